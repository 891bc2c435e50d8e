#!/bin/bash
# the reference arm as the driver launches it at N = 1 (unmodified plf() on all host threads over the whole workload)
set -u
mkdir -p gpurun_out
SECONDS=0; timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/c65_bench_ref_n1.json 2> gpurun_out/c65_bench_ref_n1.err; echo "rc=$?"
cat gpurun_out/c65_bench_ref_n1.json | cut -c1-600; echo "wall seconds: $SECONDS"
