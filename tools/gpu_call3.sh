#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
tail -8 gpurun_out/c3_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err; echo "bench rc=$?"
echo done
