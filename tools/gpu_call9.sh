#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
tail -6 gpurun_out/c9_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c9_bench_n1.json 2> gpurun_out/c9_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/c9_bench_ref.json 2> gpurun_out/c9_bench_ref.err; echo "ref rc=$?"
