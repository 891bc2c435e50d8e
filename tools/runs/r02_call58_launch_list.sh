#!/bin/bash
# the streamed-path tests after the last ABI edit, then the ncu launch list of the default bench command (after the same
# command without ncu): which kernels make up the timed region and their share of it
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_hosts.py -m gpu -q > gpurun_out/c58_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c58_pytest.log
python bench.py --steps 3 --warmup 3 --no-side --no-cpu-baseline > gpurun_out/c58_bench_plain.json 2> gpurun_out/c58_bench_plain.err && \
ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/c58_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-side --no-cpu-baseline > gpurun_out/c58_ncu_bench.log 2>&1
echo "launch list rc=$?"
