#!/bin/bash
# full GPU suite, smoke, default bench (the round-end sequence of the driver)
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c62_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c62_pytest.log
tail -4 gpurun_out/c62_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c62_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c62_smoke.log
timeout 600 python bench.py > gpurun_out/c62_bench_n1.json 2> gpurun_out/c62_bench_n1.err; echo "bench rc=$?"
