#!/bin/bash
# small-launch study after the producer-side finish count: tests, 1 Mi-site launches, 8-shape sweep
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tree.py tests/test_stress.py -m gpu -q -x > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c15_pytest.log
timeout 300 python tools/small_launch.py --shapes 0:0,1432:512,2332:256 --out gpurun_out/c15_small.json > gpurun_out/c15_small.log 2>&1; echo "small rc=$?"; grep -E '"flags": "default"' gpurun_out/c15_small.log | cut -c1-330
timeout 300 python tools/sweep.py --sites 8388608 --reps 60 --only 1432:512:0:0 > gpurun_out/c15_sweep8.log 2>&1; grep "v=" gpurun_out/c15_sweep8.log
timeout 300 python tools/tree_bench.py --tips 1024 --sites 131072 --reps 10 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tree dense', d['ms_per_traversal'], d['hbm_gbs_per_gpu'])"
timeout 300 python tools/tree_bench.py --tips 1024 --sites 131072 --reps 10 --tip-codes | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tree codes', d['ms_per_traversal'], d['hbm_gbs_per_gpu'])"
