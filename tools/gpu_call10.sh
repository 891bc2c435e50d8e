#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_protein.py tests/test_tree.py tests/test_protein_tc.py tests/test_felsenstein.py tests/test_states_api.py tests/test_stress.py -m gpu -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
tail -6 gpurun_out/c10_pytest.log
for codes in "" "--tip-codes"; do
  timeout 300 python tools/tree_bench.py --tips 1024 --sites 131072 --reps 10 $codes --out gpurun_out/c10_tree.jsonl > /dev/null 2>> gpurun_out/c10_tree.err
  PLF_NO_TIPTIP_TABLES=1 timeout 300 python tools/tree_bench.py --tips 1024 --sites 131072 --reps 10 $codes --out gpurun_out/c10_tree_notab.jsonl > /dev/null 2>> gpurun_out/c10_tree.err
done
cat gpurun_out/c10_tree.jsonl gpurun_out/c10_tree_notab.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tip_codes'], d['ms_per_traversal'], d['hbm_gbs_per_gpu'], d['total_scalings'])"
