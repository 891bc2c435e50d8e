#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/tc_check.py time > gpurun_out/c6_tc_time.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:plf_newview_aa_tc -s 1 -c 1 -o gpurun_out/c6_tc python tools/tc_check.py time > gpurun_out/c6_ncu_tc.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/c6_tc_time.log
timeout 200 python tools/ncu_targets.py evaluate && timeout 300 python - <<'P'
import sys, os
sys.path.insert(0, os.getcwd())
import bench, torch, numpy as np
pkg = bench.load_pkg()
print(bench.side_evaluate(pkg, torch, torch.device("cuda",0), 6543.4, 64<<20, 1, __import__("plf_b200").sharding))
P
