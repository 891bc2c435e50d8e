#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/tc_check.py check > gpurun_out/c8_tc_check.log 2>&1; echo "tc check rc=$?"; tail -3 gpurun_out/c8_tc_check.log
timeout 120 python tools/tc_check.py time > gpurun_out/c8_tc_time.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:plf_newview_aa_tc -s 1 -c 1 -o gpurun_out/c8_tc python tools/tc_check.py time > gpurun_out/c8_ncu_tc.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/c8_tc_time.log
