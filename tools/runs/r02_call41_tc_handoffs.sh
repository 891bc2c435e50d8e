#!/bin/bash
# tcgen05 20-state kernel v10 + hand-off latency counters: parity, timing, wait counters
set -u
mkdir -p gpurun_out
timeout 90 python tools/tc_check.py check > gpurun_out/c41_tc_check.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/c41_tc_check.log
timeout 120 python tools/tc_check.py time > gpurun_out/c41_tc_time.log 2>&1; head -4 gpurun_out/c41_tc_time.log
PLF_TC_TRACE=gpurun_out/c41_tc_trace.txt timeout 120 python tools/tc_check.py time > /dev/null 2>&1; ls -la gpurun_out/c41_tc_trace.txt
