#!/bin/bash
# tcgen05 20-state kernel (staging boxes + tensor stores): parity and timing
set -u
mkdir -p gpurun_out
timeout 120 python tools/tc_check.py check > gpurun_out/c16_tc_check.log 2>&1; echo "tc check rc=$?"; tail -1 gpurun_out/c16_tc_check.log
timeout 120 python tools/tc_check.py time > gpurun_out/c16_tc_time.log 2>&1; echo "time rc=$?"; head -3 gpurun_out/c16_tc_time.log
