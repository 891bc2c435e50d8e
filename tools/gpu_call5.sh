#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/tc_check.py check > gpurun_out/c5_tc_check.log 2>&1; echo "tc check rc=$?"
tail -12 gpurun_out/c5_tc_check.log
timeout 120 python tools/tc_check.py time > gpurun_out/c5_tc_time.log 2>&1; echo "tc time rc=$?"
cat gpurun_out/c5_tc_time.log | tail -5
nvidia-smi --query-gpu=name,memory.used --format=csv
