#!/bin/bash
# tools/microbench_tma.cu: feeding a shared-memory ring by TMA request shape
set -u
mkdir -p gpurun_out
timeout 120 build/microbench_tma > gpurun_out/c44_microbench_tma.txt 2>&1; echo "rc=$?"; cat gpurun_out/c44_microbench_tma.txt
