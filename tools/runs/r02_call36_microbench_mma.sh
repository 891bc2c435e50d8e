#!/bin/bash
# tools/microbench_mma.cu: cost of small tcgen05.mma by shape, issuing idiom and number of issuing warps
set -u
mkdir -p gpurun_out
timeout 60 build/microbench_mma > gpurun_out/c36_microbench_mma.txt 2>&1; echo "rc=$?"; cat gpurun_out/c36_microbench_mma.txt
