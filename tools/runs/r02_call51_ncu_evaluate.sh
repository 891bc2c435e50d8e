#!/bin/bash
# ncu --set full of the shipped plf_evaluate_kernel<4,4> (after the same command without ncu)
set -u
mkdir -p gpurun_out
python tools/ncu_targets.py evaluate > gpurun_out/c51_eval.log 2>&1 && \
ncu --clock-control none --set full --import-source on -k regex:plf_evaluate_kernel -s 1 -c 1 -o gpurun_out/c51_evaluate python tools/ncu_targets.py evaluate > gpurun_out/c51_ncu_eval.log 2>&1
echo "evaluate rc=$?"; tail -3 gpurun_out/c51_eval.log
