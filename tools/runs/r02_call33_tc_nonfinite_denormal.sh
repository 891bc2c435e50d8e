#!/bin/bash
# tests of the tcgen05 kernel with non-finite entries; denormal operand probe
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_protein_tc.py -m gpu -q > gpurun_out/c33_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c33_pytest.log
timeout 120 python tools/tc_denormal_probe.py > gpurun_out/c33_denormal.log 2>&1; cat gpurun_out/c33_denormal.log | tail -5
