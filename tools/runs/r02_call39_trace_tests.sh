#!/bin/bash
# tests of the two debug instruments (stream event timeline, counting twin of the tcgen05 kernel)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_protein_tc.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/c39_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c39_pytest.log
