#!/bin/bash
# the ring kernels' test files (incl. the ring-fed evaluate kernel and the final tensor-core kernel) under both
# process-wide slot-release modes
set -u
mkdir -p gpurun_out
for mode in 1 0; do
  PLF_SAFE_RELEASE=$mode timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tree.py tests/test_protein.py tests/test_protein_tc.py tests/test_stress.py tests/test_states_api.py tests/test_evaluate.py tests/test_felsenstein.py -m gpu -q > gpurun_out/c59_pytest_release$mode.log 2>&1
  echo "PLF_SAFE_RELEASE=$mode rc=$?"; tail -2 gpurun_out/c59_pytest_release$mode.log
done
