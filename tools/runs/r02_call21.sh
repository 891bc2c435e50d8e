#!/bin/bash
# 20-state, stress and host tests after the release-default change, timing of the tcgen05 kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_protein.py tests/test_protein_tc.py tests/test_states_api.py tests/test_stress.py tests/test_hosts.py -m gpu -q > gpurun_out/c21_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c21_pytest.log
timeout 120 python tools/tc_check.py time > gpurun_out/c21_tc_time.log 2>&1; head -3 gpurun_out/c21_tc_time.log
