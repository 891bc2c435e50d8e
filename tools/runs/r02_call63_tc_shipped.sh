#!/bin/bash
# the shipped tcgen05 kernel (converged issuer warps, per-warp whole-row staging with rotated chunk order): 20-state tests, stress, timing,
# wait counters, then ONE ncu capture (after the same command without ncu), SASS census
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_protein.py tests/test_protein_tc.py tests/test_states_api.py tests/test_stress.py tests/test_hosts.py -m gpu -q > gpurun_out/c63_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c63_pytest.log
timeout 120 python tools/tc_check.py time > gpurun_out/c63_tc_time.log 2>&1 && \
timeout 600 ncu --clock-control none --set full --import-source on -k regex:plf_newview_aa_tc -s 2 -c 1 -o gpurun_out/c63_tc python tools/tc_check.py time > gpurun_out/c63_ncu_tc.log 2>&1
echo "ncu tc rc=$?"; head -3 gpurun_out/c63_tc_time.log | cut -c1-130
PLF_TC_TRACE=gpurun_out/c63_tc_trace.txt timeout 120 python tools/tc_check.py time > /dev/null 2>&1
cuobjdump -sass amd-versal-phylogenetic-likelihood-function_b200/build/plf_protein_tc.o | grep -E "UTCHMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTCATOM|UTMACMDFLUSH|ELECT" | sed 's/^ *//' | awk '{ $1=""; print }' | sed 's/\[.*//; s/ R[0-9]*.*//; s/ UR[0-9]*.*//; s/ P[0-9].*//' | sort | uniq -c | sort -rn | head -30 > gpurun_out/c63_tc_sass.txt
