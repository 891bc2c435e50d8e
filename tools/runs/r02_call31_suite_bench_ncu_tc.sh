#!/bin/bash
# full GPU suite, smoke, default bench, timing of the tcgen05 20-state kernel, then ONE ncu capture of it
# (after the same command has run without ncu), and the SASS census of the shipped object
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c31_pytest.log
tail -4 gpurun_out/c31_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c31_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c31_smoke.log
timeout 600 python bench.py > gpurun_out/c31_bench_n1.json 2> gpurun_out/c31_bench_n1.err; echo "bench rc=$?"
timeout 120 python tools/tc_check.py time > gpurun_out/c31_tc_time.log 2>&1 && \
timeout 600 ncu --clock-control none --set full --import-source on -k regex:plf_newview_aa_tc -s 2 -c 1 -o gpurun_out/c31_tc python tools/tc_check.py time > gpurun_out/c31_ncu_tc.log 2>&1
echo "ncu tc rc=$?"; head -3 gpurun_out/c31_tc_time.log
PLF_TC_TRACE=gpurun_out/c31_tc_trace.txt timeout 120 python tools/tc_check.py time > /dev/null 2>&1
cuobjdump -sass amd-versal-phylogenetic-likelihood-function_b200/build/plf_protein_tc.o | grep -E "UTCHMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTCATOM|UTMACMDFLUSH|UTMACCTL" | sed 's/^ *//' | awk '{ $1=""; print }' | sed 's/\[.*//; s/ R[0-9]*.*//; s/ UR[0-9]*.*//' | sort | uniq -c | sort -rn | head -30 > gpurun_out/c31_tc_sass.txt
