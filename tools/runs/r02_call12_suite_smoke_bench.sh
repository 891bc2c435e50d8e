#!/bin/bash
# full GPU suite, smoke, default bench, SASS census of the tensor-core object (mid-round)
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest.log
tail -6 gpurun_out/c12_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c12_smoke.log
timeout 600 python bench.py > gpurun_out/c12_bench_n1.json 2> gpurun_out/c12_bench_n1.err; echo "bench rc=$?"
cuobjdump -sass amd-versal-phylogenetic-likelihood-function_b200/build/plf_protein_tc.o | grep -E "UTCHMMA|LDTM|STTM|UTMALDG|UTCBAR|UTCATOM" | sed 's/^ *//' | awk '{ $1=""; print }' | sort | uniq -c | sort -rn | head -20 > gpurun_out/c12_tc_sass.txt
