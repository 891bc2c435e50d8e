#!/bin/bash
# rotated staging writes (v21): parity first, then a same-box A/B against v15 (build/alt), then the wait counters
set -u
P=amd-versal-phylogenetic-likelihood-function_b200
mkdir -p gpurun_out
cp $P/libb200plf.so /tmp/lib_shipped.so
timeout 600 python -m pytest tests/test_protein_tc.py tests/test_states_api.py -m gpu -q > gpurun_out/c61_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c61_pytest.log
for round in 1 2 3; do
  for v in v15 v21; do
    cp $P/build/alt/lib_$v.so $P/libb200plf.so
    echo "== $v round $round"; timeout 120 python tools/tc_check.py time 2>&1 | sed -n 2p | cut -c1-120
  done
done > gpurun_out/c61_ab.log 2>&1
cat gpurun_out/c61_ab.log
cp /tmp/lib_shipped.so $P/libb200plf.so
PLF_TC_TRACE=gpurun_out/c61_tc_trace.txt timeout 120 python tools/tc_check.py time > /dev/null 2>&1
