#!/bin/bash
# ring-depth sensitivity of the tcgen05 20-state kernel: alternative builds (build/alt, -DPLF_TC_RING=n) swapped in on the box
set -u
P=amd-versal-phylogenetic-likelihood-function_b200
mkdir -p gpurun_out
cp $P/libb200plf.so /tmp/lib_default.so
for v in default ring3 ring2; do
  if [ $v = default ]; then cp /tmp/lib_default.so $P/libb200plf.so; else cp $P/build/alt/lib_$v.so $P/libb200plf.so; fi
  echo "== $v"; timeout 120 python tools/tc_check.py time 2>&1 | head -2
done > gpurun_out/c25_ring.log 2>&1
cat gpurun_out/c25_ring.log
