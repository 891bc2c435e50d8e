#!/bin/bash
# same-box A/B of two builds of the tcgen05 kernel, alternating.  The alternative libraries are built beforehand into
# build/alt/lib_<name>.so (the other objects of build/ linked with an alternative plf_protein_tc.o) and are not kept.
set -u
P=amd-versal-phylogenetic-likelihood-function_b200
mkdir -p gpurun_out
for round in 1 2 3; do
  for v in v15 v17; do
    cp $P/build/alt/lib_$v.so $P/libb200plf.so
    echo "== $v round $round"; timeout 120 python tools/tc_check.py time 2>&1 | head -2 | cut -c1-120
  done
done > gpurun_out/c47_ab.log 2>&1
cat gpurun_out/c47_ab.log
