#!/bin/bash
# round-2 GPU call 4 (one B200): ncu captures of the shipped kernels.  Every ncu run is preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
TB="python tools/tree_bench.py --tips 1024 --sites 131072 --reps 1"
# tree, dense tips: level 1 (512 tip-tip ops on dense CLVs) and level 2 (inner-inner)
$TB > gpurun_out/c4_tree_dense.log 2>&1 && \
$NCU --set full --import-source on -k regex:plf_newview_batch -s 30 -c 2 -o gpurun_out/c4_tree_dense $TB > gpurun_out/c4_ncu_tree_dense.log 2>&1
echo "tree dense rc=$?"
# tree, state-code tips: level 1 (tip-tip from codes) and level 2
$TB --tip-codes > gpurun_out/c4_tree_codes.log 2>&1 && \
$NCU --set full --import-source on -k regex:plf_newview_batch -s 30 -c 2 -o gpurun_out/c4_tree_codes $TB --tip-codes > gpurun_out/c4_ncu_tree_codes.log 2>&1
echo "tree codes rc=$?"
# evaluate kernel
python tools/ncu_targets.py evaluate > gpurun_out/c4_eval.log 2>&1 && \
$NCU --set full --import-source on -k regex:plf_evaluate_kernel -s 1 -c 1 -o gpurun_out/c4_evaluate python tools/ncu_targets.py evaluate > gpurun_out/c4_ncu_eval.log 2>&1
echo "evaluate rc=$?"
# headline kernel (64 Mi sites), refreshed: the dynamically scheduled ring with the static prologue
python tools/ncu_targets.py cfg3 > gpurun_out/c4_cfg3.log 2>&1 && \
$NCU --set full --import-source on -k regex:plf_newview_tma_dyn -s 1 -c 1 -o gpurun_out/c4_cfg3 python tools/ncu_targets.py cfg3 > gpurun_out/c4_ncu_cfg3.log 2>&1
echo "cfg3 rc=$?"
# cfg2: per-launch device time of 1 Mi-site launches (share of the 33.4 us per launch seen with events)
python tools/ncu_targets.py cfg2 > gpurun_out/c4_cfg2.log 2>&1 && \
$NCU --metrics gpu__time_duration.sum -k regex:plf_newview_tma_dyn -s 10 -c 30 --csv --log-file gpurun_out/c4_cfg2_launches.csv python tools/ncu_targets.py cfg2 > gpurun_out/c4_ncu_cfg2.log 2>&1
echo "cfg2 rc=$?"
# launch list of the default bench command (kernel share of the timed region)
python bench.py --steps 3 --warmup 3 --no-side --no-cpu-baseline > gpurun_out/c4_bench_plain.json 2> gpurun_out/c4_bench_plain.err && \
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/c4_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-side --no-cpu-baseline > gpurun_out/c4_ncu_bench.log 2>&1
echo "bench launches rc=$?"
ls -la gpurun_out/*.ncu-rep
echo done
