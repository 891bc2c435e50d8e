#!/bin/bash
# ring-fed evaluate kernel (24 consumer warps, 2 stages of 768 sites): tests that reach it, bench side record with and
# without it, then ONE ncu capture (after the same command without ncu)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_evaluate.py tests/test_felsenstein.py tests/test_tree.py tests/test_states_api.py tests/test_multi.py -m gpu -q > gpurun_out/c54_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c54_pytest.log
for v in ring noring; do
  if [ $v = noring ]; then export PLF_EVAL_NO_RING=1; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c54_bench_$v.json 2> gpurun_out/c54_bench_$v.err; echo "bench $v rc=$?"
done
unset PLF_EVAL_NO_RING
python - <<'P'
import json
for v in ("ring", "noring"):
    d = json.loads(open(f"gpurun_out/c54_bench_{v}.json").read().strip().splitlines()[-1])
    e = d["side"]["evaluate"]
    print(v, e["value"] / 1e9, e["ms_per_launch"], e["roofline"]["frac"], e["bitwise_reproducible"], e["log_likelihood_rank0"])
P
python tools/ncu_targets.py evaluate --time > gpurun_out/c54_eval.log 2>&1 && \
ncu --clock-control none --set full --import-source on -k regex:plf_evaluate_ring -s 1 -c 1 -o gpurun_out/c54_evaluate_ring python tools/ncu_targets.py evaluate > gpurun_out/c54_ncu_eval.log 2>&1
echo "ncu rc=$?"; grep "G sites" gpurun_out/c54_eval.log
