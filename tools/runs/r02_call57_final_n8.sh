#!/bin/bash
# final eight-GPU record of the round: the driver's scaling sequence N = 1, 2, 4, 8 on ONE box (so the efficiency is not
# a comparison between boxes), the reference arm at N = 8, the multi-GPU tests
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-side > gpurun_out/c57_bench_n1.json 2> gpurun_out/c57_bench_n1.err; echo "bench N=1 rc=$?"
for N in 2 4 8; do
  extra=""; if [ $N != 8 ]; then extra="--no-side"; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 $extra > gpurun_out/c57_bench_n$N.json 2> gpurun_out/c57_bench_n$N.err; echo "bench N=$N rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/c57_bench_ref_n8.json 2> gpurun_out/c57_bench_ref_n8.err; echo "ref rc=$?"
timeout 300 python -m pytest tests/test_multi.py -m gpu -q > gpurun_out/c57_pytest_multi.log 2>&1; tail -2 gpurun_out/c57_pytest_multi.log
python - <<'P'
import json
v = {}
for n in (1, 2, 4, 8):
    d = json.loads(open(f"gpurun_out/c57_bench_n{n}.json").read().strip().splitlines()[-1])
    v[n] = d["value"]
    print(n, d["value"] / 1e9, d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"] / 1e6, d["e2e"]["roofline"]["frac"])
print("efficiency", [round(v[n] / (n * v[1]), 4) for n in (1, 2, 4, 8)])
P
