#!/usr/bin/env python
"""tools/tree_bench.py -- throughput of the chained newview over a synthetic tree (cfg5).

    python tools/tree_bench.py --tips 1024 --sites 131072 --shape balanced --reps 20
    torchrun --nproc-per-node 8 tools/tree_bench.py --tips 1024 --sites 1048576   # sites split over ranks

Tips are generated on the device (counter hash), matrices are random; every rank traverses the
whole tree on its own site range; the per-rank scaling totals are all-reduced (NCCL) at the end.
Reports the device time per traversal (max over ranks), newview-sites/s and achieved HBM GB/s at
the algorithmic byte count plf_tree_info returns (192 B/site/node + 4 B per count vector touched)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tips", type=int, default=1024)
    ap.add_argument("--sites", type=int, default=1 << 20, help="total sites (split over ranks)")
    ap.add_argument("--shape", default="balanced", choices=["balanced", "random", "caterpillar"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--u", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--math", default="strict")
    ap.add_argument("--tip-codes", action="store_true", help="tips as 1 B/site state codes + tip vector table")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    pkg = bench.load_pkg()
    from plf_b200 import sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    first, n = sharding.shard_for_rank(args.sites, rank, world)
    if args.shape == "caterpillar":      # one op per level; every op but the first has one inner child
        left = np.array([0] + [args.tips + k for k in range(args.tips - 2)], np.int32)
        right = np.arange(1, args.tips, dtype=np.int32)
    else:
        left, right = (pkg.balanced_tree if args.shape == "balanced" else pkg.random_tree)(args.tips)
    rng = np.random.RandomState(1)
    def stochastic(*shape):      # rows of every 4x4 matrix sum to 1: CLV magnitudes stay bounded
        m = rng.random_sample(shape + (4, 4)) + 0.05
        return (m / m.sum(axis=-1, keepdims=True)).astype(np.float32)
    ev = stochastic().reshape(16)
    pl = stochastic(args.tips - 1, 4).reshape(args.tips - 1, 64)
    pr = stochastic(args.tips - 1, 4).reshape(args.tips - 1, 64)

    t = pkg.Tree(left, right, n, device=local, tip_codes=args.tip_codes)
    t.set_tuning(args.u, args.chunk)
    t.set_math(pkg.MATH_FMA if args.math == "fma" else pkg.MATH_STRICT)
    if args.tip_codes:
        tv = (rng.random_sample((16, 4)) * np.where(np.arange(16)[:, None] % 4 == 0, 1e-10, 1.0)).astype(np.float32)
        t.write_tip_vector(tv)
        for tip in range(args.tips):
            t.write_tip_codes(tip, np.random.RandomState(first + tip).randint(0, 16, n).astype(np.uint8))
    else:
        scratch = torch.empty((n, 16), device=device)
        for tip in range(args.tips):      # x1-stream of the generator for even tips, x2-stream for odd ones
            a, b = (t.tip_ptr(tip), scratch.data_ptr()) if tip % 2 == 0 else (scratch.data_ptr(), t.tip_ptr(tip))
            pkg.generate_device(a, b, first + tip * 7919, n, 1000 + tip)
        torch.cuda.synchronize()
    t.write_matrices(ev, pl, pr)
    info = t.info()
    for _ in range(3):
        t.run_async()
    t.wait()
    times = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        t.run_async()
        t.wait()
        times.append(t.last_ms())
    ms = float(np.median(times))
    ms_max = sharding.max_over_ranks(ms, device)
    total = sharding.reduce_scaler_increment(t.total_scalings(), device)
    # the optional final reduction of the path: per-rank log-likelihood across the root branch, summed over NCCL
    diag = np.exp(-np.linspace(0.0, 1.5, 16)).astype(np.float32)
    try:
        lnl = sharding.reduce_log_likelihood(t.evaluate_root(diag), device)
    except pkg.PlfError:          # a child of the root is a compressed tip: no dense CLV to evaluate
        lnl = None
    root, cnt = t.read_root(0, min(n, 4096))
    if rank == 0:
        nodes = args.tips - 1
        row = {"tips": args.tips, "sites_total": args.sites, "sites_per_gpu": n, "gpus": world, "shape": args.shape,
               "levels": info["levels"], "clv_slots": info["clv_slots"], "device_GiB": info["device_bytes"] / 2 ** 30,
               "ms_per_traversal": ms_max, "ms_min": float(min(times)),
               "newview_sites_per_s": nodes * args.sites / (ms_max * 1e-3),
               "hbm_gbs_per_gpu": info["traversal_bytes"] / (ms_max * 1e-3) / 1e9,
               "total_scalings": total, "log_likelihood": lnl, "root_count_max": int(cnt.max()), "root_finite": bool(np.isfinite(root).all()),
               "u": args.u, "chunk": args.chunk, "math": args.math, "tip_codes": args.tip_codes,
               "traversal_GB": info["traversal_bytes"] / 1e9}
        print(json.dumps(row), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(row) + "\n")
    t.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
