#!/usr/bin/env python
"""tools/sweep.py -- kernel-variant sweep on one B200 (developer tool, not part of the product).

Times plf_newview_device for a list of (variant, threads, blocks_per_sm, math) on device-resident
synthetic CLVs with CUDA events and prints sites/s, GB/s at 193 B/site and the fraction of the
measured copy peak.  Every variant's output is compared bit-for-bit with the first strict run.

    python tools/sweep.py --sites 16777216 --reps 20 --out gpurun_out/sweep.json [--grid small|full]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=16 << 20)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="")
    ap.add_argument("--grid", default="small")
    ap.add_argument("--only", default="", help="comma list of variant:threads:bps:math")
    args = ap.parse_args()

    import torch
    pkg = bench.load_pkg()
    dev = torch.device("cuda", 0)
    n = args.sites
    ev, left, right = bench.stimulus_matrices(42)
    d_ev, d_pl, d_pr = (torch.from_numpy(a).to(dev) for a in (ev, left, right))
    x1 = torch.empty((n, 16), device=dev)
    x2 = torch.empty((n, 16), device=dev)
    x3 = torch.empty((n, 16), device=dev)
    ref3 = None
    sc = torch.empty(n, dtype=torch.uint8, device=dev)
    dsum = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    pkg.generate_device(x1.data_ptr(), x2.data_ptr(), 0, n, 42, stream)
    peak, _ = bench.measured_peak()

    # plain copy kernel of torch as a same-run yardstick (read n*128 B, write n*128 B)
    y = torch.empty_like(x1)
    for _ in range(3):
        y.copy_(x1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        y.copy_(x1)
    e1.record()
    torch.cuda.synchronize()
    copy_gbs = 2 * n * 64 * args.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del y
    print(f"torch copy yardstick: {copy_gbs:.0f} GB/s   (MEASURED_PEAKS hbm_gbs = {peak:.0f})", flush=True)

    if args.only:
        combos = [tuple(int(v) for v in c.split(":")) for c in args.only.split(",")]
    elif args.grid == "small":
        combos = []
        for math in (0, 1):
            for v, t in ((1, 256), (2, 256), (4, 256), (2002, 256), (3002, 256), (3001, 256), (4001, 256),
                         (12, 256), (2, 128), (2, 512), (2002, 512),
                         (2422, 256), (1422, 256), (3422, 256), (2421, 256), (3421, 256), (2322, 256),
                         (2622, 256), (1424, 256), (2324, 256), (1422, 512), (1622, 512), (1421, 512), (1821, 512),
                         (4422, 128), (4822, 128), (3421, 128), (4821, 128)):
                combos.append((v, t, 0, math))
    elif args.grid == "cand":
        combos = []
        for math in (0, 1):
            for v, t in ((1322, 512), (1422, 512), (1622, 512), (1222, 512), (1321, 512), (1421, 512), (1621, 512),
                         (1324, 512), (2322, 128), (3322, 128), (3222, 128), (1324, 256), (2322, 256), (2222, 256),
                         (1422, 256), (1424, 128), (3002, 256), (2, 256)):
                combos.append((v, t, 0, math))
    elif args.grid == "dyn":        # static vs dynamic stage scheduling, same shapes
        combos = []
        for math in (0, 1):
            for d, u, t, b in ((3, 2, 512, 1), (4, 2, 512, 1), (2, 2, 512, 1), (3, 4, 256, 1), (3, 2, 128, 2), (3, 2, 128, 3),
                               (3, 2, 256, 2), (4, 1, 512, 1), (6, 1, 512, 1)):
                for kind in (2, 3):
                    combos.append((1000 * b + 100 * d + 10 * kind + u, t, 0, math))
    elif args.grid == "tma":
        combos = []
        for math in (0, 1):
            for t in (128, 256, 512):
                for u in (1, 2, 4):
                    for d in (2, 3, 4, 6, 8):
                        combos.append((1000 + 100 * d + 20 + u, t, 0, math))
            for t, b in ((128, 2), (128, 3), (256, 2)):
                for u in (1, 2):
                    for d in (2, 3, 4):
                        combos.append((1000 * b + 100 * d + 20 + u, t, 0, math))
    else:
        combos = []
        for math in (0, 1):
            for u in (1, 2, 4):
                for b in (1, 2, 3, 4):
                    for t in (128, 256, 512):
                        combos.append((1000 * b + u, t, 0, math))
                for d in (2, 3, 4, 6, 8):
                    for b, t in ((1, 128), (2, 128), (3, 128), (4, 128), (1, 256), (2, 256), (3, 256), (1, 512)):
                        combos.append((1000 * b + 100 * d + 20 + u, t, 0, math))

    rows = []
    for variant, threads, bps, math in combos:
        opts = pkg.make_opts(math, variant, threads, bps)
        try:
            info = pkg.kernel_info(variant, math, threads)
            x3.zero_()
            dsum.zero_()
            a = (x1.data_ptr(), x2.data_ptr(), x3.data_ptr(), sc.data_ptr(), d_ev.data_ptr(), d_pl.data_ptr(),
                 d_pr.data_ptr(), None, n, dsum.data_ptr(), opts, stream)
            for _ in range(3):
                pkg.newview_device(*a)
            torch.cuda.synchronize()
            ok = int(dsum.item()) == 3 * ((n + 3) // 4)
            if math == 0:
                if ref3 is None:
                    ref3 = x3.clone()
                else:
                    ok = ok and bool(torch.equal(x3.view(torch.int32), ref3.view(torch.int32)))
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
            evs[0].record()
            for i in range(args.reps):
                pkg.newview_device(*a)
                evs[i + 1].record()
            torch.cuda.synchronize()
            ts = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.reps))
            mean = evs[0].elapsed_time(evs[-1]) / args.reps
            med = ts[len(ts) // 2]
            gbs = 193 * n / (mean * 1e-3) / 1e9
            row = {"variant": variant, "threads": threads, "bps": bps or info["blocks_per_sm"], "math": math,
                   "regs": info["regs"], "ms_mean": mean, "ms_med": med, "ms_min": ts[0],
                   "gbs": gbs, "frac_measured": gbs / peak, "gsites": n / (mean * 1e-3) / 1e9, "ok": ok}
        except Exception as e:  # unknown variant / too much smem
            row = {"variant": variant, "threads": threads, "bps": bps, "math": math, "error": str(e)[:80]}
        rows.append(row)
        if "error" in row:
            print(f"v={variant:5d} t={threads:3d} math={math}  ERROR {row['error']}", flush=True)
        else:
            print(f"v={variant:5d} t={threads:3d} bps={row['bps']:2d} math={math} regs={row['regs']:3d} "
                  f"mean={mean:7.4f} ms min={ts[0]:7.4f}  {gbs:7.0f} GB/s  {row['frac_measured']:.3f} of measured  "
                  f"{'ok' if ok else 'MISMATCH'}", flush=True)
    # INPUT_SRC=gen analogue: cfg4a (write) and cfg4b (discard = pure issue rate, no HBM traffic)
    gen_rows = []
    chk = torch.zeros(1, dtype=torch.float64, device=dev)
    for math in (0, 1):
        for sink, label in ((pkg.GEN_WRITE, "write"), (pkg.GEN_DISCARD, "discard")):
            opts = pkg.make_opts(math)
            a = (x3.data_ptr(), sc.data_ptr(), n, dsum.data_ptr(), chk.data_ptr(), sink, opts, stream)
            for _ in range(3):
                pkg.newview_gen_device(*a)
            e0.record()
            for _ in range(args.reps):
                pkg.newview_gen_device(*a)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            gs = n / (ms * 1e-3) / 1e9
            gen_rows.append({"math": math, "sink": label, "ms": ms, "gsites": gs,
                             "gbs_written": (65 * n / (ms * 1e-3) / 1e9) if sink == pkg.GEN_WRITE else 0.0})
            print(f"gen math={math} sink={label:7s} {ms:7.4f} ms  {gs:6.2f} G sites/s", flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"sites": n, "reps": args.reps, "copy_gbs": copy_gbs, "peak": peak, "rows": rows, "gen": gen_rows}, f, indent=1)


if __name__ == "__main__":
    main()
