#!/usr/bin/env python
"""tools/size_curve.py -- time per launch of the default newview kernel against the number of sites (developer tool).

K back-to-back launches on one stream between two events, one buffer set (inputs far larger than L2 from 1 Mi sites
up), for a few kernel shapes and with / without programmatic dependent launch.  Fits T(n) = a + n / b over the large
sizes and prints each size's excess over the fit: where the strong-scaling efficiency of the 8-GPU run goes.

    python tools/size_curve.py [--shapes 0:0,1432:512,1332:512] [--out gpurun_out/size_curve.json]
"""
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
    ap.add_argument("--shapes", default="0:0")
    ap.add_argument("--sizes", default="1,2,4,8,16,32,64", help="Mi sites")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    pkg = bench.load_pkg()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    ev, left, right = bench.stimulus_matrices(42)
    d_ev, d_pl, d_pr = (torch.from_numpy(a).to(dev) for a in (ev, left, right))
    sizes = [int(float(v) * (1 << 20)) for v in args.sizes.split(",")]
    nmax = max(sizes)
    x1 = torch.empty((nmax, 16), device=dev)
    x2 = torch.empty((nmax, 16), device=dev)
    x3 = torch.empty((nmax, 16), device=dev)
    sc = torch.empty(nmax, dtype=torch.uint8, device=dev)
    dsum = torch.zeros(1, dtype=torch.int64, device=dev)
    pkg.generate_device(x1.data_ptr(), x2.data_ptr(), 0, nmax, 42, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for shape in args.shapes.split(","):
        variant, threads = (int(v) for v in shape.split(":"))
        for flags, flabel in ((0, "pdl"), (pkg.LAUNCH_NO_PDL, "no-pdl")):
            opts = pkg.make_opts(0, variant, threads, 0, 0, flags)
            ts = []
            for n in sizes:
                reps = max(10, int(40e-3 / (n * 193 / 7e12)))
                a = (x1.data_ptr(), x2.data_ptr(), x3.data_ptr(), sc.data_ptr(), d_ev.data_ptr(), d_pl.data_ptr(), d_pr.data_ptr(),
                     None, n, dsum.data_ptr(), opts, stream)
                for _ in range(5):
                    pkg.newview_device(*a)
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(3):
                    e0.record()
                    for _ in range(reps):
                        pkg.newview_device(*a)
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / reps * 1e3)
                ts.append(best)
            ts = np.array(ts)
            ns = np.array(sizes, dtype=np.float64)
            big = ns >= 16 * (1 << 20)
            slope, icpt = np.polyfit(ns[big], ts[big], 1) if big.sum() >= 2 else (ts[-1] / ns[-1], 0.0)
            row = {"variant": variant, "threads": threads, "flags": flabel, "fit_gbs": 193 / slope / 1e3, "fit_intercept_us": icpt,
                   "sizes_Mi": [n / (1 << 20) for n in sizes], "us": ts.tolist(), "gbs": (193 * ns / ts / 1e3).tolist(),
                   "excess_over_slope_us": (ts - slope * ns).tolist()}
            out.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
