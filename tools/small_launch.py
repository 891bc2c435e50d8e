#!/usr/bin/env python
"""tools/small_launch.py -- where do the microseconds of a SMALL launch go?  (developer tool)

cfg2 of BASELINE.json is one PLF instance over 1 Mi sites: 193 MB of traffic, 27-29 us at the HBM rate, and the
launch's fixed costs (launch gap, ramp, tail) are a third of that.  This tool times, for a list of kernel shapes:

  serial   K back-to-back launches on ONE stream between two events (no events in between), over `sets` rotating
           buffer sets (cold: 6 sets = 1.1 GB > L2) or one set (warm), with and without programmatic dependent
           launch, with the fenced and the data-dependency slot release;
  streams  the same K launches spread over S streams with their own buffers (independent instances,
           NUM_ACCELERATORS of the reference): wall time between a device-wide sync on both sides.

    python tools/small_launch.py --sites 1048576 --reps 200 --out gpurun_out/small.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SHAPES = [(0, 0), (1432, 512), (1422, 512), (2334, 256), (2332, 256), (2632, 256), (2432, 256), (3332, 128), (3632, 128),
          (1334, 256), (2322, 256)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--streams", type=int, default=9)
    ap.add_argument("--out", default="")
    ap.add_argument("--shapes", default="", help="comma list of variant:threads")
    ap.add_argument("--math", type=int, default=0)
    args = ap.parse_args()

    import torch
    pkg = bench.load_pkg()
    dev = torch.device("cuda", 0)
    n = args.sites
    peak, _ = bench.measured_peak()
    ev, left, right = bench.stimulus_matrices(42)
    d_ev, d_pl, d_pr = (torch.from_numpy(a).to(dev) for a in (ev, left, right))
    nsets = max(6, args.streams)
    x1 = [torch.empty((n, 16), device=dev) for _ in range(nsets)]
    x2 = [torch.empty((n, 16), device=dev) for _ in range(nsets)]
    x3 = [torch.empty((n, 16), device=dev) for _ in range(nsets)]
    sc = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(nsets)]
    dsum = torch.zeros(1, dtype=torch.int64, device=dev)
    main_stream = torch.cuda.current_stream().cuda_stream
    for k in range(nsets):
        pkg.generate_device(x1[k].data_ptr(), x2[k].data_ptr(), 0, n, 42 + k, main_stream)
    torch.cuda.synchronize()
    ref = None
    shapes = [tuple(int(v) for v in c.split(":")) for c in args.shapes.split(",")] if args.shapes else SHAPES
    streams = [torch.cuda.Stream() for _ in range(args.streams)]

    # yardstick: torch's copy kernel at the same size (reads 64 MiB, writes 64 MiB per launch)
    y = torch.empty_like(x1[0])
    for k in range(6):
        y.copy_(x1[k % 6])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.reps):
        y.copy_(x1[k % 6])
    e1.record()
    torch.cuda.synchronize()
    copy_us = e0.elapsed_time(e1) / args.reps * 1e3
    print(f"torch copy_ of {n * 64 >> 20} MiB: {copy_us:.2f} us per launch = {2 * n * 64 / copy_us / 1e3:.0f} GB/s", flush=True)
    del y

    rows = []

    def args_for(k, opts, stream):
        return (x1[k].data_ptr(), x2[k].data_ptr(), x3[k].data_ptr(), sc[k].data_ptr(), d_ev.data_ptr(), d_pl.data_ptr(),
                d_pr.data_ptr(), None, n, dsum.data_ptr(), opts, stream)

    for variant, threads in shapes:
        for flags, flabel in ((0, "default"), (pkg.LAUNCH_NO_PDL, "no-pdl"), (pkg.LAUNCH_DEP_RELEASE, "dep-release"),
                              (pkg.LAUNCH_FENCED_RELEASE, "fenced-release")):
            opts = pkg.make_opts(args.math, variant, threads, 0, 0, flags)
            try:
                info = pkg.kernel_info(variant, args.math, threads)
            except Exception as e:
                print(f"v={variant} t={threads}: {e}", flush=True)
                break
            row = {"variant": variant, "threads": threads, "flags": flabel, "regs": info["regs"], "bps": info["blocks_per_sm"]}
            for label, sets in (("cold", 6), ("warm", 1)):
                for i in range(sets + 3):
                    pkg.newview_device(*args_for(i % sets, opts, main_stream))
                torch.cuda.synchronize()
                dsum.zero_()
                e0.record()
                for i in range(args.reps):
                    pkg.newview_device(*args_for(i % sets, opts, main_stream))
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / args.reps * 1e3
                ok = int(dsum.item()) == args.reps * ((n + 3) // 4)
                if args.math == 0:
                    if ref is None:
                        ref = x3[0].clone()
                    ok = ok and bool(torch.equal(x3[0].view(torch.int32), ref.view(torch.int32)))
                row[label + "_us"] = us
                row[label + "_gbs"] = 193 * n / us / 1e3
                row["ok"] = row.get("ok", True) and ok
            if flabel == "default":
                # independent instances: S streams, own buffers each
                for i in range(2 * args.streams):
                    pkg.newview_device(*args_for(i % args.streams, opts, streams[i % args.streams].cuda_stream))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(args.reps):
                    k = i % args.streams
                    pkg.newview_device(*args_for(k, opts, streams[k].cuda_stream))
                issue = time.perf_counter() - t0
                torch.cuda.synchronize()
                wall = time.perf_counter() - t0
                row["streams_us"] = wall / args.reps * 1e6
                row["streams_issue_us"] = issue / args.reps * 1e6
                row["streams_gbs"] = 193 * n / row["streams_us"] / 1e3
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"sites": n, "reps": args.reps, "copy_us": copy_us, "peak": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
