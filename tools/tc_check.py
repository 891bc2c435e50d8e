#!/usr/bin/env python
"""tools/tc_check.py -- the tcgen05 20-state kernel (variant 9, FMA mode) against the CPU restatement, then its speed.

    python tools/tc_check.py check       # parity on small and ragged sizes (prints the relative-error distribution)
    python tools/tc_check.py time        # G sites/s at 2 Mi sites next to the CUDA-core FMA and strict kernels
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    import torch
    import oracle
    pkg = bench.load_pkg()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.RandomState(7)
    ev, left, right = (rng.random_sample(k).astype(np.float32) for k in (400, 1600, 1600))
    if mode == "check":
        co = oracle.COracle()
        worst = 0.0
        for n in (1, 7, 128, 129, 300, 4099, 100003):
            x1, x2 = pkg.generate_states_host(20, 0, n, 11 + n)
            o3, osc, oinc = co.newview_states(20, x1, x2, ev, left, right)
            d1, d2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
            g3 = torch.full((n, 80), float("nan"), device=dev)
            gsc = torch.full((n,), 9, dtype=torch.uint8, device=dev)
            gsum = torch.zeros(1, dtype=torch.int64, device=dev)
            opts = pkg.make_opts(pkg.MATH_FMA, 9, 0)
            pkg.newview_states_device(20, d1.data_ptr(), d2.data_ptr(), g3.data_ptr(), gsc.data_ptr(), ev, left, right, None, n,
                                      gsum.data_ptr(), opts, stream)
            torch.cuda.synchronize()
            got = g3.cpu().numpy()
            rel = np.abs(got.astype(np.float64) - o3) / np.maximum(np.abs(o3), 1e-300)
            sc_same = bool(np.array_equal(gsc.cpu().numpy(), osc))
            print(json.dumps({"n": n, "finite": bool(np.isfinite(got).all()), "rel_max": float(rel.max()), "rel_p50": float(np.median(rel)),
                              "rel_p999": float(np.quantile(rel, 0.999)), "scaler_bytes_equal": sc_same,
                              "sum": int(gsum.item()), "sum_want": int(oinc)}), flush=True)
            worst = max(worst, float(rel.max()))
        print("worst relative error", worst, "PASS" if worst <= 1e-5 else "FAIL")
        return 0 if worst <= 1e-5 else 1
    n = 2 << 20
    x1 = torch.empty(n * 80, device=dev)
    x2 = torch.empty(n * 80, device=dev)
    x3 = torch.empty(n * 80, device=dev)
    sc = torch.empty(n, dtype=torch.uint8, device=dev)
    dsum = torch.zeros(1, dtype=torch.int64, device=dev)
    pkg.generate_states_device(20, x1.data_ptr(), x2.data_ptr(), 0, n, 42, stream)
    for label, math, variant, fl in (("tcgen05 3xTF32 (fenced release)", 1, 9, pkg.LAUNCH_FENCED_RELEASE),
                                     ("tcgen05 3xTF32 (dependency release)", 1, 9, pkg.LAUNCH_DEP_RELEASE),
                                     ("FMA default (= the tcgen05 kernel at this size)", 1, 0, 0), ("cuda-core fma (variant 4, 256 threads)", 1, 4, 0),
                                     ("cuda-core strict", 0, 0, 0)):
        opts = pkg.make_opts(math, variant, 256 if variant == 4 else 0, 0, 0, fl)
        a = (20, x1.data_ptr(), x2.data_ptr(), x3.data_ptr(), sc.data_ptr(), ev, left, right, None, n, dsum.data_ptr(), opts, stream)
        dsum.zero_()
        for _ in range(2):
            pkg.newview_states_device(*a)
        torch.cuda.synchronize()
        ok = int(dsum.item()) == 2 * ((n + 3) // 4)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            pkg.newview_states_device(*a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"kernel": label, "ms": ms, "gsites": n / ms / 1e6, "gbs": 961 * n / ms / 1e6, "scaler_sum_ok": ok}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
