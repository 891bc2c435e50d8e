#!/usr/bin/env python
"""tools/ncu_targets.py -- small, fixed workloads to put under ncu (developer tool).

    python tools/ncu_targets.py evaluate      # 3 launches of plf_evaluate_kernel over 16 Mi sites (with scaler counts)
    python tools/ncu_targets.py cfg2          # 40 back-to-back 1 Mi-site newview launches on one stream (6 buffer sets)
    python tools/ncu_targets.py cfg3          # 3 launches of the default newview kernel over 64 Mi sites
    python tools/ncu_targets.py protein       # 2 launches each of the strict and the FMA 20-state kernel, 2 Mi sites
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    what = sys.argv[1]
    import torch
    pkg = bench.load_pkg()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    ev, left, right = bench.stimulus_matrices(42)
    d_ev, d_pl, d_pr = (torch.from_numpy(a).to(dev) for a in (ev, left, right))
    if what == "evaluate":
        n = 16 << 20
        x1 = torch.empty((n, 16), device=dev)
        x2 = torch.empty((n, 16), device=dev)
        pkg.generate_device(x1.data_ptr(), x2.data_ptr(), 0, n, 42, stream)
        c1 = torch.ones(n, dtype=torch.int32, device=dev)
        c2 = torch.zeros(n, dtype=torch.int32, device=dev)
        diag = torch.from_numpy(np.exp(-np.linspace(0.0, 1.5, 16)).astype(np.float32)).to(dev)
        lnl = torch.zeros(4, dtype=torch.float64, device=dev)
        for i in range(3):
            pkg.evaluate_device(x1.data_ptr(), x2.data_ptr(), c1.data_ptr(), c2.data_ptr(), None, diag.data_ptr(), n,
                                lnl[i:].data_ptr(), stream)
        torch.cuda.synchronize()
        v = lnl.cpu().numpy()
        assert v[0] == v[1] == v[2] and np.isfinite(v[0])
        if "--time" in sys.argv:                                # events around 20 launches: G sites/s, GB/s (136 B/site)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20):
                pkg.evaluate_device(x1.data_ptr(), x2.data_ptr(), c1.data_ptr(), c2.data_ptr(), None, diag.data_ptr(), n,
                                    lnl[3:].data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"evaluate {n} sites: {ms:.4f} ms = {n / ms / 1e6:.2f} G sites/s = {136 * n / ms / 1e6:.0f} GB/s, lnL {v[0]!r}")
    elif what in ("cfg2", "cfg3"):
        n, sets, reps = ((1 << 20), 6, 40) if what == "cfg2" else ((64 << 20), 1, 3)
        x1 = [torch.empty((n, 16), device=dev) for _ in range(sets)]
        x2 = [torch.empty((n, 16), device=dev) for _ in range(sets)]
        x3 = [torch.empty((n, 16), device=dev) for _ in range(sets)]
        sc = torch.empty(n, dtype=torch.uint8, device=dev)
        dsum = torch.zeros(1, dtype=torch.int64, device=dev)
        for k in range(sets):
            pkg.generate_device(x1[k].data_ptr(), x2[k].data_ptr(), 0, n, 42 + k, stream)
        for i in range(reps):
            k = i % sets
            pkg.newview_device(x1[k].data_ptr(), x2[k].data_ptr(), x3[k].data_ptr(), sc.data_ptr(), d_ev.data_ptr(),
                               d_pl.data_ptr(), d_pr.data_ptr(), None, n, dsum.data_ptr(), None, stream)
        torch.cuda.synchronize()
        assert int(dsum.item()) == reps * ((n + 3) // 4)
    elif what == "protein":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import protein_bench
        res = protein_bench.measure(pkg, torch, 2 << 20, 2, shapes=[(0, 0)], maths=(0, 1), verbose=False)
        assert all(r["ok"] for r in res["rows"])
    else:
        raise SystemExit(f"unknown target {what}")
    print("ok", what)


if __name__ == "__main__":
    main()
