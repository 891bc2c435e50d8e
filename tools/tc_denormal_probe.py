#!/usr/bin/env python
"""tools/tc_denormal_probe.py -- what the tensor-core 20-state kernel does with denormal operands (developer tool).

Scales x1 of every site into the fp32 denormal range and compares the tcgen05 kernel and the CUDA-core FMA kernel with
the CPU loop nest.  Prints the fraction of results that are zero / within 1e-5 / off."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    import torch
    import oracle
    from test_protein import matrices, run_states
    pkg = bench.load_pkg()
    co = oracle.COracle()
    n = 4096
    ev, left, right = matrices(7)
    x1, x2 = pkg.generate_states_host(20, 0, n, 5)
    x1 = x1.reshape(n, 80) * np.float32(1e-30)          # 1e-30 .. 1e-42: around and below the smallest normal (1.2e-38)
    x2 = x2.reshape(n, 80)
    o3, osc, _ = co.newview_states(20, x1, x2, ev, left, right)
    for label, shape in (("tcgen05", (9, 0)), ("cuda-core fma", (4, 256))):
        g3, gsc, _ = run_states(pkg, torch, 20, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=shape)
        nz = o3 != 0
        rel = np.abs(g3[nz].astype(np.float64) - o3[nz]) / np.abs(o3[nz].astype(np.float64))
        print(f"{label:14s} oracle nonzero {nz.mean():.3f}; of those: kernel zero {(g3[nz] == 0).mean():.4f}, within 1e-5 {(rel <= 1e-5).mean():.4f}, "
              f"worst rel {rel.max():.3e}; scaler bytes equal {(gsc == osc).mean():.4f}")


if __name__ == "__main__":
    main()
