#!/usr/bin/env python
"""tools/stream_timeline.py -- the overlap of plf_newview_stream's pipeline from a PLF_STREAM_TRACE dump (developer tool;
the stand-in for an nsys timeline, which this image does not have).

    PLF_STREAM_TRACE=trace.txt NO_CORRECTNESS_CHECK=1 host_stream.exe <config> 0 16777216 2 ; python tools/stream_timeline.py trace.txt

Per chunk the library records four events on the chunk's stream: before its H2D copies, after them, after the kernel,
after its D2H copies.  The kernel interval is an upper bound (it opens when the copies end, the kernel may start later).
"""
import sys

import numpy as np


def union(iv):
    iv = sorted(iv)
    out, total = [], 0.0
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return sum(b - a for a, b in out)


def main():
    path = sys.argv[1]
    head = open(path).readline().strip()
    rows = np.loadtxt(path, comments="#", ndmin=2)
    t = rows[:, 3:7]
    sites = rows[:, 2]
    end = t[:, 3].max()
    grid = np.linspace(0.0, end, 20001)
    mid = 0.5 * (grid[1:] + grid[:-1])
    act = np.zeros((3, mid.size), dtype=bool)
    for r in t:
        for j in range(3):
            act[j] |= (mid >= r[j]) & (mid < r[j + 1])
    dt = end / mid.size
    n_act = act.sum(axis=0)
    print(head)
    print(f"chunks {len(rows)}, sites {int(sites.sum())}, wall {end:.3f} ms")
    for j, name in enumerate(("H2D copies", "kernel (upper bound)", "D2H copies")):
        print(f"  {name:22s} active {act[j].sum() * dt:8.3f} ms = {100 * act[j].mean():5.1f} % of the call;  "
              f"sum of intervals {np.sum(t[:, j + 1] - t[:, j]):8.3f} ms")
    for k in (1, 2, 3):
        print(f"  >= {k} of the three active   {(n_act >= k).sum() * dt:8.3f} ms = {100 * (n_act >= k).mean():5.1f} %")
    print(f"  H2D and D2H at once        {(act[0] & act[2]).sum() * dt:8.3f} ms = {100 * (act[0] & act[2]).mean():5.1f} %")
    # a character timeline of the first chunks: H = H2D, K = kernel, D = D2H
    width = 100
    show = min(len(rows), 12)
    span = t[show - 1, 3]
    print(f"\nfirst {show} chunks, {span:.2f} ms across {width} columns (H = H2D, K = kernel, D = D2H):")
    for i in range(show):
        line = [" "] * width
        for j, ch in enumerate("HKD"):
            a, b = int(t[i, j] / span * width), int(np.ceil(t[i, j + 1] / span * width))
            for c in range(a, min(max(b, a + 1), width)):
                line[c] = ch
        print(f"  chunk {i:2d} slot {int(rows[i, 1])} |{''.join(line)}|")


if __name__ == "__main__":
    main()
