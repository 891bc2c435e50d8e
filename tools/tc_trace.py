"""Read a PLF_TC_TRACE dump of the tcgen05 20-state kernel (debug build selected by the environment variable): per worker
warp, the cycles spent in each kind of wait and in each segment of a step over the whole launch.

    PLF_TC_TRACE=trace.txt python tools/tc_check.py time ; python tools/tc_trace.py trace.txt

Prints, per (tile, category) step and averaged over all worker warps of all CTAs, where a worker's time goes.
"""
import sys

import numpy as np

WAITS = ["x1 box (TMA)", "x2 box (TMA)", "branch MMAs (mma_ab)", "EV MMAs (mma_x)", "staging rows free", "(unused)"]
# segment -> the waits that happen inside it
SEGMENTS = [("convert x1, x2 of the next step", 8, (0, 1)), ("finish_ab: p = a*b", 9, (2,)), ("results of this category", 10, (3, 4)),
            ("tile end (rescale, stores)", 11, (5,))]


def main():
    allrows = np.loadtxt(sys.argv[1], dtype=np.int64)
    issuers = allrows[allrows[:, 1] >= 10]
    rows = allrows[allrows[:, 1] < 8]
    rows = rows[rows[:, 9] > 0]
    steps = rows[:, 9] * 4
    total = rows[:, 8] / steps
    print(f"{len(rows)} worker warps, {int(rows[:, 9].sum()) // 4} tiles; cycles per step: mean {total.mean():.0f}, "
          f"min {total.min():.0f}, max {total.max():.0f}")
    for k, name in enumerate(WAITS):
        per = rows[:, 2 + k] / steps
        print(f"  wait for {name:22s} {per.mean():7.0f}  ({100 * per.mean() / total.mean():4.1f} %)   min {per.min():6.0f} max {per.max():6.0f}")
    if rows.shape[1] >= 14:
        for name, col, waits in SEGMENTS:
            seg = rows[:, 2 + col] / steps
            inside = sum(rows[:, 2 + w] / steps for w in waits)
            print(f"  segment {name:32s} {seg.mean():7.0f}  of which waiting {inside.mean():6.0f}, busy {seg.mean() - inside.mean():6.0f}")


    if rows.shape[1] >= 15:
        per = rows[:, 14] / steps
        print(f"  hand-off: issuer's commit issued -> worker awake {per.mean():7.0f}   (the tail of b's chain + the mbarrier hop)")
    issuers = issuers[issuers[:, 4] > 0]
    if len(issuers):
        print(f"  hand-off: last worker arrive (x2 operand) -> issuer awake {np.mean(issuers[:, 2] / issuers[:, 4]):7.0f};  "
              f"issue of b's five MMAs + commit {np.mean(issuers[:, 3] / issuers[:, 4]):7.0f}")


if __name__ == "__main__":
    main()
