"""oracle/plf_numpy.py -- numpy float32 restatement of the reference PLF newview step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  A third, independent statement of
app/src/plf.cpp:19-65 (/root/reference) used to cross-check the C restatement: every
multiply and every add is a separate float32 numpy ufunc call, so each result is rounded
to fp32 exactly once per operation and in the reference's order (no FMA, no pairwise sums).
"""
from __future__ import annotations

import numpy as np

F = np.float32
MINLIKELIHOOD = F(2.0 ** -32)   # plf.cpp:5-6
TWOTOTHE32 = F(2.0 ** 32)       # plf.cpp:5


def newview(x1, x2, ev, left, right, wgt=None):
    """Returns (x3[n,16] f32, scaler[n] u8, scaler_increment int).

    x1,x2: [n,16] site-major [site][cat j][state l]; ev: [16] = [k][l] (or [64] = per-category
    [j][k][l], the INPUT_SRC=gen analogue); left,right: [64] = [j][k][l]."""
    x1 = np.ascontiguousarray(x1, dtype=F).reshape(-1, 4, 4)
    x2 = np.ascontiguousarray(x2, dtype=F).reshape(-1, 4, 4)
    n = x1.shape[0]
    ev = np.asarray(ev, dtype=F)
    ev = np.broadcast_to(ev.reshape(1, 4, 4), (4, 4, 4)) if ev.size == 16 else ev.reshape(4, 4, 4)
    left = np.asarray(left, dtype=F).reshape(4, 4, 4)
    right = np.asarray(right, dtype=F).reshape(4, 4, 4)
    x3 = np.empty((n, 4, 4), dtype=F)
    zero = np.zeros(n, dtype=F)
    with np.errstate(all="ignore"):
        for j in range(4):
            p = []
            for k in range(4):
                a = zero
                b = zero
                for l in range(4):                           # plf.cpp:35-39
                    a = a + x1[:, j, l] * left[j, k, l]
                    b = b + x2[:, j, l] * right[j, k, l]
                p.append(a * b)                              # plf.cpp:41
            for l in range(4):                               # plf.cpp:45-50
                acc = zero
                for k in range(4):
                    acc = acc + p[k] * ev[j, k, l]
                x3[:, j, l] = acc
        x3 = x3.reshape(n, 16)
        mag = np.where(x3 < 0, -x3, x3)                      # ABS(), plf.cpp:4
        scaled = np.all(mag < MINLIKELIHOOD, axis=1)         # plf.cpp:53-56 (NaN -> False)
        x3[scaled] = x3[scaled] * TWOTOTHE32                 # plf.cpp:58-61
    sc = scaled.astype(np.uint8)
    if wgt is None:
        inc = int(sc.sum())
    else:
        inc = int((sc.astype(np.int64) * np.asarray(wgt, dtype=np.int64)).sum())  # plf.cpp:63
    return x3, sc, inc
