"""oracle/tree_oracle.py -- chained newview over a tree on the CPU.  TEST INFRASTRUCTURE ONLY.

The reference has no tree code; a traversal is, by definition, the reference's plf()
(app/src/plf.cpp:8-68, here through the pinned C restatement) applied to every inner node in
post-order, with the per-site scaler bytes of each call summed up the tree."""
from __future__ import annotations

import numpy as np


def traverse(coracle, left, right, tips, ev, p_left, p_right, wgt=None, states: int = 4):
    """tips: [n_tips, n, 4*states].  Returns (root CLV [n, 4*states], per-site counts [n], total scalings).
    states = 4 goes through the pinned DNA restatement, any other count through the general-S loop nest."""
    n_tips = tips.shape[0]
    clv = {i: tips[i] for i in range(n_tips)}
    cnt = {i: np.zeros(tips.shape[1], np.int32) for i in range(n_tips)}
    total = 0
    for k, (a, b) in enumerate(zip(left, right)):
        if states == 4:
            x3, sc, inc = coracle.newview(clv[int(a)], clv[int(b)], ev, p_left[k], p_right[k], wgt)
        else:
            x3, sc, inc = coracle.newview_states(states, clv[int(a)], clv[int(b)], ev, p_left[k], p_right[k], wgt)
        node = n_tips + k
        clv[node] = x3
        cnt[node] = cnt[int(a)] + cnt[int(b)] + sc.astype(np.int32)
        total += inc
        for c in (int(a), int(b)):          # children are dead once their parent exists
            del clv[c], cnt[c]
    root = n_tips + len(left) - 1
    return clv[root], cnt[root], total
