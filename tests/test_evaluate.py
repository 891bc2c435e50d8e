"""Root log-likelihood across a branch.  The reference has no such step; the CUDA kernel is
checked here against oracle/evaluate_oracle.py (RAxML's evaluateGTRGAMMA restated), fp64 on both sides, tolerance
1e-9 relative, and that restatement is pinned to an independent textbook model in tests/test_felsenstein.py."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import evaluate_oracle, tree_oracle
from test_tree import tree_inputs

REL_TOL = 1e-9


def test_evaluate_oracle_hand_case():
    x1 = np.full((2, 16), 0.5, np.float32)
    x2 = np.full((2, 16), 0.25, np.float32)
    diag = np.ones(16, np.float32)
    # per site: 16 * 0.125 = 2 -> log(0.5); second site rescaled 3 times in total
    got = evaluate_oracle.evaluate(x1, x2, diag, cnt1=[0, 1], cnt2=[0, 2], wgt=[2, 1])
    want = 2 * np.log(0.5) + (np.log(0.5) + 3 * (-32 * np.log(2)))
    assert abs(got - want) < 1e-12


@pytest.mark.gpu
# from 2^18 sites on, 4-state calls take the ring-fed kernel (stages of 768 sites: all of these end in a ragged stage)
@pytest.mark.parametrize("n", [1, 7, 8, 1000, 65537, (1 << 18) - 1, 1 << 18, 300001, 1 << 20])
def test_evaluate_device_matches_oracle(pkg, n):
    import torch
    rng = np.random.RandomState(n)
    x1 = (rng.random_sample((n, 16)) * 10.0 ** rng.uniform(-9, 0, (n, 1))).astype(np.float32)
    x2 = rng.random_sample((n, 16)).astype(np.float32)
    diag = rng.random_sample(16).astype(np.float32)
    c1 = rng.randint(0, 4, n).astype(np.int32)
    c2 = rng.randint(0, 4, n).astype(np.int32)
    wgt = rng.randint(1, 9, n).astype(np.int32)
    d = [torch.from_numpy(a).cuda() for a in (x1, x2, c1, c2, wgt, diag)]
    for use_cnt, use_w in ((True, True), (False, False), (True, False)):
        lnl = torch.zeros(1, dtype=torch.float64, device="cuda")
        pkg.evaluate_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr() if use_cnt else None,
                            d[3].data_ptr() if use_cnt else None, d[4].data_ptr() if use_w else None,
                            d[5].data_ptr(), n, lnl.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = evaluate_oracle.evaluate(x1, x2, diag, c1 if use_cnt else None, c2 if use_cnt else None,
                                        wgt if use_w else None)
        assert abs(lnl.item() - want) <= REL_TOL * abs(want), (n, use_cnt, use_w, lnl.item(), want)
        again = torch.zeros(1, dtype=torch.float64, device="cuda")           # same launch, same bits
        pkg.evaluate_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr() if use_cnt else None,
                            d[3].data_ptr() if use_cnt else None, d[4].data_ptr() if use_w else None,
                            d[5].data_ptr(), n, again.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert again.item() == lnl.item()


@pytest.mark.gpu
def test_tree_root_log_likelihood(pkg, coracle):
    """Traversal + evaluation across the root branch == oracle traversal of the two root subtrees
    + oracle evaluation; also invariant under re-running the traversal."""
    n_tips, n = 40, 3000
    left, right = pkg.random_tree(n_tips, seed=3)
    tips, ev, pl, pr, wgt = tree_inputs(n_tips, n, seed=8)
    diag = np.random.RandomState(2).random_sample(16).astype(np.float32)
    # oracle: CLVs and counts of the root's two children
    sub = {}
    for child in (int(left[-1]), int(right[-1])):
        if child < n_tips:
            sub[child] = (tips[child], np.zeros(n, np.int32))
        else:
            k = child - n_tips
            # post-order prefix up to inner node k is a forest that contains k's whole subtree
            clv = {i: tips[i] for i in range(n_tips)}
            cnt = {i: np.zeros(n, np.int32) for i in range(n_tips)}
            for q in range(k + 1):
                x3, sc, _ = coracle.newview(clv[int(left[q])], clv[int(right[q])], ev, pl[q], pr[q], wgt)
                clv[n_tips + q] = x3
                cnt[n_tips + q] = cnt[int(left[q])] + cnt[int(right[q])] + sc.astype(np.int32)
            sub[child] = (clv[child], cnt[child])
    (xa, ca), (xb, cb) = sub[int(left[-1])], sub[int(right[-1])]
    want = evaluate_oracle.evaluate(xa, xb, diag, ca, cb, wgt)
    with pkg.Tree(left, right, n) as t:
        for i in range(n_tips):
            t.write_tip(i, tips[i])
        t.write_matrices(ev, pl, pr)
        t.write_wgt(wgt)
        with pytest.raises(pkg.PlfError):
            t.evaluate_root(diag)                 # not traversed yet
        for _ in range(2):
            t.run_async()
            got = t.evaluate_root(diag)
            assert abs(got - want) <= REL_TOL * abs(want), (got, want)
    assert np.isfinite(want) and want < 0
