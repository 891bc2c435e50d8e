"""The likelihood CHAIN -- newview over a tree, per-site scaler counts, evaluate across the root branch -- against an
independent textbook model (oracle/felsenstein_fp64.py: state-space Felsenstein pruning, explicit GTR+Gamma4 transition
matrices, float64 with log-space renormalisation; itself checked against 50-digit mpmath).

Tolerance: |lnL - lnL_textbook| <= 1e-6 * |lnL_textbook|.  The path computes in fp32 (about 1e-7 relative per newview,
random in sign over sites), the textbook in fp64; measured agreement is 2-4e-8."""
from __future__ import annotations

import sys

import numpy as np
import pytest

from oracle import evaluate_oracle, felsenstein_fp64 as F

REL_TOL = 1e-6


def chain_inputs(pkg, n_tips, n, ambiguity, seed, shape="random"):
    model = F.GtrGamma()
    left, right = pkg.random_tree(n_tips, seed=seed) if shape == "random" else pkg.balanced_tree(n_tips)
    tl, tr = F.random_branch_lengths(n_tips - 1, seed + 1)
    codes = F.simulate_alignment(model, left, right, tl, tr, n, seed + 2, ambiguity=ambiguity)
    wgt = np.random.RandomState(seed + 3).randint(1, 5, n).astype(np.int32)
    return model, left, right, tl, tr, codes, wgt


def subtree(coracle, node, n_tips, left, right, tips, ev, pl, pr):
    """CLV and per-site scaler counts of `node` by the pinned newview oracle, post-order over its subtree."""
    n = tips.shape[1]
    if node < n_tips:
        return tips[node], np.zeros(n, np.int32)
    order, stack = [], [node]
    while stack:
        x = stack.pop()
        if x >= n_tips:
            order.append(x - n_tips)
            stack += [int(left[x - n_tips]), int(right[x - n_tips])]
    clv = {i: tips[i] for i in range(n_tips)}
    cnt = {i: np.zeros(n, np.int32) for i in range(n_tips)}
    for k in sorted(order):                          # post-order ids: children always have smaller indices
        a, b = int(left[k]), int(right[k])
        x3, sc, _ = coracle.newview(clv[a], clv[b], ev, pl[k], pr[k], None)
        clv[n_tips + k], cnt[n_tips + k] = x3, cnt[a] + cnt[b] + sc.astype(np.int32)
    return clv[node], cnt[node]


def cpu_chain_lnl(coracle, n_tips, left, right, codes, wgt, ev, pl, pr, tv, diag):
    tips = np.stack([np.tile(tv[codes[i]], (1, 4)) for i in range(n_tips)]).astype(np.float32)
    (xa, ca), (xb, cb) = (subtree(coracle, int(c), n_tips, left, right, tips, ev, pl, pr) for c in (left[-1], right[-1]))
    return evaluate_oracle.evaluate(xa, xb, diag, ca, cb, wgt), int((ca + cb).max())


def test_textbook_model_is_self_consistent():
    m = F.GtrGamma()
    assert abs(m.cat_rates.mean() - 1.0) < 1e-12 and (np.diff(m.cat_rates) > 0).all()
    assert np.allclose(m.u @ np.diag(m.lam) @ m.u_inv, m.q, atol=1e-12)          # the eigen-decomposition is Q
    assert abs(-(m.pi * np.diag(m.q)).sum() - 1.0) < 1e-12                       # one substitution per unit time
    p = m.p_matrix(0.37, 1.3)
    assert np.allclose(p.sum(axis=1), 1.0) and (p > 0).all()
    assert np.allclose(m.pi[:, None] * p, (m.pi[:, None] * p).T)                 # detailed balance


def test_fp64_pruning_matches_mpmath(pkg):
    model, left, right, tl, tr, codes, wgt = chain_inputs(pkg, 7, 10, 0.3, seed=11)
    a = F.log_likelihood(model, left, right, tl, tr, codes, wgt)
    b = F.mp_log_likelihood(model, left, right, tl, tr, codes, wgt)
    assert abs(a - b) <= 1e-12 * abs(b)


@pytest.mark.parametrize("n_tips,n,ambiguity", [(8, 400, 0.05), (64, 1500, 0.05), (64, 600, 1.0), (200, 300, 0.3)])
def test_reference_style_chain_matches_textbook_on_cpu(pkg, coracle, n_tips, n, ambiguity):
    """Pins oracle/evaluate_oracle.py and the EV / P / tip-vector conventions: the pinned newview oracle chained up
    the tree + the evaluate restatement give the textbook log-likelihood."""
    model, left, right, tl, tr, codes, wgt = chain_inputs(pkg, n_tips, n, ambiguity, seed=n_tips)
    want = F.log_likelihood(model, left, right, tl, tr, codes, wgt)
    ev, pl, pr, tv, diag = F.eigen_inputs(model, left, right, tl, tr)
    got, max_cnt = cpu_chain_lnl(coracle, n_tips, left, right, codes, wgt, ev, pl, pr, tv, diag)
    assert abs(got - want) <= REL_TOL * abs(want), (got, want)
    if n_tips >= 64:
        assert max_cnt >= 1, "the tree should be deep enough to rescale"


def test_the_check_has_teeth(pkg, coracle):
    """Each bookkeeping mistake the chain could make moves lnL far outside the tolerance."""
    n_tips, n = 64, 300
    model, left, right, tl, tr, codes, wgt = chain_inputs(pkg, n_tips, n, 0.05, seed=5)
    want = F.log_likelihood(model, left, right, tl, tr, codes, wgt)
    ev, pl, pr, tv, diag = F.eigen_inputs(model, left, right, tl, tr)
    good, max_cnt = cpu_chain_lnl(coracle, n_tips, left, right, codes, wgt, ev, pl, pr, tv, diag)
    assert abs(good - want) <= REL_TOL * abs(want) and max_cnt >= 1
    ev_t = ev.reshape(4, 4).T.reshape(16).copy()                                          # EV used transposed
    pl_t = pl.reshape(-1, 4, 4, 4).transpose(0, 1, 3, 2).reshape(-1, 64).copy()           # P used as [j][l][k]
    for bad in (cpu_chain_lnl(coracle, n_tips, left, right, codes, wgt, ev_t, pl, pr, tv, diag)[0],
                cpu_chain_lnl(coracle, n_tips, left, right, codes, wgt, ev, pl_t, pr, tv, diag)[0],
                cpu_chain_lnl(coracle, n_tips, left, right, codes, wgt, ev, pl, pr, tv, diag[::-1].copy())[0]):
        assert not abs(bad - want) <= 1e-3 * abs(want)
    # dropping the scaler counts (or one log 2^-32 per count) is off by 22.18 per event
    tips = np.stack([np.tile(tv[codes[i]], (1, 4)) for i in range(n_tips)]).astype(np.float32)
    (xa, ca), (xb, cb) = (subtree(coracle, int(c), n_tips, left, right, tips, ev, pl, pr) for c in (left[-1], right[-1]))
    no_counts = evaluate_oracle.evaluate(xa, xb, diag, None, None, wgt)
    assert abs(no_counts - want) > 20.0


@pytest.mark.gpu
@pytest.mark.parametrize("n_tips,n,ambiguity,shape", [(8, 777, 0.05, "random"), (64, 3000, 0.05, "balanced"),
                                                      (64, 1111, 1.0, "random"), (48, 5000, 0.3, "random")])
@pytest.mark.parametrize("tip_codes", [False, True])
def test_cuda_chain_matches_textbook(pkg, n_tips, n, ambiguity, shape, tip_codes):
    """plf_tree_run_async + plf_tree_evaluate_root (CUDA, through the C ABI) against the textbook model."""
    model, left, right, tl, tr, codes, wgt = chain_inputs(pkg, n_tips, n, ambiguity, seed=n_tips + n, shape=shape)
    if tip_codes and (left[-1] < n_tips or right[-1] < n_tips):
        pytest.skip("a child of the root is a compressed tip")
    want = F.log_likelihood(model, left, right, tl, tr, codes, wgt)
    ev, pl, pr, tv, diag = F.eigen_inputs(model, left, right, tl, tr)
    with pkg.Tree(left, right, n, tip_codes=tip_codes) as t:
        if tip_codes:
            t.write_tip_vector(tv)
        for i in range(n_tips):
            if tip_codes:
                t.write_tip_codes(i, codes[i])
            else:
                t.write_tip(i, np.tile(tv[codes[i]], (1, 4)).astype(np.float32))
        t.write_matrices(ev, pl, pr)
        t.write_wgt(wgt)
        t.run_async()
        lnl = [t.evaluate_root(diag) for _ in range(3)]
        _, cnt = t.read_root()
    assert lnl[0] == lnl[1] == lnl[2], "the device reduction must be reproducible bit for bit"
    assert abs(lnl[0] - want) <= REL_TOL * abs(want), (lnl[0], want)
    if n_tips >= 48:
        assert cnt.max() >= 1, "the tree should be deep enough to rescale"
