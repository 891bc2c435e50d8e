"""The STATES knob as a first-class citizen (reference Makefile:31, README.md:36,202): the instance API, the streamed
host path, the multi-GPU wrapper, the tree executor and evaluate, each run for S = 4 (DNA) and S = 20 (protein) against
the same checker -- the reference's loop nest with the state count as a parameter (pinned at S = 4; at S = 20 the
reference has no path, golden vector or test, so parity there is against the restatement only: see DESIGN.md)."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import bits, first_mismatch
from oracle import evaluate_oracle, tree_oracle


def inputs(pkg, S, n, seed):
    rng = np.random.RandomState(seed)
    ev = rng.random_sample(S * S).astype(np.float32)
    left = rng.random_sample(4 * S * S).astype(np.float32)
    right = rng.random_sample(4 * S * S).astype(np.float32)
    x1, x2 = pkg.generate_states_host(S, 0, n, seed)           # every 4th site rescales
    wgt = rng.randint(1, 5, n).astype(np.int32)
    return ev, left, right, x1, x2, wgt


def test_testbench_info_scales_with_the_state_count(pkg):
    for S, hl, hs in ((4, 80, 64), (20, 2000, 1600)):
        tb = pkg.TestbenchInfo(1000, 3, elements_per_alignment=4 * S)
        assert tb.states == S and tb.header_left() == hl
        assert pkg.TestbenchInfo(1000, 3, layout=pkg.LAYOUT_SEP, elements_per_alignment=4 * S).header_right() == hs
        assert tb.instance_active_elements_left(0) == 334 * 4 * S + hl
    ev, pl, x = np.zeros(400, np.float32), np.ones(1600, np.float32), np.full((3, 80), 2, np.float32)
    assert pkg.pack_left(ev, pl, x).shape == (2240,) and pkg.pack_right(ev, pl, x, pkg.LAYOUT_SEP).shape == (1840,)


def test_states_contexts_fail_loudly_without_gpu_or_with_bad_knobs(pkg):
    import ctypes
    lib, c = pkg.load(), ctypes.c_void_p()
    assert lib.plf_ctx_create_states(ctypes.byref(c), 0, 1, 0, 0, 5) == -1 and b"STATES=5" in lib.plf_last_error(None)
    assert lib.plf_ctx_create_states(ctypes.byref(c), 0, 1, 0, pkg.INPUT_GEN, 20) == -1       # gen movers are DNA only
    t = ctypes.c_void_p()
    l, r = (ctypes.c_int * 1)(0), (ctypes.c_int * 1)(1)
    assert lib.plf_tree_create_states(ctypes.byref(t), 0, 2, l, r, 10, 1, 20) == -1               # code tips are DNA only


@pytest.mark.gpu
@pytest.mark.parametrize("S", [4, 20])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n,instances", [(1, 1), (100, 1), (4099, 3), (30011, 9)])
def test_instance_api_for_both_state_counts(pkg, coracle, S, layout, n, instances):
    ev, left, right, x1, x2, wgt = inputs(pkg, S, n, seed=S + n)
    o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right, wgt)
    with pkg.Context(0, 9, layout, pkg.INPUT_MEM, states=S) as ctx:
        for _ in range(2):                         # second call: buffers are reused, matrices re-read from the head
            x3, sc, inc = ctx.newview(ev, left, right, x1, x2, wgt, instances=instances)
            assert np.array_equal(bits(x3), bits(o3)), first_mismatch(x3, o3)
            assert np.array_equal(sc, osc) and inc == oinc


@pytest.mark.gpu
@pytest.mark.parametrize("S", [4, 20])
def test_streamed_path_and_multi_for_both_state_counts(pkg, coracle, S):
    n = 70001
    ev, left, right, x1, x2, wgt = inputs(pkg, S, n, seed=S)
    o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right, wgt)
    with pkg.Context(0, 1, states=S) as ctx:
        x3 = np.empty_like(x1)
        sc = np.empty(n, np.uint8)
        inc = ctx.newview_stream(ev, left, right, x1, x2, x3, sc, wgt, chunk_sites=8192)
        assert np.array_equal(bits(x3), bits(o3)), first_mismatch(x3, o3)
        assert np.array_equal(sc, osc) and inc == oinc
    with pkg.Multi([0], states=S) as m:
        x3, sc, inc = m.newview(ev, left, right, x1, x2, wgt)
        assert np.array_equal(bits(x3), bits(o3)) and np.array_equal(sc, osc) and inc == oinc


@pytest.mark.gpu
@pytest.mark.parametrize("S", [4, 20])
@pytest.mark.parametrize("shape,n_tips,n", [("balanced", 16, 1001), ("random", 33, 517), ("random", 64, 3000)])
def test_tree_and_root_likelihood_for_both_state_counts(pkg, coracle, S, shape, n_tips, n):
    left, right = pkg.balanced_tree(n_tips) if shape == "balanced" else pkg.random_tree(n_tips, seed=n)
    rng = np.random.RandomState(n_tips * S)
    sf = 4 * S
    tips = (rng.random_sample((n_tips, n, sf)) * 10.0 ** rng.uniform(-12, 0, (n_tips, n, 1))).astype(np.float32)
    ev = (rng.random_sample(S * S) * (4.0 / S)).astype(np.float32)          # keep the 20-term sums near the 4-term ones
    pl = (rng.random_sample((n_tips - 1, 4 * S * S)) * (4.0 / S)).astype(np.float32)
    pr = (rng.random_sample((n_tips - 1, 4 * S * S)) * (4.0 / S)).astype(np.float32)
    wgt = rng.randint(1, 6, n).astype(np.int32)
    diag = rng.random_sample(sf).astype(np.float32)
    o_root, o_cnt, o_total = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, wgt, states=S)
    assert o_cnt.max() >= 1 and np.isfinite(o_root).all()
    # children of the root for the evaluate check
    sub = {}
    for child in (int(left[-1]), int(right[-1])):
        if child < n_tips:
            sub[child] = (tips[child], np.zeros(n, np.int32))
        else:
            k = child - n_tips                                   # post-order: nodes 0..k are a closed sub-forest
            # recompute the subtree by truncating the post-order list at node k
            r_, c_, _ = _prefix_traverse(coracle, left, right, tips, ev, pl, pr, k, S)
            sub[child] = (r_, c_)
    (xa, ca), (xb, cb) = sub[int(left[-1])], sub[int(right[-1])]
    want_lnl = evaluate_oracle.evaluate(xa, xb, diag, ca, cb, wgt)
    with pkg.Tree(left, right, n, states=S) as t:
        for i in range(n_tips):
            t.write_tip(i, tips[i])
        t.write_matrices(ev, pl, pr)
        t.write_wgt(wgt)
        for _ in range(2):
            t.run_async()
            root, cnt = t.read_root()
            assert np.array_equal(bits(root), bits(o_root)), first_mismatch(root, o_root)
            assert np.array_equal(cnt, o_cnt) and t.total_scalings() == o_total
        lnl = [t.evaluate_root(diag) for _ in range(2)]
        assert lnl[0] == lnl[1] and abs(lnl[0] - want_lnl) <= 1e-9 * abs(want_lnl), (lnl, want_lnl)


def _prefix_traverse(coracle, left, right, tips, ev, pl, pr, k, S):
    """CLV and counts of inner node k: run the post-order list up to k, keeping every node alive."""
    n_tips, n = tips.shape[0], tips.shape[1]
    clv = {i: tips[i] for i in range(n_tips)}
    cnt = {i: np.zeros(n, np.int32) for i in range(n_tips)}
    for j in range(k + 1):
        a, b = int(left[j]), int(right[j])
        if S == 4:
            x3, sc, _ = coracle.newview(clv[a], clv[b], ev, pl[j], pr[j], None)
        else:
            x3, sc, _ = coracle.newview_states(S, clv[a], clv[b], ev, pl[j], pr[j], None)
        clv[n_tips + j], cnt[n_tips + j] = x3, cnt[a] + cnt[b] + sc.astype(np.int32)
    return clv[n_tips + k], cnt[n_tips + k], 0
