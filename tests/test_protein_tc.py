"""The tensor-core 20-state kernel (csrc/plf_protein_tc.cu: tcgen05.mma kind::tf32 with the 3xTF32 split, accumulators in
tensor memory, operands by TMA) against the CPU restatement of the reference's loop nest with 20 states.

Tolerance mode only: tensor cores cannot reproduce the reference's rounding sequence.  Bound: 1e-5 relative on positive
data (the designed stimulus; measured 2.6e-6 worst, 1.7e-6 median), and for signed data 1e-5 of the condition-free
magnitude |x1||P_l| . |x2||P_r| . |EV| (cancellation makes a relative bound meaningless there).  Scaler bytes are compared
exactly away from the 2^-32 threshold."""
from __future__ import annotations

import numpy as np
import pytest

from test_protein import S, SITE, matrices, run_states

TC = (9, 0)              # (variant, threads): the tcgen05 kernel
REL_TOL = 1e-5


def magnitude_bound(x1, x2, ev, left, right):
    """The newview evaluated on absolute values in float64: the scale against which absolute errors are judged."""
    n = x1.shape[0]
    a1 = np.abs(x1.astype(np.float64)).reshape(n, 4, S)
    a2 = np.abs(x2.astype(np.float64)).reshape(n, 4, S)
    pl = np.abs(left.astype(np.float64)).reshape(4, S, S)
    pr = np.abs(right.astype(np.float64)).reshape(4, S, S)
    e = np.abs(ev.astype(np.float64)).reshape(S, S)
    a = np.einsum("njl,jkl->njk", a1, pl)
    b = np.einsum("njl,jkl->njk", a2, pr)
    return np.einsum("njk,kl->njl", a * b, e).reshape(n, SITE)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 127, 128, 129, 255, 257, 1000, 4099, 65536 + 5, 300001])
def test_tc_designed_stimulus_within_tolerance(pkg, coracle, n):
    import torch
    ev, left, right = matrices(3)
    x1, x2 = pkg.generate_states_host(S, 1000, n, 42)
    wgt = np.random.RandomState(n).randint(1, 7, n).astype(np.int32)
    o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right, wgt)
    g3, gsc, ginc = run_states(pkg, torch, S, ev, left, right, x1, x2, wgt=wgt, math=pkg.MATH_FMA, shape=TC)
    rel = np.abs(g3.astype(np.float64) - o3) / np.maximum(np.abs(o3.astype(np.float64)), 1e-300)
    assert rel.max() <= REL_TOL, f"max relative error {rel.max():.3e}"
    assert np.array_equal(gsc, osc) and ginc == oinc      # designed stimulus: no site near the threshold


@pytest.mark.gpu
def test_tc_signed_wide_range_normwise(pkg, coracle):
    import torch
    n = 20011
    rng = np.random.RandomState(17)
    ev, left, right = matrices(18, signed=True)
    x1 = (rng.standard_normal((n, SITE)) * 10.0 ** rng.uniform(-14, 1, (n, 1))).astype(np.float32)
    x2 = (rng.standard_normal((n, SITE)) * 10.0 ** rng.uniform(-3, 1, (n, 1))).astype(np.float32)
    x1[5] = -0.0
    x1[12] = 0.0
    o3, osc, _ = coracle.newview_states(S, x1, x2, ev, left, right)
    g3, gsc, _ = run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=TC)
    # scaler decisions may differ only for sites whose largest entry is within the tolerance of 2^-32
    unscaled = np.abs(o3).max(axis=1) / np.where(osc == 1, 2.0 ** 32, 1.0)
    near = np.abs(unscaled / 2.0 ** -32 - 1.0) < 1e-3
    differ = gsc != osc
    assert not (differ & ~near).any(), f"{int((differ & ~near).sum())} scaler bytes differ away from the threshold"
    # values: against the condition-free magnitude (times 2^32 where the site was rescaled), on sites with equal decisions
    bound = magnitude_bound(x1, x2, ev, left, right) * np.where(osc == 1, 2.0 ** 32, 1.0)[:, None]
    err = np.abs(g3.astype(np.float64) - o3.astype(np.float64))[~differ]
    assert (err <= REL_TOL * bound[~differ] + 1e-45).all(), \
        f"worst error / magnitude = {(err / np.maximum(bound[~differ], 1e-300)).max():.3e}"


@pytest.mark.gpu
def test_tc_is_the_fma_default_for_long_calls_and_refuses_strict(pkg, coracle):
    import torch
    info_tc = pkg.states_kernel_info(20, pkg.MATH_FMA, 9, 0)
    assert info_tc["threads"] == 384 and info_tc["tile_sites"] == 128 and info_tc["smem_bytes"] > 170 * 1024
    ev, left, right = matrices(5)
    n = 40000
    x1, x2 = pkg.generate_states_host(S, 0, n, 3)
    before = pkg.launch_count()
    d = run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=(0, 0))          # default shape, FMA
    t = run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=TC)
    assert np.array_equal(d[0].view(np.uint32), t[0].view(np.uint32)), "FMA default for a long call should be the tensor-core kernel"
    with pytest.raises(pkg.PlfError):
        run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_STRICT, shape=TC)         # no strict mode on tensor cores
    assert pkg.launch_count() > before


@pytest.mark.gpu
def test_tc_tree_with_scaler_counts(pkg, coracle):
    """FMA-mode 20-state tree (long enough per node to take the tensor-core kernel): counts and root within tolerance."""
    from oracle import tree_oracle
    n_tips, n = 9, 20000
    left, right = pkg.random_tree(n_tips, seed=4)
    rng = np.random.RandomState(9)
    tips = (rng.random_sample((n_tips, n, SITE)) * 10.0 ** rng.uniform(-12, 0, (n_tips, n, 1))).astype(np.float32)
    ev = (rng.random_sample(S * S) * 0.2).astype(np.float32)
    pl = (rng.random_sample((n_tips - 1, 4 * S * S)) * 0.2).astype(np.float32)
    pr = (rng.random_sample((n_tips - 1, 4 * S * S)) * 0.2).astype(np.float32)
    o_root, o_cnt, _ = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, None, states=S)
    with pkg.Tree(left, right, n, states=S) as t:
        t.set_math(pkg.MATH_FMA)
        for i in range(n_tips):
            t.write_tip(i, tips[i])
        t.write_matrices(ev, pl, pr)
        t.run_async()
        root, cnt = t.read_root()
    same = cnt == o_cnt
    assert same.mean() > 0.999
    rel = np.abs(root[same].astype(np.float64) - o_root[same]) / np.maximum(np.abs(o_root[same]), 1e-300)
    assert rel.max() <= 8 * REL_TOL          # error compounds over the depth of the tree
