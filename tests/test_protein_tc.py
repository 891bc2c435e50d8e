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


@pytest.mark.gpu
def test_tc_non_finite_inputs_stay_in_their_site_and_category(pkg, coracle):
    """NaN / Inf entries in one (site, category) of a child: the tensor-core kernel's results are non-finite exactly where
    the reference's are (that site's category), every other value stays within tolerance, and such a site is never
    rescaled (a NaN or Inf is not below the threshold).  Rows are TMEM lanes and the operand has no K padding, so nothing
    can leak between sites; this pins it."""
    import torch
    n = 1000
    ev, left, right = matrices(7)
    x1, x2 = pkg.generate_states_host(S, 0, n, 5)
    x1 = x1.reshape(n, SITE).copy()
    x2 = x2.reshape(n, SITE).copy()
    x1[3, 1 * S + 4] = np.nan            # site 3, category 1
    x2[130, 2 * S + 19] = np.inf         # site 130 (second tile), category 2
    x1[255, 0] = -np.inf                 # last row of the second tile, category 0
    x1[4] *= 1e-20                       # a small site next to a poisoned one keeps its rescale
    o3, osc, _ = coracle.newview_states(S, x1, x2, ev, left, right)
    g3, gsc, _ = run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=TC)
    bad_o, bad_g = ~np.isfinite(o3), ~np.isfinite(g3)
    assert np.array_equal(bad_o, bad_g), "non-finite results differ in position from the reference's"
    assert bad_o.sum() == 3 * S and bad_o[3, S:2 * S].all() and bad_o[130, 2 * S:3 * S].all() and bad_o[255, :S].all()
    ok = ~bad_o
    rel = np.abs(g3[ok].astype(np.float64) - o3[ok]) / np.maximum(np.abs(o3[ok].astype(np.float64)), 1e-300)
    assert rel.max() <= REL_TOL
    assert np.array_equal(gsc, osc) and gsc[3] == 0 and gsc[130] == 0 and gsc[255] == 0 and gsc[4] == 1


@pytest.mark.gpu
def test_tc_denormal_operands_are_flushed_not_garbled(pkg, coracle):
    """The tensor core reads fp32 denormals as zero (tools/tc_denormal_probe.py: the CUDA-core FMA kernel keeps them).  A
    CLV entry that small is >= 2^94 below the rescaling threshold of its own site, so the likelihood cannot see it; what
    this pins is that such a value becomes exactly zero -- never garbage -- and that the scaler bytes do not move."""
    import torch
    n = 4096
    ev, left, right = matrices(7)
    x1, x2 = pkg.generate_states_host(S, 0, n, 5)
    x1 = x1.reshape(n, SITE) * np.float32(1e-30)          # 1e-30 ... 1e-42: the designed-small sites drop below FLT_MIN
    x2 = x2.reshape(n, SITE)
    o3, osc, _ = coracle.newview_states(S, x1, x2, ev, left, right)
    g3, gsc, _ = run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=TC)
    rel = np.abs(g3.astype(np.float64) - o3) / np.maximum(np.abs(o3.astype(np.float64)), 1e-300)
    flushed = g3 == 0
    assert ((rel <= REL_TOL) | flushed).all()
    assert flushed.any() and (np.abs(x1[flushed.any(axis=1)]).max(axis=1) < np.finfo(np.float32).tiny).all(), \
        "only sites whose x1 is entirely denormal may come out as zero"
    assert np.array_equal(gsc, osc)


@pytest.mark.gpu
def test_tc_trace_twin_computes_the_same_bits(pkg, coracle, tmp_path, monkeypatch):
    """PLF_TC_TRACE selects the compile-time twin of the kernel that counts wait cycles: same results bit for bit, one
    line of counters per worker warp, and the counters add up (waits <= total, tiles = all tiles)."""
    import torch
    n = 50000
    ev, left, right = matrices(9)
    x1, x2 = pkg.generate_states_host(S, 0, n, 21)
    plain = run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=TC)
    path = tmp_path / "tc_trace.txt"
    monkeypatch.setenv("PLF_TC_TRACE", str(path))
    traced = run_states(pkg, torch, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=TC)
    monkeypatch.delenv("PLF_TC_TRACE")
    assert np.array_equal(plain[0].view(np.uint32), traced[0].view(np.uint32)) and np.array_equal(plain[1], traced[1]) and plain[2] == traced[2]
    rows = np.loadtxt(path, dtype=np.int64, ndmin=2)
    rows = rows[rows[:, 1] < 8]                                 # worker warps (the issuer lanes follow)
    assert rows.shape[1] == 15 and rows.shape[0] % 8 == 0
    assert rows[:, 9].sum() == 4 * ((n + 127) // 128)            # four worker warps per group count the group's tiles
    busy = rows[rows[:, 9] > 0]
    assert (busy[:, 2:8].sum(axis=1) <= busy[:, 8]).all() and (busy[:, 8] > 0).all()
