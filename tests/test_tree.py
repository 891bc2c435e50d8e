"""Chained newview over a tree (BASELINE.json configs[4]): the level-batched CUDA traversal, through
the C ABI, against the oracle applied node by node in post-order.  Strict math: bit-exact CLVs and
identical per-site scaler counts."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import bits, first_mismatch
from oracle import tree_oracle


def tree_inputs(n_tips, n, seed, scale_lo=-14.0):
    """Random positive tip CLVs whose magnitude varies per (tip, site) so that rescaling events
    happen at different depths of the tree; P/EV positive as in the reference stimulus."""
    rng = np.random.RandomState(seed)
    tips = (rng.random_sample((n_tips, n, 16)) * 10.0 ** rng.uniform(scale_lo, 0, (n_tips, n, 1))).astype(np.float32)
    ev = rng.random_sample(16).astype(np.float32)
    pl = rng.random_sample((n_tips - 1, 64)).astype(np.float32)
    pr = rng.random_sample((n_tips - 1, 64)).astype(np.float32)
    wgt = rng.randint(1, 6, n).astype(np.int32)
    return tips, ev, pl, pr, wgt


def test_tree_builders_are_postorder(pkg):
    for n_tips in (2, 3, 7, 64, 1024):
        for left, right in (pkg.balanced_tree(n_tips), pkg.random_tree(n_tips, 1)):
            assert left.size == right.size == n_tips - 1
            seen = set()
            for k, (a, b) in enumerate(zip(left, right)):
                assert a < n_tips + k and b < n_tips + k and a != b
                assert a not in seen and b not in seen
                seen.update((int(a), int(b)))
            assert len(seen) == 2 * (n_tips - 1)      # everything but the root is somebody's child


def test_tree_oracle_counts_are_sums_of_bytes(coracle):
    tips, ev, pl, pr, wgt = tree_inputs(8, 50, 0)
    left, right = np.array([0, 2, 8, 4, 6, 11, 10], np.int32), np.array([1, 3, 9, 5, 7, 12, 13], np.int32)
    root, cnt, total = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, wgt)
    assert root.shape == (50, 16) and cnt.min() >= 0 and total >= int((cnt * wgt).sum()) - 0
    assert total == int((cnt.astype(np.int64) * wgt).sum())      # every event reaches the root count


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n_tips,n", [("balanced", 64, 3001), ("random", 33, 1000), ("balanced", 2, 129),
                                             ("random", 257, 517), ("caterpillar", 12, 4096)])
@pytest.mark.parametrize("u,chunk", [(0, 0), (1, 1), (2, 1), (2, 3), (1, 1000), (3, 0), (3, 7), (4, 0), (4, 2), (0, 5)])
def test_tree_traversal_matches_oracle(pkg, coracle, shape, n_tips, n, u, chunk):
    if shape == "balanced":
        left, right = pkg.balanced_tree(n_tips)
    elif shape == "random":
        left, right = pkg.random_tree(n_tips, seed=n)
    else:       # caterpillar: every level has one node -> n_tips-1 launches
        left = np.array([0] + [n_tips + k for k in range(n_tips - 2)], np.int32)
        right = np.arange(1, n_tips, dtype=np.int32)
    tips, ev, pl, pr, wgt = tree_inputs(n_tips, n, seed=n_tips)
    o_root, o_cnt, o_total = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, wgt)
    assert o_cnt.max() >= min(2, n_tips - 1), "stimulus should rescale repeatedly along the tree"
    with pkg.Tree(left, right, n) as t:
        t.set_tuning(u, chunk)
        for i in range(n_tips):
            t.write_tip(i, tips[i])
        t.write_matrices(ev, pl, pr)
        t.write_wgt(wgt)
        for _ in range(2):          # replay the captured graph: buffers are recycled correctly
            t.run_async()
            root, cnt = t.read_root()
            assert np.array_equal(bits(root), bits(o_root)), first_mismatch(root, o_root)
            assert np.array_equal(cnt, o_cnt)
            assert t.total_scalings() == o_total
        info = t.info()
        assert info["levels"] >= int(np.ceil(np.log2(n_tips))) and info["clv_slots"] <= n_tips - 1
        if shape == "balanced" and n_tips == 64:
            assert info["levels"] == 6 and info["clv_slots"] <= 48      # recycling: 32 + 16 live at most
        assert t.last_ms() > 0


@pytest.mark.gpu
def test_tree_fma_mode_within_tolerance(pkg, coracle):
    left, right = pkg.balanced_tree(16)
    tips, ev, pl, pr, wgt = tree_inputs(16, 2000, seed=5, scale_lo=-3.0)
    o_root, o_cnt, _ = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, None)
    with pkg.Tree(left, right, 2000) as t:
        t.set_math(pkg.MATH_FMA)
        for i in range(16):
            t.write_tip(i, tips[i])
        t.write_matrices(ev, pl, pr)
        t.run_async()
        root, cnt = t.read_root()
    same = cnt == o_cnt
    assert same.mean() > 0.999          # threshold-boundary sites may differ by one rescale
    rel = np.abs(root[same].astype(np.float64) - o_root[same]) / np.maximum(np.abs(o_root[same]), 1e-300)
    assert rel.max() <= 4 * 1e-5        # error compounds over log2(16) = 4 chained newviews


@pytest.mark.gpu
def test_tree_error_behaviour(pkg):
    with pytest.raises(pkg.PlfError):
        pkg.Tree(np.array([0, 1], np.int32), np.array([1, 3], np.int32), 10)      # node 1 used twice
    with pytest.raises(pkg.PlfError):
        pkg.Tree(np.array([0], np.int32), np.array([5], np.int32), 10)            # child id not earlier
    left, right = pkg.balanced_tree(4)
    with pkg.Tree(left, right, 10) as t:
        with pytest.raises(pkg.PlfError):
            t.read_root()                                                         # not run yet
        with pytest.raises(pkg.PlfError):
            t.write_tip(4, np.zeros((10, 16), np.float32))                        # tip out of range
        with pytest.raises(pkg.PlfError):
            t.write_tip(0, np.zeros((11, 16), np.float32))                        # too large


@pytest.mark.gpu
def test_tree_stress_single_busy_cta_with_poisoned_memory(pkg, coracle):
    """Regression for the ring's write-after-read hazard (profiles/r01_tree.md): one CTA does a whole
    level (chunk >= level size), refills come from L2 and are fast, device memory is poisoned with
    NaNs between iterations.  Before the proxy fence in the slot release ~1 in 3 iterations failed."""
    import torch
    cases = []
    for shape, n_tips, n in (("random", 257, 517), ("balanced", 64, 3001)):
        left, right = pkg.random_tree(n_tips, seed=n) if shape == "random" else pkg.balanced_tree(n_tips)
        tips, ev, pl, pr, wgt = tree_inputs(n_tips, n, seed=n_tips)
        cases.append((left, right, n_tips, n, tips, ev, pl, pr, wgt,
                      tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, wgt)))
    for it in range(12):
        poison = torch.full((32 << 20,), float("nan"), device="cuda")
        del poison
        torch.cuda.empty_cache()
        for left, right, n_tips, n, tips, ev, pl, pr, wgt, (o_root, o_cnt, o_total) in cases:
            for u in (1, 2, 3, 4, 0):
                with pkg.Tree(left, right, n) as t:
                    t.set_tuning(u, 1000)
                    for i in range(n_tips):
                        t.write_tip(i, tips[i])
                    t.write_matrices(ev, pl, pr)
                    t.write_wgt(wgt)
                    for _ in range(2):
                        t.run_async()
                        root, cnt = t.read_root()
                        assert np.array_equal(bits(root), bits(o_root)), (it, u, first_mismatch(root, o_root))
                        assert np.array_equal(cnt, o_cnt) and t.total_scalings() == o_total


def expand_tip_codes(codes, tip_vector):
    """Dense CLV of a compressed tip: x[i][j][l] = tip_vector[code_i][l] for every category j."""
    tv = np.asarray(tip_vector, np.float32).reshape(16, 4)
    return np.tile(tv[codes], (1, 4)).astype(np.float32)          # [n, 16]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n_tips,n", [("balanced", 64, 3001), ("random", 33, 1000), ("balanced", 2, 129),
                                             ("random", 257, 517), ("caterpillar", 12, 4099)])
@pytest.mark.parametrize("u,chunk", [(0, 0), (1, 1), (2, 3), (1, 1000), (3, 5), (4, 1)])
def test_tree_with_compressed_tips_matches_dense_oracle(pkg, coracle, shape, n_tips, n, u, chunk):
    """tips stored as one state code per site + a 16x4 tip-vector table give exactly the
    traversal of the expanded dense tips (tip-tip, tip-inner and inner-inner nodes all occur)."""
    if shape == "balanced":
        left, right = pkg.balanced_tree(n_tips)
    elif shape == "random":
        left, right = pkg.random_tree(n_tips, seed=n)
    else:
        left = np.array([0] + [n_tips + k for k in range(n_tips - 2)], np.int32)
        right = np.arange(1, n_tips, dtype=np.int32)
    rng = np.random.RandomState(n_tips + n)
    codes = rng.randint(0, 16, (n_tips, n)).astype(np.uint8)
    tip_vector = (rng.random_sample((16, 4)) * 10.0 ** rng.uniform(-9, 0, (16, 1))).astype(np.float32)
    _, ev, pl, pr, wgt = tree_inputs(n_tips, 4, seed=n_tips)
    wgt = rng.randint(1, 6, n).astype(np.int32)
    dense = np.stack([expand_tip_codes(codes[i], tip_vector) for i in range(n_tips)])
    o_root, o_cnt, o_total = tree_oracle.traverse(coracle, left, right, dense, ev, pl, pr, wgt)
    with pkg.Tree(left, right, n, tip_codes=True) as t:
        t.set_tuning(u, chunk)
        for i in range(n_tips):
            t.write_tip_codes(i, codes[i])
        t.write_tip_vector(tip_vector)
        t.write_matrices(ev, pl, pr)
        t.write_wgt(wgt)
        with pytest.raises(pkg.PlfError):
            t.write_tip(0, dense[0])                      # wrong tip format for this tree
        for _ in range(2):
            t.run_async()
            root, cnt = t.read_root()
            assert np.array_equal(bits(root), bits(o_root)), first_mismatch(root, o_root)
            assert np.array_equal(cnt, o_cnt) and t.total_scalings() == o_total
        dense_bytes = n * 64 * n_tips
        assert t.info()["device_bytes"] < dense_bytes + n * 68 * (n_tips - 1)     # tips cost 1 B/site, not 64
    assert o_cnt.max() >= 1


@pytest.mark.gpu
def test_dense_tree_rejects_tip_codes(pkg):
    left, right = pkg.balanced_tree(4)
    with pkg.Tree(left, right, 10) as t:
        with pytest.raises(pkg.PlfError):
            t.write_tip_codes(0, np.zeros(10, np.uint8))
        with pytest.raises(pkg.PlfError):
            t.write_tip_vector(np.zeros(64, np.float32))


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[4] at its NAMED shape: 1024 taxa.  The site count of the full configuration (1 Mi over
# 8 GPUs = 131 072 per GPU) is covered twice: every site at n = 4 099 (ragged against all stage sizes), and at
# 131 072 sites per GPU through oracle-checked slices of device-generated tips.
# ---------------------------------------------------------------------------------------------------------
def _stochastic(rng, *shape):
    m = rng.random_sample(shape + (4, 4)) + 0.05
    return (m / m.sum(axis=-1, keepdims=True)).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["balanced", "random"])
@pytest.mark.parametrize("tip_codes", [False, True])
def test_cfg5_shape_1024_taxa_every_site(pkg, coracle, shape, tip_codes):
    n_tips, n = 1024, 4099
    left, right = pkg.balanced_tree(n_tips) if shape == "balanced" else pkg.random_tree(n_tips, seed=5)
    rng = np.random.RandomState(1024)
    ev = rng.random_sample(16).astype(np.float32)
    pl = _stochastic(rng, n_tips - 1, 4).reshape(n_tips - 1, 64)
    pr = _stochastic(rng, n_tips - 1, 4).reshape(n_tips - 1, 64)
    wgt = rng.randint(1, 4, n).astype(np.int32)
    if tip_codes:
        codes = rng.randint(0, 16, (n_tips, n)).astype(np.uint8)
        tip_vector = (rng.random_sample((16, 4)) * 10.0 ** rng.uniform(-6, 0, (16, 1))).astype(np.float32)
        tips = np.stack([expand_tip_codes(codes[i], tip_vector) for i in range(n_tips)])
    else:
        tips = (rng.random_sample((n_tips, n, 16)) * 10.0 ** rng.uniform(-8, 0, (n_tips, n, 1))).astype(np.float32)
    o_root, o_cnt, o_total = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, wgt)
    assert o_cnt.max() >= 3 and np.isfinite(o_root).all(), "the 1024-taxon stimulus should rescale repeatedly and stay finite"
    with pkg.Tree(left, right, n, tip_codes=tip_codes) as t:
        if tip_codes:
            t.write_tip_vector(tip_vector)
        for i in range(n_tips):
            t.write_tip_codes(i, codes[i]) if tip_codes else t.write_tip(i, tips[i])
        t.write_matrices(ev, pl, pr)
        t.write_wgt(wgt)
        for _ in range(2):
            t.run_async()
            root, cnt = t.read_root()
            assert np.array_equal(bits(root), bits(o_root)), first_mismatch(root, o_root)
            assert np.array_equal(cnt, o_cnt) and t.total_scalings() == o_total
        info = t.info()
        assert info["levels"] == (10 if shape == "balanced" else info["levels"]) and info["levels"] >= 10


@pytest.mark.gpu
def test_cfg5_per_gpu_size_slices_match_oracle(pkg, coracle):
    """1024 taxa x 131 072 sites (one GPU's share of cfg5 on 8 GPUs), tips generated on the device exactly as
    bench.py's side workload does; three slices (first sites, an unaligned middle run, the last sites) are recomputed
    by the oracle from the host twin of the generator and compared bit for bit, counts included."""
    import torch
    n_tips, n, first = 1024, 131072, 5 * 131072          # rank 5's site range
    left, right = pkg.balanced_tree(n_tips)
    rng = np.random.RandomState(1)
    ev = _stochastic(rng).reshape(16)
    pl = _stochastic(rng, n_tips - 1, 4).reshape(n_tips - 1, 64)
    pr = _stochastic(rng, n_tips - 1, 4).reshape(n_tips - 1, 64)
    slices = [(0, 384), (65519, 419), (n - 257, 257)]
    with pkg.Tree(left, right, n) as t:
        scratch = torch.empty((n, 16), device="cuda")
        for tip in range(n_tips):
            a, b = (t.tip_ptr(tip), scratch.data_ptr()) if tip % 2 == 0 else (scratch.data_ptr(), t.tip_ptr(tip))
            pkg.generate_device(a, b, first + tip * 7919, n, 1000 + tip)
        torch.cuda.synchronize()
        t.write_matrices(ev, pl, pr)
        t.run_async()
        total = t.total_scalings()
        got = [t.read_root(lo, cnt) for lo, cnt in slices]
        # the counts of the WHOLE root, for the total
        _, all_cnt = t.read_root()
    assert total == int(all_cnt.astype(np.int64).sum()) and all_cnt.max() >= 2
    for (lo, cnt), (g_root, g_cnt) in zip(slices, got):
        tips = np.empty((n_tips, cnt, 16), np.float32)
        for tip in range(n_tips):
            h1, h2 = pkg.generate_host(first + tip * 7919 + lo, cnt, 1000 + tip)
            tips[tip] = (h1 if tip % 2 == 0 else h2).reshape(cnt, 16)
        o_root, o_cnt, _ = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, None)
        assert np.array_equal(bits(g_root), bits(o_root)), (lo, first_mismatch(g_root, o_root))
        assert np.array_equal(g_cnt, o_cnt), lo
