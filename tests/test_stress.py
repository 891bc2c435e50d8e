"""Slot-release stress for the bulk-copy ring kernels.

The write-after-read hazard of a shared-memory ring: a slot handed back to the bulk-copy engine while a warp is still
reading it.  It shows when refills are FAST (inputs resident in L2) and one block does all the work, so these tests run
the DNA ring kernel (plf_newview_tma_dyn, the default), the 20-state kernel (plf_newview_aa) and the tensor-core
20-state kernel (plf_newview_aa_tc) on ONE block
(PLF_LAUNCH_SINGLE_CTA), with freshly poisoned device memory every iteration, in BOTH release modes (fence.proxy.async
and data dependency), and compare every bit with the oracle.  The tree kernel's twin lives in tests/test_tree.py."""
from __future__ import annotations

import numpy as np
import pytest

import oracle
from conftest import bits, first_mismatch

ITERS = 25


def _poison(torch):
    p = torch.full((48 << 20,), float("nan"), device="cuda")        # 192 MiB of NaN through the allocator's pool
    del p
    torch.cuda.empty_cache()


@pytest.mark.gpu
@pytest.mark.parametrize("variant,threads", [(0, 0), (1432, 512), (1422, 512), (2332, 256), (1232, 512), (3632, 128)])
def test_dna_ring_single_block_poisoned_memory_both_release_modes(pkg, coracle, variant, threads):
    import torch
    n = 24001                                                       # 94 ragged stages of 256 sites: 3 MB, L2-resident
    ev, left, right, x1, x2, _ = oracle.host_mem_inputs(n, seed=3)
    wgt = np.random.RandomState(3).randint(1, 4, n).astype(np.int32)
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    stream = torch.cuda.current_stream().cuda_stream
    for it in range(ITERS):
        _poison(torch)
        d = [torch.from_numpy(a).cuda() for a in (x1, x2, ev, left, right, wgt)]
        for mode in (pkg.LAUNCH_FENCED_RELEASE, pkg.LAUNCH_DEP_RELEASE):
            g3 = torch.full((n, 16), float("nan"), device="cuda")
            gsc = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
            gsum = torch.zeros(1, dtype=torch.int64, device="cuda")
            opts = pkg.make_opts(pkg.MATH_STRICT, variant, threads, 0, 0, mode | pkg.LAUNCH_SINGLE_CTA)
            for _ in range(2):                                      # back to back: the second launch finds everything in L2
                pkg.newview_device(d[0].data_ptr(), d[1].data_ptr(), g3.data_ptr(), gsc.data_ptr(), d[2].data_ptr(),
                                   d[3].data_ptr(), d[4].data_ptr(), d[5].data_ptr(), n, gsum.data_ptr(), opts, stream)
            torch.cuda.synchronize()
            got = g3.cpu().numpy()
            assert np.array_equal(bits(got), bits(o3)), (it, mode, first_mismatch(got, o3))
            assert np.array_equal(gsc.cpu().numpy(), osc) and int(gsum.item()) == 2 * oinc


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(0, 0), (4, 256), (2, 384), (1, 512)])
def test_protein_single_block_poisoned_memory_both_release_modes(pkg, coracle, shape):
    import torch
    n = 6007
    rng = np.random.RandomState(20)
    ev, left, right = (rng.random_sample(k).astype(np.float32) for k in (400, 1600, 1600))
    x1, x2 = pkg.generate_states_host(20, 0, n, 9)
    o3, osc, oinc = coracle.newview_states(20, x1, x2, ev, left, right)
    stream = torch.cuda.current_stream().cuda_stream
    for it in range(ITERS):
        _poison(torch)
        d1, d2 = torch.from_numpy(x1).cuda(), torch.from_numpy(x2).cuda()
        for mode in (pkg.LAUNCH_FENCED_RELEASE, pkg.LAUNCH_DEP_RELEASE):
            g3 = torch.full((n * 80,), float("nan"), device="cuda")
            gsc = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
            gsum = torch.zeros(1, dtype=torch.int64, device="cuda")
            opts = pkg.make_opts(pkg.MATH_STRICT, shape[0], shape[1], 0, 0, mode | pkg.LAUNCH_SINGLE_CTA)
            for _ in range(2):
                pkg.newview_states_device(20, d1.data_ptr(), d2.data_ptr(), g3.data_ptr(), gsc.data_ptr(), ev, left, right,
                                          None, n, gsum.data_ptr(), opts, stream)
            torch.cuda.synchronize()
            got = g3.cpu().numpy().reshape(o3.shape)
            assert np.array_equal(bits(got), bits(o3)), (it, mode, first_mismatch(got, o3))
            assert np.array_equal(gsc.cpu().numpy(), osc) and int(gsum.item()) == 2 * oinc


@pytest.mark.gpu
def test_tensor_core_kernel_single_block_poisoned_memory_both_release_modes(pkg, coracle):
    """The tcgen05 20-state kernel's operand ring (4 boxes per group, refilled by TMA tensor copies) and its staging boxes
    (read by TMA tensor stores): one block, L2-resident inputs, both release modes, poisoned memory; results within the
    mode's tolerance, identical between the two release modes and from launch to launch."""
    import torch
    n = 20011                                                       # 157 ragged tiles of 128 sites, 19 MB: L2-resident
    rng = np.random.RandomState(21)
    ev, left, right = (rng.random_sample(k).astype(np.float32) for k in (400, 1600, 1600))
    x1, x2 = pkg.generate_states_host(20, 0, n, 13)
    o3, osc, oinc = coracle.newview_states(20, x1, x2, ev, left, right)
    stream = torch.cuda.current_stream().cuda_stream
    first = None
    for it in range(10):
        _poison(torch)
        d1, d2 = torch.from_numpy(x1).cuda(), torch.from_numpy(x2).cuda()
        for mode in (pkg.LAUNCH_FENCED_RELEASE, pkg.LAUNCH_DEP_RELEASE):
            g3 = torch.full((n * 80,), float("nan"), device="cuda")
            gsc = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
            gsum = torch.zeros(1, dtype=torch.int64, device="cuda")
            opts = pkg.make_opts(pkg.MATH_FMA, 9, 0, 0, 0, mode | pkg.LAUNCH_SINGLE_CTA)
            for _ in range(2):
                pkg.newview_states_device(20, d1.data_ptr(), d2.data_ptr(), g3.data_ptr(), gsc.data_ptr(), ev, left, right,
                                          None, n, gsum.data_ptr(), opts, stream)
            torch.cuda.synchronize()
            got = g3.cpu().numpy().reshape(o3.shape)
            if first is None:
                first = got.copy()
                rel = np.abs(got.astype(np.float64) - o3) / np.maximum(np.abs(o3.astype(np.float64)), 1e-300)
                assert rel.max() <= 1e-5, rel.max()
            assert np.array_equal(bits(got), bits(first)), (it, mode, first_mismatch(got, first))
            assert np.array_equal(gsc.cpu().numpy(), osc) and int(gsum.item()) == 2 * oinc


@pytest.mark.gpu
def test_release_mode_switch_reaches_the_tree_kernel(pkg, coracle):
    """plf_set_release_mode(1) (= PLF_SAFE_RELEASE=1) re-captures the tree's graph with the fenced release: same bits."""
    from oracle import tree_oracle
    from test_tree import tree_inputs
    left, right = pkg.random_tree(65, seed=2)
    tips, ev, pl, pr, wgt = tree_inputs(65, 1500, seed=65)
    o_root, o_cnt, o_total = tree_oracle.traverse(coracle, left, right, tips, ev, pl, pr, wgt)
    try:
        with pkg.Tree(left, right, 1500) as t:
            for i in range(65):
                t.write_tip(i, tips[i])
            t.write_matrices(ev, pl, pr)
            t.write_wgt(wgt)
            for mode in (1, 0, -1, 1):
                pkg.set_release_mode(mode)
                t.set_tuning(0, 1000)
                t.run_async()
                root, cnt = t.read_root()
                assert np.array_equal(bits(root), bits(o_root)), (mode, first_mismatch(root, o_root))
                assert np.array_equal(cnt, o_cnt) and t.total_scalings() == o_total
    finally:
        pkg.set_release_mode(-1)
