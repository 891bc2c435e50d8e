"""Several GPUs from one process (plf_multi_*): the reference's site split (app/src/include.h:181-192) over the GPUs
of a box and the final NCCL reduction.  The split rule and the error behaviour run on the CPU; the one-GPU case runs on
every GPU box; the two-GPU case (real ncclAllReduce) is skipped where fewer than two GPUs are visible."""
from __future__ import annotations

import ctypes

import numpy as np
import pytest

import oracle
from conftest import bits, first_mismatch


def test_partition_is_the_reference_rule(pkg):
    lib = pkg.load()
    from plf_b200 import sharding
    first, cnt = ctypes.c_size_t(0), ctypes.c_size_t(0)
    for n in (0, 1, 7, 100, 4099, 1 << 20, (64 << 20) + 5):
        for parts in (1, 2, 3, 4, 8, 9):
            covered = 0
            for r in range(parts):
                assert lib.plf_multi_partition(n, parts, r, ctypes.byref(first), ctypes.byref(cnt)) == 0
                per = -(-n // parts)                                   # ceil(n / parts), include.h:181-186
                assert first.value == min(per * r, n)
                assert cnt.value == max(0, min(per, n - per * r))      # the last non-empty part takes the remainder
                assert (first.value, cnt.value) == tuple(sharding.shard_for_rank(n, r, parts)) or cnt.value == 0
                covered += cnt.value
            assert covered == n
    assert lib.plf_multi_partition(10, 0, 0, ctypes.byref(first), ctypes.byref(cnt)) != 0
    assert lib.plf_multi_partition(10, 2, 2, ctypes.byref(first), ctypes.byref(cnt)) != 0


def test_multi_create_fails_loudly(pkg):
    lib = pkg.load()
    m = ctypes.c_void_p()
    dup = (ctypes.c_int * 2)(0, 0)
    assert lib.plf_multi_create(ctypes.byref(m), dup, 2, 1, 0, 0) == -1 and b"twice" in lib.plf_multi_last_error(None)
    assert lib.plf_multi_create(ctypes.byref(m), None, 1, 1, 0, 0) == -1
    if pkg.device_count() == 0:
        one = (ctypes.c_int * 1)(0)
        assert lib.plf_multi_create(ctypes.byref(m), one, 1, 1, 0, 0) == -2          # PLF_ERR_CUDA: no CPU fallback
        assert not m.value


@pytest.mark.gpu
def test_multi_single_gpu_matches_oracle(pkg, coracle):
    n = 70001
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(n, seed=7)
    wgt = np.random.RandomState(1).randint(1, 5, n).astype(np.int32)
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    with pkg.Multi([0]) as m:
        x3, sc, inc = m.newview(ev, left, right, x1, x2, wgt)
        assert np.array_equal(bits(x3), bits(o3)), first_mismatch(x3, o3)
        assert np.array_equal(sc, osc) and inc == oinc
        assert m.reduce([5], [1.5]) == (5, 1.5)
        assert m.info() == {"n_devices": 1, "nccl_version": 0, "reductions": 0}     # one GPU: NCCL is never touched
        assert m.partition(n, 0) == (0, n)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [9, 4099, 1 << 20])
def test_multi_two_gpus_shards_and_nccl_totals(pkg, coracle, n):
    if pkg.device_count() < 2:
        pytest.skip("needs two GPUs")
    ev, left, right, x1, x2, _ = oracle.host_mem_inputs(n, seed=n)
    wgt = np.random.RandomState(n).randint(1, 5, n).astype(np.int32)
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    with pkg.Multi([0, 1], n_instances=3) as m:
        assert m.info()["nccl_version"] >= 20000
        # (1) the library's own multi-GPU newview: host arrays in, NCCL-reduced increment out
        x3, sc, inc = m.newview(ev, left, right, x1, x2, wgt)
        assert np.array_equal(bits(x3), bits(o3)), first_mismatch(x3, o3)
        assert np.array_equal(sc, osc) and inc == oinc
        assert m.info()["reductions"] == 1
        # (2) the caller drives the per-GPU contexts (instance API) and reduces: every shard bit for bit
        incs, lnls = [], []
        for r, ctx in enumerate(m.contexts):
            lo, cnt = m.partition(n, r)
            g3, gsc, ginc = ctx.newview(ev, left, right, x1[lo:lo + cnt], x2[lo:lo + cnt], wgt[lo:lo + cnt],
                                        instances=3 if cnt >= 9 else 1)
            assert np.array_equal(bits(g3), bits(o3[lo:lo + cnt])), (r, first_mismatch(g3, o3[lo:lo + cnt]))
            assert np.array_equal(gsc, osc[lo:lo + cnt])
            incs.append(ginc)
            lnls.append(float(np.log(np.abs(g3[:, 0]).astype(np.float64) + 1e-300).sum()))
        tot_inc, tot_lnl = m.reduce(incs, lnls)
        assert tot_inc == oinc == sum(incs)
        assert tot_lnl == lnls[0] + lnls[1]                        # two addends: the fp64 sum is order-independent
        assert m.info()["reductions"] == 2
