"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle, the committed
golden fixtures of the reference, and size-independent properties at BASELINE.json's full sizes.

Bar: STRICT math is bit-exact (CLVs, per-site scaler bytes, scaler increment); FMA math is
within 1e-5 relative (north_star tolerance for the reference's fp32 type) with identical scaler
bytes except for reported threshold-boundary sites."""
from __future__ import annotations

import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, bits, first_mismatch

pytestmark = pytest.mark.gpu

REL_TOL_FMA = 1e-5
# kernel variants exercised everywhere: (variant id, compute threads per block)
VARIANTS = [(0, 0), (2, 256), (1, 128), (4, 256), (12, 256), (3002, 256), (2422, 256), (1421, 128),
            (1324, 256), (1622, 512), (1322, 512), (3222, 128), (1221, 512),
            (1332, 512), (3332, 128), (1334, 256), (1431, 128)]          # K=3: dynamic stage scheduling
RAGGED = [1, 7, 8, 9, 31, 33, 127, 128, 129, 1000, 4097, 65536 + 5]


@pytest.fixture(scope="module")
def gpu(pkg):
    if not os.path.exists(pkg.LIB_PATH):
        raise RuntimeError("libb200plf.so missing on the GPU box -- build() must run first")
    assert pkg.device_count() > 0, "no CUDA device visible to libb200plf.so"
    import torch
    assert torch.cuda.is_available()
    return torch


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_device(pkg, torch, ev, left, right, x1, x2, wgt=None, math=0, variant=0, threads=0,
               ev4=False, blocks_per_sm=0):
    """plf_newview_device on torch-owned device memory; returns numpy (x3, scaler, inc)."""
    n = x1.shape[0]
    d1, d2 = dev(torch, x1), dev(torch, x2)
    dev_ev, dl, dr = dev(torch, ev), dev(torch, left), dev(torch, right)
    d3 = torch.full((max(n, 1), 16), float("nan"), device="cuda")
    dsc = torch.full((max(n, 1),), 7, dtype=torch.uint8, device="cuda")
    dsum = torch.zeros(1, dtype=torch.int64, device="cuda")
    dw = dev(torch, wgt.astype(np.int32)) if wgt is not None else None
    opts = pkg.make_opts(math, variant, threads, blocks_per_sm, 1 if ev4 else 0)
    pkg.newview_device(d1.data_ptr(), d2.data_ptr(), d3.data_ptr(), dsc.data_ptr(),
                       dev_ev.data_ptr(), dl.data_ptr(), dr.data_ptr(),
                       dw.data_ptr() if dw is not None else None, n, dsum.data_ptr(), opts,
                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d3[:n].cpu().numpy(), dsc[:n].cpu().numpy(), int(dsum.item())


def signed_inputs(n, seed):
    rng = np.random.RandomState(seed)
    x1 = (rng.standard_normal((n, 16)) * 10.0 ** rng.uniform(-12, 1, (n, 1))).astype(np.float32)
    x2 = (rng.standard_normal((n, 16)) * 10.0 ** rng.uniform(-12, 1, (n, 1))).astype(np.float32)
    ev = rng.standard_normal(16).astype(np.float32)
    left = rng.standard_normal(64).astype(np.float32)
    right = rng.standard_normal(64).astype(np.float32)
    wgt = rng.randint(0, 50, n).astype(np.int32)
    return ev, left, right, x1, x2, wgt


# ---------------------------------------------------------------------------------------------
# golden vectors of the reference
# ---------------------------------------------------------------------------------------------
def test_aie_golden_kat_through_host_api(pkg, gpu):
    from test_oracle import kat_arrays
    ev, left, right, x1, x2, gold, _ = kat_arrays()
    for layout in (pkg.LAYOUT_COMB, pkg.LAYOUT_SEP):
        with pkg.Context(0, 1, layout) as ctx:
            x3, sc, inc = ctx.newview(ev, left, right, x1, x2)
        assert np.array_equal(x3, gold) and sc[0] == 0 and inc == 0


@pytest.mark.parametrize("name", ["hostmem100", "hostmem333w", "edge256", "threshold8"])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("instances", [1, 3, 9])
def test_reference_fixtures_bit_exact_through_host_api(pkg, gpu, ref_cases, name, layout, instances):
    g = lambda k: ref_cases[f"{name}__{k}"]
    wgt = ref_cases[f"{name}__wgt"] if f"{name}__wgt" in ref_cases.files else None
    n = g("x1").shape[0]
    if not pkg.TestbenchInfo(n, instances).valid():
        pytest.skip("split leaves an empty instance (reference rule)")
    with pkg.Context(0, 9, layout) as ctx:      # NUM_ACCELERATORS=9, `instances` of them used
        x3, sc, inc = ctx.newview(g("ev"), g("left"), g("right"), g("x1"), g("x2"), wgt,
                                  instances=instances)
    assert inc == int(g("inc"))
    assert np.array_equal(bits(x3), bits(g("x3"))), first_mismatch(x3, g("x3"))
    w = np.ones(n, np.int64) if wgt is None else wgt.astype(np.int64)
    assert int((sc.astype(np.int64) * w).sum()) == inc


def test_reference_checksum_1M_sites(pkg, gpu):
    """cfg2 size: 1 000 000 sites, the reference's own plf() output as a SHA-256 fixture."""
    with open(os.path.join(GOLDEN, "ref_checksums.json")) as f:
        c = json.load(f)["hostmem_n1000000_seed42"]
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(c["n"], c["seed"])
    with pkg.Context(0, 9) as ctx:
        for instances in (1, 9):
            x3, sc, inc = ctx.newview(ev, left, right, x1, x2, wgt, instances=instances)
            assert inc == c["scaler_increment"] == 250000
            assert hashlib.sha256(x3.tobytes()).hexdigest() == c["x3_sha256"]
            assert np.array_equal(np.nonzero(sc)[0], np.arange(0, c["n"], 4))


# ---------------------------------------------------------------------------------------------
# every kernel variant vs the oracle, ragged sizes, both arithmetic modes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant,threads", VARIANTS)
def test_variants_strict_bit_exact(pkg, gpu, coracle, variant, threads):
    for n in RAGGED:
        ev, left, right, x1, x2, wgt = signed_inputs(n, seed=n)
        o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
        x3, sc, inc = run_device(pkg, gpu, ev, left, right, x1, x2, wgt, 0, variant, threads)
        assert np.array_equal(bits(x3), bits(o3)), (variant, n, first_mismatch(x3, o3))
        assert np.array_equal(sc, osc), (variant, n)
        assert inc == oinc, (variant, n)
        assert 0 < osc.sum() < n or n < 8


@pytest.mark.parametrize("variant,threads", VARIANTS)
def test_variants_fma_within_tolerance(pkg, gpu, coracle, variant, threads):
    boundary_total = 0
    for n in (129, 4097, 200_000):
        # positive data (CLVs and P are probabilities in the reference stimulus): no cancellation
        ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(n, seed=n)
        rng = np.random.RandomState(n)
        x1 = (x1 * 10.0 ** rng.uniform(-6, 0, (n, 1))).astype(np.float32)
        o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
        x3, sc, inc = run_device(pkg, gpu, ev, left, right, x1, x2, wgt, 1, variant, threads)
        differ = np.nonzero(sc != osc)[0]
        # threshold-boundary disagreements: allowed only where max|x3| is within tol of 2^-32
        for s in differ:
            unscaled = np.abs(o3[s]).max() / (2.0 ** 32 if osc[s] else 1.0)
            assert abs(unscaled - 2.0 ** -32) <= 2 * REL_TOL_FMA * 2.0 ** -32, (variant, n, s)
        boundary_total += len(differ)
        same = sc == osc
        err = np.abs(x3[same].astype(np.float64) - o3[same]) / np.maximum(np.abs(o3[same]), 1e-300)
        assert err.max() <= REL_TOL_FMA, (variant, n, err.max())
    print(f"variant {variant}: {boundary_total} threshold-boundary scaler disagreements")


def test_fma_mode_signed_data_normwise(pkg, gpu, coracle):
    """With cancellation the 1e-5 bound holds relative to the magnitude of the terms summed."""
    n = 50_000
    ev, left, right, x1, x2, wgt = signed_inputs(n, seed=3)
    o3, osc, _ = coracle.newview(x1, x2, ev, left, right, wgt)
    x3, sc, _ = run_device(pkg, gpu, ev, left, right, x1, x2, wgt, 1)
    a1 = np.abs(x1).reshape(n, 4, 4)
    a2 = np.abs(x2).reshape(n, 4, 4)
    al = np.abs(left).reshape(4, 4, 4)
    ar = np.abs(right).reshape(4, 4, 4)
    aev = np.abs(ev).reshape(4, 4)
    pa = np.einsum("njl,jkl->njk", a1.astype(np.float64), al)
    pb = np.einsum("njl,jkl->njk", a2.astype(np.float64), ar)
    bound = np.einsum("njk,kl->njl", pa * pb, aev).reshape(n, 16)
    bound = bound * np.where(osc[:, None] == 1, 2.0 ** 32, 1.0)
    same = sc == osc
    err = np.abs(x3.astype(np.float64) - o3)[same] / np.maximum(bound[same], 1e-300)
    assert err.max() <= REL_TOL_FMA
    assert (~same).sum() <= 2


def test_per_category_ev(pkg, gpu, coracle):
    n = 1000
    ev, left, right, x1, x2, wgt = signed_inputs(n, seed=17)
    ev4 = np.random.RandomState(1).standard_normal(64).astype(np.float32)
    o3, osc, oinc = coracle.newview(x1, x2, ev4, left, right, wgt, ev4=True)
    for variant, threads in ((2, 256), (2422, 256)):
        x3, sc, inc = run_device(pkg, gpu, ev4, left, right, x1, x2, wgt, 0, variant, threads, ev4=True)
        assert np.array_equal(bits(x3), bits(o3)) and np.array_equal(sc, osc) and inc == oinc


@pytest.mark.parametrize("variant,threads", [(0, 0), (2, 256), (4, 256), (1322, 512), (1421, 128), (2422, 256),
                                             (1332, 512), (3332, 128)])
def test_no_out_of_bounds_writes(pkg, gpu, coracle, variant, threads):
    """Guard bands around x3 and the scaler bytes stay untouched for ragged site counts, and input
    buffers that end exactly at the last site are enough (compute-sanitizer is closed on this pool)."""
    torch = gpu
    guard = 64
    for n in (1, 5, 127, 257, 4097, 70001):
        ev, left, right, x1, x2, _ = signed_inputs(n, seed=n + 1)
        o3, osc, oinc = coracle.newview(x1, x2, ev, left, right)
        d1, d2 = dev(torch, x1), dev(torch, x2)                     # exact-size inputs
        mats = [dev(torch, a) for a in (ev, left, right)]
        d3 = torch.full((n + 2 * guard, 16), -7.0, device="cuda")
        dsc = torch.full((n + 2 * guard,), 9, dtype=torch.uint8, device="cuda")
        dsum = torch.zeros(1, dtype=torch.int64, device="cuda")
        pkg.newview_device(d1.data_ptr(), d2.data_ptr(), d3[guard:].data_ptr(), dsc[guard:].data_ptr(),
                           mats[0].data_ptr(), mats[1].data_ptr(), mats[2].data_ptr(), None, n,
                           dsum.data_ptr(), pkg.make_opts(0, variant, threads), 0)
        torch.cuda.synchronize()
        h3, hsc = d3.cpu().numpy(), dsc.cpu().numpy()
        assert (h3[:guard] == -7.0).all() and (h3[guard + n:] == -7.0).all(), (variant, n)
        assert (hsc[:guard] == 9).all() and (hsc[guard + n:] == 9).all(), (variant, n)
        assert np.array_equal(bits(h3[guard:guard + n]), bits(o3)) and np.array_equal(hsc[guard:guard + n], osc)
        assert int(dsum.item()) == oinc


def test_null_scaler_and_null_sum(pkg, gpu, coracle):
    torch = gpu
    n = 777
    ev, left, right, x1, x2, _ = signed_inputs(n, seed=5)
    o3, _, _ = coracle.newview(x1, x2, ev, left, right)
    t = [dev(torch, a) for a in (x1, x2, ev, left, right)]
    d3 = torch.empty((n, 16), device="cuda")
    pkg.newview_device(t[0].data_ptr(), t[1].data_ptr(), d3.data_ptr(), None, t[2].data_ptr(),
                       t[3].data_ptr(), t[4].data_ptr(), None, n, None, None, 0)
    torch.cuda.synchronize()
    assert np.array_equal(bits(d3.cpu().numpy()), bits(o3))


def test_empty_and_misaligned_inputs(pkg, gpu):
    torch = gpu
    d = torch.zeros(256, device="cuda")
    pkg.newview_device(d.data_ptr(), d.data_ptr(), d.data_ptr(), None, d.data_ptr(), d.data_ptr(),
                       d.data_ptr(), None, 0, None, None, 0)            # n == 0 is a no-op
    with pytest.raises(pkg.PlfError) as e:
        pkg.newview_device(d.data_ptr() + 4, d.data_ptr(), d.data_ptr(), None, d.data_ptr(),
                           d.data_ptr(), d.data_ptr(), None, 1, None, None, 0)
    assert e.value.code == -1
    with pytest.raises(pkg.PlfError):
        pkg.newview_device(d.data_ptr(), d.data_ptr(), d.data_ptr(), None, d.data_ptr(),
                           d.data_ptr(), d.data_ptr(), None, 1, None, pkg.make_opts(variant=9999), 0)


# ---------------------------------------------------------------------------------------------
# host-API behaviour (the XRT-like surface)
# ---------------------------------------------------------------------------------------------
def test_host_api_error_behaviour(pkg, gpu):
    with pkg.Context(0, 2) as ctx:
        buf = np.zeros(80 + 16 * 10, np.float32)
        with pytest.raises(pkg.PlfError) as e:
            ctx.run_async(0, 10)                      # run before alloc
        assert e.value.code == -4
        with pytest.raises(pkg.PlfError):
            ctx.instance_alloc(2, 10)                 # instance out of range
        ctx.instance_alloc(0, 10)
        with pytest.raises(pkg.PlfError):
            ctx.write_left(0, buf, buf.nbytes + 4)    # larger than the buffer object
        with pytest.raises(pkg.PlfError):
            ctx.run_async(0, 11)                      # more sites than allocated
        out = np.zeros(16 * 10, np.float32)
        with pytest.raises(pkg.PlfError):
            ctx.read_out(0, out, out.nbytes, 64)      # offset + size beyond the buffer
        ctx.write_left(0, buf)
        ctx.write_right(0, buf)
        ctx.run_async(0, 10)
        ctx.read_out(0, out)
        ctx.wait(0)
        assert ctx.scaler_increment(0) == 10          # all-zero CLVs: every site rescales
        assert not out.any()


def test_instances_are_independent_streams(pkg, gpu, coracle):
    """Different data on each of 9 instances, enqueued back to back, waited at the end."""
    n = 5000
    with pkg.Context(0, 9) as ctx:
        jobs = []
        for k in range(9):
            ev, left, right, x1, x2, wgt = signed_inputs(n + k, seed=100 + k)
            lb, rb = pkg.pack_left(ev, left, x1), pkg.pack_right(ev, right, x2)
            out = np.empty((n + k, 16), np.float32)
            sc = np.empty(n + k, np.uint8)
            ctx.instance_alloc(k, n + k)
            ctx.write_left(k, lb)
            ctx.write_right(k, rb)
            ctx.write_wgt(k, wgt)
            ctx.run_async(k, n + k)
            ctx.read_out(k, out)
            ctx.read_scaler(k, sc)
            jobs.append((lb, rb, out, sc, coracle.newview(x1, x2, ev, left, right, wgt)))
        for k, (_, _, out, sc, (o3, osc, oinc)) in enumerate(jobs):
            ctx.wait(k)
            assert np.array_equal(bits(out), bits(o3)) and np.array_equal(sc, osc)
            assert ctx.scaler_increment(k) == oinc
        assert len({ctx.stream(k) for k in range(9)}) == 9


def test_instances_driven_from_host_threads(pkg, gpu, coracle):
    """The ABI is thread-compatible: distinct instances of one ctx driven from distinct host threads
    (the reference drives each instance from its own xrt::queue workers, host_mem.cpp:249-260)."""
    import threading
    n, calls, inst = 20000, 20, 6
    data = [signed_inputs(n, seed=300 + k) for k in range(inst)]
    want = [coracle.newview(x1, x2, ev, left, right, wgt) for ev, left, right, x1, x2, wgt in data]
    errors = []
    with pkg.Context(0, inst) as ctx:
        def worker(k):
            try:
                ev, left, right, x1, x2, wgt = data[k]
                lb, rb = pkg.pack_left(ev, left, x1), pkg.pack_right(ev, right, x2)
                out = np.empty((n, 16), np.float32)
                sc = np.empty(n, np.uint8)
                ctx.instance_alloc(k, n)
                for _ in range(calls):
                    ctx.write_left(k, lb)
                    ctx.write_right(k, rb)
                    ctx.write_wgt(k, wgt)
                    ctx.run_async(k, n)
                    ctx.read_out(k, out)
                    ctx.read_scaler(k, sc)
                    ctx.wait(k)
                    assert ctx.scaler_increment(k) == want[k][2]
                    assert np.array_equal(bits(out), bits(want[k][0])) and np.array_equal(sc, want[k][1])
            except Exception as e:          # surfaced in the main thread
                errors.append((k, repr(e)))
        threads = [threading.Thread(target=worker, args=(k,)) for k in range(inst)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    assert not errors, errors


def test_repeated_calls_reuse_buffers(pkg, gpu, coracle):
    """plf_calls > 1: the run handle and buffers are reused (host_mem.cpp:283-325)."""
    n = 3000
    ev, left, right, x1, x2, wgt = signed_inputs(n, seed=8)
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    with pkg.Context(0, 1) as ctx:
        ctx.instance_alloc(0, n)
        lb, rb = pkg.pack_left(ev, left, x1), pkg.pack_right(ev, right, x2)
        out = np.empty((n, 16), np.float32)
        for _ in range(5):
            ctx.write_left(0, lb)
            ctx.write_right(0, rb)
            ctx.write_wgt(0, wgt)
            ctx.run_async(0, n)
            ctx.read_out(0, out)
            assert ctx.scaler_increment(0) == oinc      # not accumulated across calls
            assert np.array_equal(bits(out), bits(o3))


def test_scaler_counter_ring_wraps(pkg, gpu, coracle):
    """The per-run scaler counters live in a ring that is re-zeroed every 1024 runs."""
    n = 64
    ev, left, right, x1, x2, wgt = signed_inputs(n, seed=21)
    x1[::2] = 0.0                                   # half of the sites rescale
    _, _, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    with pkg.Context(0, 1) as ctx:
        ctx.instance_alloc(0, n)
        assert ctx.scaler_increment(0) == 0          # nothing has run yet
        ctx.write_left(0, pkg.pack_left(ev, left, x1))
        ctx.write_right(0, pkg.pack_right(ev, right, x2))
        ctx.write_wgt(0, wgt)
        for call in range(2100):
            ctx.run_async(0, n)
            if call in (0, 1, 1022, 1023, 1024, 1025, 2047, 2048, 2099):
                assert ctx.scaler_increment(0) == oinc, call


def test_pinned_host_buffers_and_marks(pkg, gpu, coracle):
    n = 100_000
    ev, left, right, x1, x2, _ = oracle.host_mem_inputs(n, seed=4)
    lbp, lb_ptr = pkg.host_alloc((80 + 16 * n) * 4, np.float32)
    rbp, rb_ptr = pkg.host_alloc((80 + 16 * n) * 4, np.float32)
    outp, out_ptr = pkg.host_alloc(16 * n * 4, np.float32)
    try:
        lbp[:] = pkg.pack_left(ev, left, x1)
        rbp[:] = pkg.pack_right(ev, right, x2)
        with pkg.Context(0, 1) as ctx:
            ctx.instance_alloc(0, n)
            ctx.mark(0, pkg.MARK_BEGIN)
            ctx.write_left(0, lbp)
            ctx.write_right(0, rbp)
            ctx.mark(0, pkg.MARK_T1)
            ctx.run_async(0, n)
            ctx.mark(0, pkg.MARK_T2)
            ctx.read_out(0, outp)
            ctx.mark(0, pkg.MARK_END)
            ctx.wait(0)
            hm, msm, mh = (ctx.elapsed_ms(0, a, b) for a, b in ((0, 1), (1, 2), (2, 3)))
            assert hm > 0 and msm > 0 and mh > 0
            o3, _, oinc = coracle.newview(x1, x2, ev, left, right)
            assert np.array_equal(bits(outp.reshape(n, 16)), bits(o3))
            assert ctx.scaler_increment(0) == oinc == n // 4
    finally:
        for p in (lb_ptr, rb_ptr, out_ptr):
            pkg.host_free(p)


# ---------------------------------------------------------------------------------------------
# INPUT_SRC=gen analogue
# ---------------------------------------------------------------------------------------------
def test_gen_mode_matches_mem_run_of_the_pattern(pkg, gpu, coracle):
    n = 4099
    p1, p2, ev4, pl, pr = pkg.gen_pattern()
    x1 = np.tile(p1, (n, 1))
    x2 = np.tile(p2, (n, 1))
    o3, osc, oinc = coracle.newview(x1, x2, ev4, pl, pr, ev4=True)
    for math_mode in (pkg.MATH_STRICT, pkg.MATH_FMA):
        with pkg.Context(0, 2, input_src=pkg.INPUT_GEN) as ctx:
            ctx.set_math(math_mode)
            ctx.instance_alloc(1, n)
            ctx.run_async(1, n)
            out = np.empty((n, 16), np.float32)
            sc = np.empty(n, np.uint8)
            ctx.read_out(1, out)
            ctx.read_scaler(1, sc)
            ctx.wait(1)
            if math_mode == pkg.MATH_STRICT:
                assert np.array_equal(bits(out), bits(o3))
            else:
                assert np.allclose(out, o3, rtol=REL_TOL_FMA, atol=0)
            assert np.array_equal(sc, osc) and ctx.scaler_increment(1) == oinc
            with pytest.raises(pkg.PlfError):
                ctx.write_left(1, x1)                  # gen instances have no input buffers
            # cfg4b: discard sink -> checksum only
            ctx.set_gen_sink(pkg.GEN_DISCARD)
            ctx.run_async(1, n)
            chk = ctx.gen_checksum(1)
            assert abs(chk - o3.astype(np.float64).sum()) <= 1e-5 * abs(o3.astype(np.float64).sum())


def test_device_generator_matches_host_generator(pkg, gpu):
    torch = gpu
    n, first = 10_000, 123_456_789
    d1 = torch.empty((n, 16), device="cuda")
    d2 = torch.empty((n, 16), device="cuda")
    pkg.generate_device(d1.data_ptr(), d2.data_ptr(), first, n, 42)
    torch.cuda.synchronize()
    h1, h2 = pkg.generate_host(first, n, 42)
    assert np.array_equal(bits(d1.cpu().numpy()), bits(h1))
    assert np.array_equal(bits(d2.cpu().numpy()), bits(h2))


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes through size-independent properties
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_sites", [1 << 20, 64 << 20])
def test_full_size_properties(pkg, gpu, coracle, n_sites):
    """cfg2 (1 Mi) and cfg3 (64 Mi sites, 12 GiB): device-generated stimulus; checks
    (a) exactly the sites i % 4 == 0 rescale, (b) the fused scaler sum equals the byte sum,
    (c) random slices are bit-identical to the oracle run on host-regenerated inputs,
    (d) a second run is bit-identical (determinism), (e) linear checksum of checksums."""
    torch = gpu
    free, _ = torch.cuda.mem_get_info()
    if free < n_sites * 200 + (1 << 30):
        pytest.skip("not enough device memory")
    seed = 42
    ev, left, right, *_ = oracle.host_mem_inputs(1, seed=seed)
    d1 = torch.empty((n_sites, 16), device="cuda")
    d2 = torch.empty((n_sites, 16), device="cuda")
    d3 = torch.empty((n_sites, 16), device="cuda")
    dsc = torch.empty(n_sites, dtype=torch.uint8, device="cuda")
    dsum = torch.zeros(1, dtype=torch.int64, device="cuda")
    dev_ev, dl, dr = dev(torch, ev), dev(torch, left), dev(torch, right)
    pkg.generate_device(d1.data_ptr(), d2.data_ptr(), 0, n_sites, seed)
    args = (d1.data_ptr(), d2.data_ptr(), d3.data_ptr(), dsc.data_ptr(), dev_ev.data_ptr(),
            dl.data_ptr(), dr.data_ptr(), None, n_sites, dsum.data_ptr(), None, 0)
    pkg.newview_device(*args)
    torch.cuda.synchronize()
    assert int(dsum.item()) == n_sites // 4                                    # (a)+(b)
    assert int(dsc.sum(dtype=torch.int64).item()) == n_sites // 4
    assert bool((dsc.view(-1, 4)[:, 0] == 1).all()) and int(dsc.view(-1, 4)[:, 1:].sum().item()) == 0
    rng = np.random.RandomState(0)
    starts = [0, n_sites - 4096] + list(rng.randint(0, n_sites - 4096, 6))
    for s in starts:                                                           # (c)
        h1, h2 = pkg.generate_host(int(s), 4096, seed)
        o3, osc, _ = coracle.newview(h1, h2, ev, left, right)
        assert np.array_equal(bits(d3[s:s + 4096].cpu().numpy()), bits(o3))
        assert np.array_equal(dsc[s:s + 4096].cpu().numpy(), osc)
    first = d3.view(torch.int32).to(torch.int64).sum().item()                  # (d)+(e)
    d3.zero_()
    dsum.zero_()
    pkg.newview_device(*args)
    torch.cuda.synchronize()
    assert d3.view(torch.int32).to(torch.int64).sum().item() == first
    assert torch.isfinite(d3).all()


@pytest.mark.parametrize("n,chunk", [(1, 0), (300001, 65536), (70000, 1 << 20), (262144, 65536)])
def test_streamed_host_path_matches_oracle(pkg, gpu, coracle, n, chunk):
    """plf_newview_stream: chunked / overlapped round trip over unpacked host arrays, bit-exact, incl.
    a ragged last chunk, weights and the scaler bytes; buffers are reused across calls and resized."""
    ev, left, right, x1, x2, wgt = signed_inputs(n, seed=n + 7)
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    with pkg.Context(0, 1) as ctx:
        for _ in range(2):
            x3 = np.full((n, 16), np.nan, np.float32)
            sc = np.full(n, 7, np.uint8)
            inc = ctx.newview_stream(ev, left, right, x1, x2, x3, sc, wgt, chunk_sites=chunk)
            assert inc == oinc and np.array_equal(sc, osc)
            assert np.array_equal(bits(x3), bits(o3)), first_mismatch(x3, o3)
        x3b = np.empty((n, 16), np.float32)
        assert ctx.newview_stream(ev, left, right, x1, x2, x3b, None, None, chunk_sites=max(256, chunk // 2)) == int(osc.sum())
        assert np.array_equal(bits(x3b), bits(o3))


def test_stream_trace_dump_is_a_consistent_timeline(pkg, gpu, coracle, tmp_path, monkeypatch):
    """PLF_STREAM_TRACE: one line per chunk with four non-decreasing event times, every site accounted for, and the traced
    call returns the same bits (the events only break the programmatic launch overlap)."""
    n, chunk = 300001, 65536
    ev, left, right, x1, x2, wgt = signed_inputs(n, seed=11)
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
    path = tmp_path / "stream_trace.txt"
    monkeypatch.setenv("PLF_STREAM_TRACE", str(path))
    with pkg.Context(0, 1) as ctx:
        x3 = np.full((n, 16), np.nan, np.float32)
        sc = np.full(n, 7, np.uint8)
        assert ctx.newview_stream(ev, left, right, x1, x2, x3, sc, wgt, chunk_sites=chunk) == oinc
    assert np.array_equal(bits(x3), bits(o3)) and np.array_equal(sc, osc)
    rows = np.loadtxt(path, comments="#", ndmin=2)
    assert rows.shape == ((n + chunk - 1) // chunk, 7)
    assert rows[:, 2].sum() == n and (rows[:, 1] == np.arange(len(rows)) % 3).all()
    t = rows[:, 3:]
    assert (np.diff(t, axis=1) >= 0).all() and t.min() >= 0 and t.max() < 10e3
