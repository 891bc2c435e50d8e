"""General state count (the reference's STATES knob): 20-state newview.

The reference implements DNA only (README.md:36,202), so the checker is the reference's plf() loop nest with
the state count as a parameter (oracle.plf_oracle_newview_states).  It is pinned where it can be: at S = 4 it
must be bit-identical to the reference's own plf() (committed fixtures + oracle/_ref live).  The CUDA path is
then held to the same bar as the DNA path: STRICT bit-exact (CLVs, scaler bytes, increment), FMA <= 1e-5."""
from __future__ import annotations

import numpy as np
import pytest

import oracle
from conftest import bits

S = 20
SITE = 4 * S
REL_TOL_FMA = 1e-5
SHAPES = [(0, 0), (1, 512), (2, 256), (2, 384), (4, 128), (4, 256)]       # (sites per lane, threads per block)
RAGGED = [1, 7, 8, 9, 15, 16, 17, 31, 33, 100, 1000, 4099, 65536 + 5]


def matrices(seed, signed=False):
    rng = np.random.RandomState(seed)
    draw = (lambda k: rng.standard_normal(k)) if signed else (lambda k: rng.random_sample(k))
    return (draw(S * S).astype(np.float32), draw(4 * S * S).astype(np.float32), draw(4 * S * S).astype(np.float32))


def numpy_newview_states(states, x1, x2, ev, left, right):
    """Independent restatement with numpy fp32 ops (vectorised over sites, sequential over the summed index)."""
    n = x1.shape[0]
    a1 = x1.reshape(n, 4, states)
    a2 = x2.reshape(n, 4, states)
    pl = left.reshape(4, states, states)
    pr = right.reshape(4, states, states)
    evm = ev.reshape(states, states)
    a = np.zeros((n, 4, states), np.float32)
    b = np.zeros((n, 4, states), np.float32)
    for l in range(states):
        a = a + a1[:, :, None, l] * pl[None, :, :, l]
        b = b + a2[:, :, None, l] * pr[None, :, :, l]
    p = a * b
    x3 = np.zeros((n, 4, states), np.float32)
    for k in range(states):
        x3 = x3 + p[:, :, k, None] * evm[None, None, k, :]
    x3 = x3.reshape(n, 4 * states)
    small = (np.abs(x3) < np.float32(2.0 ** -32)).all(axis=1)
    x3[small] = x3[small] * np.float32(2.0 ** 32)
    return x3, small.astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# CPU: the checker itself
# ---------------------------------------------------------------------------------------------
def test_states_oracle_is_pinned_at_four_states(coracle, ref_cases):
    """S = 4 through the general-S loop nest == outputs of the reference's plf() (committed fixtures)."""
    names = sorted({k.rsplit("_", 1)[0] for k in ref_cases.files})
    checked = 0
    for name in names:
        keys = {k.rsplit("_", 1)[1]: k for k in ref_cases.files if k.startswith(name + "_")}
        if not {"x1", "x2", "ev", "left", "right", "x3"} <= set(keys):
            continue
        x1, x2 = ref_cases[keys["x1"]], ref_cases[keys["x2"]]
        wgt = ref_cases[keys["wgt"]] if "wgt" in keys else None
        x3, sc, inc = coracle.newview_states(4, x1, x2, ref_cases[keys["ev"]], ref_cases[keys["left"]],
                                             ref_cases[keys["right"]], wgt)
        assert np.array_equal(bits(x3), bits(ref_cases[keys["x3"]])), name
        if "inc" in keys:
            assert inc == int(ref_cases[keys["inc"]])
        checked += 1
    assert checked >= 1


def test_states_oracle_equals_reference_live(coracle):
    if not oracle.RefOracle.available():
        pytest.skip("oracle/_ref not built")
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(20000, seed=11)
    ref = oracle.RefOracle().newview(x1, x2, ev, left, right)
    x3, sc, inc = coracle.newview_states(4, x1, x2, ev, left, right)
    assert np.array_equal(bits(x3), bits(ref[0])) and inc == ref[-1] == 5000


def test_states_oracle_matches_numpy_restatement_at_twenty(coracle, pkg):
    ev, left, right = matrices(5)
    x1, x2 = pkg.generate_states_host(S, 0, 513, 42)
    x3, sc, inc = coracle.newview_states(S, x1, x2, ev, left, right)
    n3, nsc = numpy_newview_states(S, x1, x2, ev, left, right)
    assert np.array_equal(bits(x3), bits(n3)) and np.array_equal(sc, nsc)
    assert inc == (513 + 3) // 4 and np.array_equal(sc, (np.arange(513) % 4 == 0).astype(np.uint8))
    # multi-threaded driver == single thread
    m3, msc, minc = coracle.newview_states(S, x1, x2, ev, left, right, nthreads=3)
    assert np.array_equal(bits(m3), bits(x3)) and np.array_equal(msc, sc) and minc == inc


def test_states_generator_host(pkg):
    a4, b4 = pkg.generate_states_host(4, 5, 64, 9)
    c4, d4 = pkg.generate_host(5, 64, 9)
    assert np.array_equal(a4, c4) and np.array_equal(b4, d4)          # S = 4 is the DNA generator
    a, b = pkg.generate_states_host(S, 0, 64, 9)
    tiny = a.max(axis=1) < 1e-13
    assert np.array_equal(tiny, np.arange(64) % 4 == 0) and (b > 0).all() and (b < 1).all()
    a2, _ = pkg.generate_states_host(S, 16, 8, 9)                      # any site range independently
    assert np.array_equal(a2, a[16:24])
    with pytest.raises(pkg.PlfError):
        pkg.generate_states_host(7, 0, 4, 1)


def test_states_entry_rejects_unknown_state_counts(pkg):
    with pytest.raises(pkg.PlfError, match="STATES=7"):
        pkg.newview_states_device(7, 16, 16, 16, None, np.zeros(49), np.zeros(196), np.zeros(196), None, 4, None)


# ---------------------------------------------------------------------------------------------
# GPU: parity through the C ABI
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu(pkg):
    assert pkg.device_count() > 0, "no CUDA device visible to libb200plf.so"
    import torch
    assert torch.cuda.is_available()
    return torch


def run_states(pkg, torch, states, ev, left, right, x1, x2, wgt=None, math=0, shape=(0, 0), want_scaler=True, guard=64):
    n = x1.shape[0]
    site = 4 * states
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d1, d2 = to(x1), to(x2)
    d3 = torch.full(((n + guard) * site,), float("nan"), device="cuda")          # guard band behind the output
    dsc = torch.full((n + guard,), 7, dtype=torch.uint8, device="cuda")
    dsum = torch.zeros(1, dtype=torch.int64, device="cuda")
    dw = to(wgt.astype(np.int32)) if wgt is not None else None
    opts = pkg.make_opts(math, shape[0], shape[1])
    pkg.newview_states_device(states, d1.data_ptr(), d2.data_ptr(), d3.data_ptr(), dsc.data_ptr() if want_scaler else None,
                              ev, left, right, dw.data_ptr() if dw is not None else None,
                              n, dsum.data_ptr(), opts, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = d3.cpu().numpy()
    scb = dsc.cpu().numpy()
    assert np.isnan(out[n * site:]).all(), "kernel wrote past the end of x3"
    assert (scb[n:] == 7).all(), "kernel wrote past the end of the scaler bytes"
    return out[: n * site].reshape(n, site), scb[:n], int(dsum.item())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("n", RAGGED)
def test_protein_strict_bit_exact_designed_stimulus(pkg, gpu, coracle, shape, n):
    ev, left, right = matrices(3)
    x1, x2 = pkg.generate_states_host(S, 1000, n, 42)
    o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right)
    g3, gsc, ginc = run_states(pkg, gpu, S, ev, left, right, x1, x2, shape=shape)
    assert np.array_equal(bits(g3), bits(o3)), f"first differing site {np.nonzero((bits(g3) != bits(o3)).any(axis=1))[0][:5]}"
    assert np.array_equal(gsc, osc) and ginc == oinc == sum(1 for s in range(1000, 1000 + n) if s % 4 == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_protein_strict_bit_exact_signed_wide_range(pkg, gpu, coracle, shape):
    """Signed inputs over 13 decades, -0.0, NaN and Inf entries, weights, threshold-boundary sites."""
    n = 20011
    rng = np.random.RandomState(17)
    ev, left, right = matrices(18, signed=True)
    x1 = (rng.standard_normal((n, SITE)) * 10.0 ** rng.uniform(-14, 1, (n, 1))).astype(np.float32)
    x2 = (rng.standard_normal((n, SITE)) * 10.0 ** rng.uniform(-3, 1, (n, 1))).astype(np.float32)
    x1[5] = -0.0
    x1[6, 3] = np.nan
    x2[9, 79] = np.inf
    x1[12] = 0.0
    wgt = rng.randint(0, 50, n).astype(np.int32)
    o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right, wgt)
    assert 0 < osc.sum() < n                              # both branches of the rescale are exercised
    g3, gsc, ginc = run_states(pkg, gpu, S, ev, left, right, x1, x2, wgt=wgt, shape=shape)
    assert np.array_equal(bits(g3), bits(o3))
    assert np.array_equal(gsc, osc) and ginc == oinc


@pytest.mark.gpu
def test_protein_without_scaler_bytes_or_sum(pkg, gpu, coracle):
    ev, left, right = matrices(3)
    x1, x2 = pkg.generate_states_host(S, 0, 777, 1)
    o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right)
    g3, gsc, ginc = run_states(pkg, gpu, S, ev, left, right, x1, x2, want_scaler=False)
    assert np.array_equal(bits(g3), bits(o3)) and ginc == oinc and (gsc == 7).all()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_protein_fma_within_tolerance(pkg, gpu, coracle, shape):
    n = 30000
    ev, left, right = matrices(3)
    x1, x2 = pkg.generate_states_host(S, 0, n, 7)
    o3, osc, oinc = coracle.newview_states(S, x1, x2, ev, left, right)
    g3, gsc, ginc = run_states(pkg, gpu, S, ev, left, right, x1, x2, math=pkg.MATH_FMA, shape=shape)
    rel = np.abs(g3.astype(np.float64) - o3) / np.maximum(np.abs(o3.astype(np.float64)), 1e-300)
    assert rel.max() <= REL_TOL_FMA, f"max relative error {rel.max():.3e}"
    assert np.array_equal(gsc, osc) and ginc == oinc      # designed stimulus: no site near the threshold


@pytest.mark.gpu
def test_protein_four_states_forwards_to_dna_kernel(pkg, gpu, coracle):
    ev, left, right, x1, x2, _ = oracle.host_mem_inputs(5003, seed=4)
    o3, osc, oinc = coracle.newview(x1, x2, ev, left, right)
    g3, gsc, ginc = run_states(pkg, gpu, 4, ev, left, right, x1, x2)
    assert np.array_equal(bits(g3), bits(o3)) and np.array_equal(gsc, osc) and ginc == oinc


@pytest.mark.gpu
def test_protein_large_run_properties(pkg, gpu, coracle):
    """4 Mi sites (3.9 GB of CLVs) generated on the device: designed scaler pattern, increment, and a few slices
    against the oracle; a second launch reproduces the first bit for bit (no schedule dependence)."""
    torch = gpu
    n = 1 << 22
    ev, left, right = matrices(3)
    d1 = torch.empty(n * SITE, device="cuda")
    d2 = torch.empty(n * SITE, device="cuda")
    d3 = torch.empty(n * SITE, device="cuda")
    dsc = torch.empty(n, dtype=torch.uint8, device="cuda")
    dsum = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    pkg.generate_states_device(S, d1.data_ptr(), d2.data_ptr(), 0, n, 42, st)
    pkg.newview_states_device(S, d1.data_ptr(), d2.data_ptr(), d3.data_ptr(), dsc.data_ptr(), ev, left, right,
                              None, n, dsum.data_ptr(), None, st)
    torch.cuda.synchronize()
    assert int(dsum.item()) == n // 4
    assert bool((dsc.view(-1, 4)[:, 0] == 1).all()) and int(dsc.sum().item()) == n // 4
    first = d3.clone()
    pkg.newview_states_device(S, d1.data_ptr(), d2.data_ptr(), d3.data_ptr(), dsc.data_ptr(), ev, left, right,
                              None, n, None, pkg.make_opts(0, 4, 256), st)
    torch.cuda.synchronize()
    assert torch.equal(first.view(torch.int32), d3.view(torch.int32))
    for lo in (0, 123457, n - 333):
        cnt = 333
        h1, h2 = pkg.generate_states_host(S, lo, cnt, 42)
        assert np.array_equal(h1.reshape(-1), d1[lo * SITE:(lo + cnt) * SITE].cpu().numpy())
        o3, _, _ = coracle.newview_states(S, h1, h2, ev, left, right)
        assert np.array_equal(bits(d3[lo * SITE:(lo + cnt) * SITE].cpu().numpy().reshape(cnt, SITE)), bits(o3))


@pytest.mark.gpu
def test_protein_rejects_bad_arguments(pkg, gpu):
    torch = gpu
    d = torch.zeros(SITE * 8 + 4, device="cuda")
    ev, left, right = matrices(1)
    st = torch.cuda.current_stream().cuda_stream
    with pytest.raises(pkg.PlfError, match="16-byte"):
        pkg.newview_states_device(S, d.data_ptr() + 4, d.data_ptr(), d.data_ptr(), None, ev, left, right, None, 8, None, None, st)
    with pytest.raises(pkg.PlfError, match="no 20-state kernel"):
        pkg.newview_states_device(S, d.data_ptr(), d.data_ptr(), d.data_ptr(), None, ev, left, right, None, 8, None,
                                  pkg.make_opts(0, 3, 96), st)
    with pytest.raises(ValueError):
        pkg.newview_states_device(S, d.data_ptr(), d.data_ptr(), d.data_ptr(), None, ev[:16], left, right, None, 8, None, None, st)
    info = pkg.states_kernel_info(S)
    assert info["tile_sites"] == 16 and info["threads"] == 384 and info["smem_bytes"] <= 227 * 1024
    info = pkg.states_kernel_info(S, pkg.MATH_FMA)                  # FMA default for long calls: the tensor-core kernel
    assert info["tile_sites"] == 128 and info["threads"] == 384 and info["regs"] <= 170
    info = pkg.states_kernel_info(S, pkg.MATH_FMA, 4, 256)          # the CUDA-core FMA register tile
    assert info["tile_sites"] == 32 and info["threads"] == 256 and info["regs"] <= 255


# ---------------------------------------------------------------------------------------------
# regression fixtures (tests/golden/aa_cases.npz, made by tests/golden/make_golden_states.py; NOT reference-derived)
# ---------------------------------------------------------------------------------------------
def aa_cases():
    import os
    from conftest import GOLDEN
    d = np.load(os.path.join(GOLDEN, "aa_cases.npz"))
    for name in sorted({k.split("__")[0] for k in d.files}):
        yield name, {k.split("__")[1]: d[k] for k in d.files if k.startswith(name + "__")}


def test_states_oracle_reproduces_committed_fixtures(coracle):
    for name, c in aa_cases():
        x3, sc, inc = coracle.newview_states(S, c["x1"], c["x2"], c["ev"], c["left"], c["right"], c.get("wgt"))
        assert np.array_equal(bits(x3), bits(c["x3"])) and np.array_equal(sc, c["scaler"]) and inc == int(c["inc"]), name


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_protein_matches_committed_fixtures(pkg, gpu, shape):
    for name, c in aa_cases():
        g3, gsc, ginc = run_states(pkg, gpu, S, c["ev"], c["left"], c["right"], c["x1"], c["x2"], wgt=c.get("wgt"), shape=shape)
        assert np.array_equal(bits(g3), bits(c["x3"])) and np.array_equal(gsc, c["scaler"]) and ginc == int(c["inc"]), name
