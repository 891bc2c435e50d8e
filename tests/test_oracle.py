"""Pins the oracle: the C and numpy restatements against the reference's golden vectors
(aie/data/golden*.txt -> aie_kat.json), against committed outputs of the reference's own
plf() (ref_cases.npz, ref_checksums.json), and live against oracle/_ref when it is present.
CPU only."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from oracle import plf_numpy
from conftest import GOLDEN, bits

CASES = ["hostmem100", "hostmem333w", "edge256", "threshold8"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kat_arrays():
    with open(os.path.join(GOLDEN, "aie_kat.json")) as f:
        kat = json.load(f)
    ev = np.array(kat["ev"], dtype=np.float32).reshape(16)
    left = np.empty((4, 4, 4), dtype=np.float32)
    right = np.empty((4, 4, 4), dtype=np.float32)
    x1 = np.empty((1, 4, 4), dtype=np.float32)
    x2 = np.empty((1, 4, 4), dtype=np.float32)
    gold = np.empty((1, 4, 4), dtype=np.float32)
    text = []
    for j, cat in enumerate(kat["categories"]):
        # fixture rows are P^T (file[l][k] == P[k][l]): undo the movers' transpose
        left[j] = np.array(cat["pT_left"], dtype=np.float32).T
        right[j] = np.array(cat["pT_right"], dtype=np.float32).T
        x1[0, j] = cat["x_left"]
        x2[0, j] = cat["x_right"]
        gold[0, j] = cat["golden"]
        text.append(cat["golden_text"])
    return ev, left.reshape(64), right.reshape(64), x1.reshape(1, 16), x2.reshape(1, 16), \
        gold.reshape(1, 16), text


def test_aie_golden_kat_c_and_numpy(coracle):
    ev, left, right, x1, x2, gold, text = kat_arrays()
    for impl in (coracle.newview, plf_numpy.newview):
        x3, sc, inc = impl(x1, x2, ev, left, right)
        assert inc == 0 and sc[0] == 0
        # the golden file prints %.9e, which round-trips fp32: compare text AND values
        got_text = [["%.9e" % abs(v) if v == 0 else "%.9e" % v for v in x3[0, 4 * j:4 * j + 4]]
                    for j in range(4)]
        assert got_text == text
        assert np.array_equal(x3, gold)


def test_transpose_matches_mover_mapping(coracle):
    m = np.arange(16, dtype=np.float32)
    t = coracle.transpose4(m)
    assert np.array_equal(t.reshape(4, 4), m.reshape(4, 4).T)   # transpose.cpp:7-22


@pytest.mark.parametrize("name", CASES)
def test_reference_fixture_bit_exact(coracle, ref_cases, name):
    g = lambda k: ref_cases[f"{name}__{k}"]
    wgt = ref_cases[f"{name}__wgt"] if f"{name}__wgt" in ref_cases.files else None
    for impl in (coracle.newview, plf_numpy.newview):
        x3, sc, inc = impl(g("x1"), g("x2"), g("ev"), g("left"), g("right"), wgt)
        assert inc == int(g("inc"))
        assert np.array_equal(bits(x3), bits(g("x3"))), name
        w = np.ones(len(sc), np.int64) if wgt is None else wgt.astype(np.int64)
        assert int((sc * w).sum()) == inc


def test_threshold_case_semantics(coracle, ref_cases):
    g = lambda k: ref_cases[f"threshold8__{k}"]
    _, sc, inc = coracle.newview(g("x1"), g("x2"), g("ev"), g("left"), g("right"))
    # strict '<' against 2^-32: rows with an element AT or ABOVE the threshold do not scale
    assert sc.tolist() == [1, 0, 0, 1, 0, 1, 1, 0]
    assert inc == 4 == int(g("inc"))


@pytest.mark.parametrize("key", ["hostmem_n4097_seed11", "hostmem_n100000_seed42",
                                 "hostmem_n1000000_seed42"])
def test_reference_checksums(coracle, key):
    with open(os.path.join(GOLDEN, "ref_checksums.json")) as f:
        c = json.load(f)[key]
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(c["n"], c["seed"])
    assert sha(x1) == c["x1_sha256"] and sha(x2) == c["x2_sha256"], "input generator drifted"
    x3, sc, inc = coracle.newview(x1, x2, ev, left, right, wgt)
    assert inc == c["scaler_increment"]
    assert sha(x3) == c["x3_sha256"]
    # the stimulus is built so that exactly sites i % 4 == 0 rescale (host_mem.cpp:198-204)
    assert np.array_equal(np.nonzero(sc)[0], np.arange(0, c["n"], 4))


def test_numpy_equals_c_on_random_signed(coracle):
    rng = np.random.RandomState(99)
    n = 5000
    x1 = (rng.standard_normal((n, 16)) * 10.0 ** rng.uniform(-12, 1, (n, 1))).astype(np.float32)
    x2 = (rng.standard_normal((n, 16)) * 10.0 ** rng.uniform(-12, 1, (n, 1))).astype(np.float32)
    ev = rng.standard_normal(16).astype(np.float32)
    left = rng.standard_normal(64).astype(np.float32)
    right = rng.standard_normal(64).astype(np.float32)
    wgt = rng.randint(1, 50, n).astype(np.int32)
    a = coracle.newview(x1, x2, ev, left, right, wgt)
    b = plf_numpy.newview(x1, x2, ev, left, right, wgt)
    assert np.array_equal(bits(a[0]), bits(b[0]))
    assert np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert 0 < a[1].sum() < n      # both branches exercised


def test_ev4_variant_matches_per_category_runs(coracle):
    rng = np.random.RandomState(5)
    n = 64
    x1 = rng.random_sample((n, 16)).astype(np.float32)
    x2 = rng.random_sample((n, 16)).astype(np.float32)
    ev4 = rng.random_sample(64).astype(np.float32)
    left = rng.random_sample(64).astype(np.float32)
    right = rng.random_sample(64).astype(np.float32)
    x3, sc, inc = coracle.newview(x1, x2, ev4, left, right, ev4=True)
    for j in range(4):
        xj, _, _ = coracle.newview(x1, x2, ev4[16 * j:16 * j + 16], left, right)
        assert np.array_equal(bits(x3[:, 4 * j:4 * j + 4]), bits(xj[:, 4 * j:4 * j + 4]))
    b = plf_numpy.newview(x1, x2, ev4, left, right)
    assert np.array_equal(bits(b[0]), bits(x3))


@pytest.mark.parametrize("layout", [0, 1])
def test_packed_front_end(coracle, layout):
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(257, seed=1)
    lb, rb = oracle.pack_buffers(ev, left, right, x1, x2, layout)
    assert lb.size == 80 + 257 * 16 and rb.size == (80 if layout == 0 else 64) + 257 * 16
    a = coracle.newview(x1, x2, ev, left, right, wgt)
    b = coracle.newview_packed(lb, rb, layout, 257, wgt)
    assert np.array_equal(bits(a[0]), bits(b[0])) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert coracle.scaler_increment(a[1], wgt) == a[2]


def test_mt_driver_equals_serial(coracle):
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(10007, seed=2)
    a = coracle.newview(x1, x2, ev, left, right, wgt)
    for t in (2, 3, 8):
        b = coracle.newview(x1, x2, ev, left, right, wgt, nthreads=t)
        assert np.array_equal(bits(a[0]), bits(b[0])) and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_empty_input(coracle):
    z = np.zeros((0, 16), dtype=np.float32)
    ev, left, right, *_ = oracle.host_mem_inputs(1)
    x3, sc, inc = coracle.newview(z, z, ev, left, right)
    assert x3.shape == (0, 16) and sc.size == 0 and inc == 0


@pytest.mark.skipif(not oracle.RefOracle.available(), reason="oracle/_ref not built here")
def test_live_against_compiled_reference(coracle):
    """The reference's own plf.cpp, compiled in place, vs. our restatement -- live."""
    ref = oracle.RefOracle()
    rng = np.random.RandomState(1234)
    for n in (1, 7, 1000, 65537):
        ev, left, right, x1, x2, _ = oracle.host_mem_inputs(n, seed=n)
        # signed matrices + random magnitudes so both branches and cancellation are hit
        left = (left - 0.5).astype(np.float32)
        x1 = (x1 * 10.0 ** rng.uniform(-8, 0, (n, 1))).astype(np.float32)
        wgt = rng.randint(0, 5, n).astype(np.int32)
        r3, rinc = ref.newview(x1, x2, ev, left, right, wgt)
        o3, osc, oinc = coracle.newview(x1, x2, ev, left, right, wgt)
        assert rinc == oinc
        assert np.array_equal(bits(r3), bits(o3))
        m3, minc = ref.newview(x1, x2, ev, left, right, wgt, nthreads=3)
        assert minc == rinc and np.array_equal(bits(m3), bits(r3))
