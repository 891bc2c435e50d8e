#!/usr/bin/env python
"""tests/golden/make_golden.py -- (re)generate the committed golden fixtures.

Runs ONLY in the build container, where /root/reference exists.  It
  1. lifts the reference's own known-answer vectors aie/data/{inputEV0,inputbranch*,inputdata*,
     golden*}.txt into tests/golden/aie_kat.json (values are facts), and
  2. runs the reference's own plf() (app/src/plf.cpp, compiled in place into
     oracle/_ref/libplf_ref.so by oracle/Makefile) on seeded inputs and stores
       - small cases with inputs+outputs   -> tests/golden/ref_cases.npz
       - large cases as SHA-256 checksums  -> tests/golden/ref_checksums.json
Nothing here is read from /root/reference at test time: the tests consume only the fixtures.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402

REF = "/root/reference"


def read_rows(path):
    with open(path) as f:
        return [[float(t) for t in line.split()] for line in f if line.strip()]


def lift_aie_kat():
    d = os.path.join(REF, "aie", "data")
    ev = read_rows(os.path.join(d, "inputEV0.txt"))
    assert len(ev) == 4
    kat = {"source": "aie/data/{inputEV0,inputbranch{left,right}N,inputdata{left,right}N,goldenN}.txt",
           "note": "branch files hold P already transposed: file[l][k] == P[k][l]; "
                   "EV is used as stored ([k][l]); golden rows are the ev-kernel output of lane N",
           "ev": ev, "categories": []}
    for j in range(4):
        cat = {}
        for side in ("left", "right"):
            rows = read_rows(os.path.join(d, f"inputdata{side}{j}.txt"))
            assert all(r == rows[0] for r in rows), "stimulus rows differ"
            pt = read_rows(os.path.join(d, f"inputbranch{side}{j}.txt"))
            assert len(pt) == 4
            cat[f"x_{side}"] = rows[0]
            cat[f"pT_{side}"] = pt
        g = read_rows(os.path.join(d, f"golden{j}.txt"))
        assert all(r == g[0] for r in g), "golden rows differ"
        cat["golden"] = g[0]
        with open(os.path.join(d, f"golden{j}.txt")) as f:
            cat["golden_text"] = f.readline().split()
        kat["categories"].append(cat)
    with open(os.path.join(HERE, "aie_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote aie_kat.json")


def edge_inputs(seed=7, n=256):
    """Signed, denormal, zero, inf/nan and mixed-magnitude CLVs with signed P/EV."""
    rng = np.random.RandomState(seed)
    ev = rng.uniform(-1, 1, 16).astype(np.float32)
    left = rng.uniform(-1.5, 1.5, 64).astype(np.float32)
    right = rng.uniform(-1.5, 1.5, 64).astype(np.float32)
    x1 = rng.uniform(-1, 1, (n, 16)).astype(np.float32)
    x2 = rng.uniform(-1, 1, (n, 16)).astype(np.float32)
    mag = (10.0 ** rng.uniform(-30, 2, (n, 1))).astype(np.float32)   # per-site magnitude sweep
    x1 = (x1 * mag).astype(np.float32)
    x1[0] = 0.0
    x1[1] = -0.0
    x2[2] = 0.0
    x1[3] = np.float32(1e-42)                  # denormal inputs
    x1[4, 5] = np.nan
    x1[5, 0] = np.inf
    x1[6] = np.float32(1e-45)
    x2[6] = np.float32(1e-3)
    x1[7] = np.float32(3e38)                   # overflow to inf inside the mat-vec
    x2[7] = np.float32(3e38)
    x1[8] = -np.abs(x1[8]) * np.float32(1e-20)
    return ev, left, right, x1, x2


def threshold_inputs():
    """Identity P/EV so that x3 == x1*x2 exactly; x1 straddles 2^-32 by ulps."""
    eye = np.eye(4, dtype=np.float32).reshape(16)
    ev = eye.copy()
    left = np.tile(eye, 4)
    right = np.tile(eye, 4)
    t = np.float32(2.0 ** -32)
    below = np.nextafter(t, np.float32(0))
    above = np.nextafter(t, np.float32(1))
    n = 8
    x2 = np.ones((n, 16), dtype=np.float32)
    x1 = np.full((n, 16), below, dtype=np.float32)
    x1[1, 15] = t            # one element exactly at the threshold: NOT scaled (strict <)
    x1[2, 0] = above
    x1[3] = -below           # all negative, just below: scaled
    x1[4, 7] = -t            # |x| == threshold: not scaled
    x1[5] = 0.0              # all zero: scaled (0 < 2^-32), stays 0
    x1[6] = np.float32(1e-45)  # smallest denormal: scaled exactly by 2^32
    x1[7, 3] = np.float32(1.0)
    return ev, left, right, x1, x2


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference():
    oracle.build(quiet=False)
    ref = oracle.RefOracle()
    cases = {}
    sums = {}

    def add_small(name, ev, left, right, x1, x2, wgt=None):
        x3, inc = ref.newview(x1, x2, ev, left, right, wgt)
        cases[f"{name}__ev"] = ev
        cases[f"{name}__left"] = left
        cases[f"{name}__right"] = right
        cases[f"{name}__x1"] = x1
        cases[f"{name}__x2"] = x2
        if wgt is not None:
            cases[f"{name}__wgt"] = wgt
        cases[f"{name}__x3"] = x3
        cases[f"{name}__inc"] = np.int64(inc)

    # cfg1: the reference's default run size (Makefile:15 ALIGNMENTS=100), seeded recipe
    ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(100, seed=42)
    add_small("hostmem100", ev, left, right, x1, x2, wgt)
    # weighted variant (wgt is all ones in the reference host; plf() honours any ints)
    rng = np.random.RandomState(3)
    ev, left, right, x1, x2, _ = oracle.host_mem_inputs(333, seed=5)
    add_small("hostmem333w", ev, left, right, x1, x2, rng.randint(0, 9, 333).astype(np.int32))
    ev, left, right, x1, x2 = edge_inputs()
    add_small("edge256", ev, left, right, x1, x2)
    ev, left, right, x1, x2 = threshold_inputs()
    add_small("threshold8", ev, left, right, x1, x2)

    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **cases)
    print("wrote ref_cases.npz")

    # large cases: checksums only (inputs are regenerated from the seed at test time)
    for n, seed in ((4097, 11), (100_000, 42), (1_000_000, 42)):
        ev, left, right, x1, x2, wgt = oracle.host_mem_inputs(n, seed=seed)
        x3, inc = ref.newview(x1, x2, ev, left, right, wgt)
        sums[f"hostmem_n{n}_seed{seed}"] = {
            "n": n, "seed": seed, "x3_sha256": sha(x3), "scaler_increment": inc,
            "x1_sha256": sha(x1), "x2_sha256": sha(x2)}
    with open(os.path.join(HERE, "ref_checksums.json"), "w") as f:
        json.dump(sums, f, indent=1)
    print("wrote ref_checksums.json")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    lift_aie_kat()
    run_reference()
