#!/usr/bin/env python
"""tests/golden/make_golden_states.py -- regression fixtures for the 20-state path (tests/golden/aa_cases.npz).

NOT derived from the reference: the reference has no protein implementation (README.md:36,202), so these are outputs
of oracle.plf_oracle_newview_states (the reference's loop nest with the state count as a parameter, pinned at S = 4
against the reference's plf()).  They freeze today's bits so that a later change of the checker or of the kernel
that alters a result is noticed; parity at S = 20 stays UNPINNED."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from conftest import load_pkg  # noqa: E402

S = 20


def main():
    co = oracle.COracle()
    pkg = load_pkg()
    cases = {}
    # 1. the designed stimulus (every 4th site rescales), uniform(0,1) matrices
    rng = np.random.RandomState(20)
    ev, left, right = (rng.random_sample(k).astype(np.float32) for k in (S * S, 4 * S * S, 4 * S * S))
    x1, x2 = pkg.generate_states_host(S, 0, 64, 42)
    x3, sc, inc = co.newview_states(S, x1, x2, ev, left, right)
    assert inc == 16
    cases.update(designed64__ev=ev, designed64__left=left, designed64__right=right, designed64__x1=x1, designed64__x2=x2,
                 designed64__x3=x3, designed64__scaler=sc, designed64__inc=np.int64(inc))
    # 2. signed matrices, magnitudes over 14 decades, -0.0 / denormal / inf entries, weights
    rng = np.random.RandomState(21)
    n = 48
    ev, left, right = (rng.standard_normal(k).astype(np.float32) for k in (S * S, 4 * S * S, 4 * S * S))
    x1 = (rng.standard_normal((n, 4 * S)) * 10.0 ** rng.uniform(-14, 1, (n, 1))).astype(np.float32)
    x2 = (rng.standard_normal((n, 4 * S)) * 10.0 ** rng.uniform(-3, 1, (n, 1))).astype(np.float32)
    x1[0] = -0.0
    x1[1] = np.float32(1e-42)
    x2[2, 5] = np.inf
    x1[3] = 0.0
    wgt = rng.randint(0, 9, n).astype(np.int32)
    x3, sc, inc = co.newview_states(S, x1, x2, ev, left, right, wgt)
    assert 0 < sc.sum() < n
    cases.update(signed48__ev=ev, signed48__left=left, signed48__right=right, signed48__x1=x1, signed48__x2=x2,
                 signed48__wgt=wgt, signed48__x3=x3, signed48__scaler=sc, signed48__inc=np.int64(inc))
    np.savez_compressed(os.path.join(HERE, "aa_cases.npz"), **cases)
    print("wrote aa_cases.npz:", sorted(cases))


if __name__ == "__main__":
    main()
