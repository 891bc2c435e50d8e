"""oracle/evaluate_oracle.py -- root log-likelihood across a branch on the CPU (float64 numpy).
TEST INFRASTRUCTURE ONLY.

This step is not in /root/reference (which stops at newview, app/src/plf.cpp), so no reference golden
vector can exist for it.  It restates evaluateGTRGAMMA of standard-RAxML (evaluateGenericSpecial.c, the
code base the reference's plf() derives from, README.md:188-189,207-208) for the reference's CLV layout:

    term_i = log(0.25 * |sum_{j,k} x1[i,j,k] * x2[i,j,k] * diag[j,k]|) + (cnt1_i + cnt2_i) * log(2^-32)
    lnL    = sum_i wgt_i * term_i

PARITY: pinned against an INDEPENDENT model instead of a reference vector -- oracle/felsenstein_fp64.py
(textbook state-space Felsenstein pruning with explicit GTR+Gamma4 transition matrices in float64, itself
checked against 50-digit mpmath).  tests/test_felsenstein.py chains the pinned newview oracle up a tree,
applies this function across the root branch and requires the textbook log-likelihood to 1e-6 relative
(measured 3e-8), including trees that rescale up to 13 times per site; the same test drives the CUDA path.
The GPU kernel is additionally compared with this restatement directly (1e-9 relative: fp64 on both sides)."""
from __future__ import annotations

import numpy as np

LOG_MIN = -32.0 * np.log(2.0)


def evaluate(x1, x2, diag, cnt1=None, cnt2=None, wgt=None) -> float:
    """diag has 4*S entries ([category][state]); S = 4 for the reference's DNA layout, 20 for protein."""
    d = np.asarray(diag, np.float32).astype(np.float64).reshape(1, -1)
    x1 = np.asarray(x1, np.float64).reshape(-1, d.shape[1])
    x2 = np.asarray(x2, np.float64).reshape(-1, d.shape[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        term = np.log(0.25 * np.abs((x1 * x2 * d).sum(axis=1)))
    c = np.zeros(x1.shape[0])
    if cnt1 is not None:
        c = c + np.asarray(cnt1, np.float64)
    if cnt2 is not None:
        c = c + np.asarray(cnt2, np.float64)
    term = term + c * LOG_MIN
    w = np.ones(x1.shape[0]) if wgt is None else np.asarray(wgt, np.float64)
    return float((w * term).sum())
