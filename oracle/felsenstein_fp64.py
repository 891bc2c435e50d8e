"""oracle/felsenstein_fp64.py -- textbook Felsenstein pruning in PROBABILITY space, float64.  TEST INFRASTRUCTURE ONLY.

An INDEPENDENT model of the quantity the newview -> scaler counts -> evaluate chain computes.  Nothing here
shares a line, a layout or a trick with the product or with the other oracles:

  * the substitution model is a real GTR + discrete-Gamma(4) model (Yang 1994 mean-of-category rates):
    Q from exchangeabilities and base frequencies, normalised to one expected substitution per unit time;
  * P(t) = U diag(exp(lambda r t)) U^-1 is formed EXPLICITLY as a 4x4 transition matrix per (branch, rate
    category) from the eigen-decomposition of the symmetrised Q (scipy/numpy eigh, float64);
  * conditional likelihoods live in state space (A, C, G, T), L_parent[a] = (sum_b P_l[a,b] L_l[b]) *
    (sum_b P_r[a,b] L_r[b])  (Felsenstein 1981), NOT in the eigen-space the reference's plf() works in
    (app/src/plf.cpp:29-50 takes eigen-space CLVs, branch matrices with the eigenvectors folded in, and EV);
  * no 2^32 threshold rescaling: every node renormalises each site by its largest entry and carries the
    logarithm of that factor (float64), the standard log-space bookkeeping;
  * lnL = sum_i w_i log( sum_j 1/4 sum_a pi_a L_root[i,j,a] ).

`eigen_inputs()` hands the SAME model to the reference-style path: tips as eigen-space vectors U^-1 e_s, per-node
branch matrices left[j][k][l] = U[k][l] exp(lambda_l r_j t) in the reference's [j][k][l] layout, EV[k][l] =
U^-1[l][k], and for the branch across the root diag[j][k] = exp(lambda_k r_j t_root).  If the path's EV/P
conventions, its scaler counts or its log(2^-32) bookkeeping drift, its lnL leaves this one.

`mp_log_likelihood()` repeats the pruning with mpmath (50 digits, no renormalisation at all) for tiny cases: the
check of this checker.
"""
from __future__ import annotations

import numpy as np

STATES = "ACGT"


class GtrGamma:
    def __init__(self, rates=(1.2, 3.1, 0.7, 0.9, 4.0, 1.0), freqs=(0.31, 0.19, 0.24, 0.26), alpha=0.7, ncat=4):
        pi = np.asarray(freqs, np.float64)
        self.pi = pi / pi.sum()
        r = np.asarray(rates, np.float64)                     # AC AG AT CG CT GT
        s = np.zeros((4, 4))
        s[np.triu_indices(4, 1)] = r
        s = s + s.T
        q = s * self.pi[None, :]
        np.fill_diagonal(q, 0.0)
        np.fill_diagonal(q, -q.sum(axis=1))
        q /= -(self.pi * np.diag(q)).sum()                    # one expected substitution per unit time
        self.q = q
        sq = np.sqrt(self.pi)
        b = (sq[:, None] * q) / sq[None, :]                   # symmetric for a reversible model
        lam, v = np.linalg.eigh((b + b.T) / 2)
        self.lam = lam
        self.u = v / sq[:, None]                              # Q = U diag(lam) U^-1
        self.u_inv = v.T * sq[None, :]
        self.cat_rates = gamma_category_rates(alpha, ncat)
        self.ncat = ncat

    def p_matrix(self, t: float, rate: float) -> np.ndarray:
        """Transition matrix P[a, b] = Pr(b at the child | a at the parent) over branch length t."""
        return (self.u * np.exp(self.lam * rate * t)[None, :]) @ self.u_inv


def gamma_category_rates(alpha: float, ncat: int) -> np.ndarray:
    """Mean rate of each of ncat equal-probability categories of a Gamma(alpha, 1/alpha) distribution."""
    from scipy.stats import gamma
    cuts = gamma.ppf(np.linspace(0, 1, ncat + 1), alpha, scale=1.0 / alpha)
    upper = gamma.cdf(cuts, alpha + 1, scale=1.0 / alpha)     # E[X; X < c] = cdf_{alpha+1}(c) for mean 1
    return (upper[1:] - upper[:-1]) * ncat


def tip_likelihoods(codes: np.ndarray) -> np.ndarray:
    """codes[n] are 4-bit ambiguity masks over (A=1, C=2, G=4, T=8); 0 is treated as 15 (gap).  -> [n, 4] of 0/1."""
    c = np.asarray(codes, np.int64)
    c = np.where(c == 0, 15, c)
    return ((c[:, None] >> np.arange(4)[None, :]) & 1).astype(np.float64)


def log_likelihood(model: GtrGamma, left, right, t_left, t_right, tip_codes, wgt=None) -> float:
    """Tree in post-order (ids 0..n_tips-1 tips, n_tips+k inner node k with children left[k], right[k] over branches
    t_left[k], t_right[k]); the last inner node is the root.  tip_codes[n_tips][n_sites]."""
    n_tips, n = tip_codes.shape
    cl = {}      # node -> (L [n, ncat, 4] normalised so that the per-site max is 1, logscale [n])
    for i in range(n_tips):
        tl = tip_likelihoods(tip_codes[i])
        cl[i] = (np.repeat(tl[:, None, :], model.ncat, axis=1), np.zeros(n))
    for k, (a, b) in enumerate(zip(left, right)):
        (la, sa), (lb, sb) = cl.pop(int(a)), cl.pop(int(b))
        out = np.empty((n, model.ncat, 4))
        for j, r in enumerate(model.cat_rates):
            pa, pb = model.p_matrix(float(t_left[k]), r), model.p_matrix(float(t_right[k]), r)
            out[:, j, :] = (la[:, j, :] @ pa.T) * (lb[:, j, :] @ pb.T)
        m = out.reshape(n, -1).max(axis=1)
        cl[n_tips + k] = (out / m[:, None, None], sa + sb + np.log(m))
    root, scale = cl[n_tips + len(left) - 1]
    site = (root * model.pi[None, None, :]).sum(axis=2).mean(axis=1)
    w = np.ones(n) if wgt is None else np.asarray(wgt, np.float64)
    return float((w * (np.log(site) + scale)).sum())


def mp_log_likelihood(model: GtrGamma, left, right, t_left, t_right, tip_codes, wgt=None, digits: int = 50) -> float:
    """The same pruning in mpmath with `digits` significant digits and NO renormalisation anywhere."""
    import mpmath as mp
    mp.mp.dps = digits
    n_tips, n = tip_codes.shape
    u = mp.matrix(model.u.tolist())
    ui = mp.matrix(model.u_inv.tolist())

    def pmat(t, r):
        d = mp.diag([mp.e ** (mp.mpf(float(l)) * mp.mpf(float(r)) * mp.mpf(float(t))) for l in model.lam])
        return u * d * ui

    total = mp.mpf(0)
    pms = [[(pmat(t_left[k], r), pmat(t_right[k], r)) for r in model.cat_rates] for k in range(len(left))]
    for i in range(n):
        site = mp.mpf(0)
        for j in range(model.ncat):
            vec = {t: mp.matrix(tip_likelihoods(tip_codes[t, i:i + 1])[0].tolist()) for t in range(n_tips)}
            for k, (a, b) in enumerate(zip(left, right)):
                pa, pb = pms[k][j]
                va, vb = pa * vec.pop(int(a)), pb * vec.pop(int(b))
                vec[n_tips + k] = mp.matrix([va[s] * vb[s] for s in range(4)])
            root = vec[n_tips + len(left) - 1]
            site += sum(mp.mpf(float(model.pi[s])) * root[s] for s in range(4)) / model.ncat
        total += (1 if wgt is None else int(wgt[i])) * mp.log(site)
    return float(total)


def eigen_inputs(model: GtrGamma, left, right, t_left, t_right):
    """The same model in the reference's operand format (float32): EV[16], P_left / P_right [n_inner][64] in the
    [j][k][l] layout of plf.cpp:37-38, the 16 x 4 tip-vector table (eigen-space image of every ambiguity code), and
    diag[16] for the branch of length t_left[-1] + t_right[-1] that joins the two children of the last inner node."""
    n_inner = len(left)
    ev = model.u_inv.T.astype(np.float32).reshape(16)                         # EV[k][l] = U^-1[l][k]
    pl = np.empty((n_inner, 4, 4, 4), np.float64)
    pr = np.empty((n_inner, 4, 4, 4), np.float64)
    for k in range(n_inner):
        for j, r in enumerate(model.cat_rates):
            pl[k, j] = model.u * np.exp(model.lam * r * float(t_left[k]))[None, :]    # [k][l] = U[k][l] e^{lam_l r t}
            pr[k, j] = model.u * np.exp(model.lam * r * float(t_right[k]))[None, :]
    tip_vector = np.stack([model.u_inv @ tip_likelihoods(np.array([c]))[0] for c in range(16)])   # [code][l]
    t_root = float(t_left[-1]) + float(t_right[-1])
    diag = np.stack([np.exp(model.lam * r * t_root) for r in model.cat_rates])                      # [j][k]
    return (ev, pl.reshape(n_inner, 64).astype(np.float32), pr.reshape(n_inner, 64).astype(np.float32),
            tip_vector.astype(np.float32), diag.astype(np.float32).reshape(16))


def random_branch_lengths(n_inner: int, seed: int, lo: float = 0.02, hi: float = 0.6):
    rng = np.random.RandomState(seed)
    return rng.uniform(lo, hi, n_inner), rng.uniform(lo, hi, n_inner)


def simulate_alignment(model: GtrGamma, left, right, t_left, t_right, n_sites: int, seed: int, ambiguity: float = 0.05):
    """Sequences evolved down the tree under the model (one Gamma category per site), as 4-bit codes; a fraction
    `ambiguity` of the tip characters is replaced by a random ambiguity code (including the all-ones gap)."""
    rng = np.random.RandomState(seed)
    n_tips = len(left) + 1
    cat = rng.randint(0, model.ncat, n_sites)
    state = {n_tips + len(left) - 1: rng.choice(4, n_sites, p=model.pi)}
    for k in range(len(left) - 1, -1, -1):
        parent = state.pop(n_tips + k)
        for child, t in ((int(left[k]), t_left[k]), (int(right[k]), t_right[k])):
            out = np.empty(n_sites, np.int64)
            for j, r in enumerate(model.cat_rates):
                p = np.clip(model.p_matrix(float(t), r), 0, None)
                p /= p.sum(axis=1, keepdims=True)
                sel = np.nonzero(cat == j)[0]
                cum = p[parent[sel]].cumsum(axis=1)
                out[sel] = (rng.random_sample(sel.size)[:, None] > cum).sum(axis=1).clip(0, 3)
            state[child] = out
    codes = np.stack([1 << state[i] for i in range(n_tips)]).astype(np.uint8)
    amb = rng.random_sample(codes.shape) < ambiguity
    codes[amb] = rng.randint(1, 16, int(amb.sum())).astype(np.uint8)
    return codes
