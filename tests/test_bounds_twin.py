"""The MATH of the line search by bounds (csrc/bigclam_tile.cuh, phase H2) restated in NumPy and checked against the oracle:
whatever candidate the oracle's exhaustive line search accepts must survive the bound (soundness), on graphs and F where
the bound has bite (most candidates excluded) and where it has none.  Independent of the CUDA code: a wrong derivation
shows up here on the CPU, a wrong implementation in tests/test_gpu_prune.py (bounds vs exhaustive, bit for bit)."""
import numpy as np
import pytest
import scipy.sparse as sps

from conftest import random_graph


def _survivors(rp, col, F, sumF, alpha=0.05, min_p=1e-4, max_p=0.9999, max_f=1000.0, nsteps=16, beta=0.1):
    """Boolean (n, nsteps): candidates the bound cannot exclude (twin of TlWarp::run H2 / LsBound::cannot_pass, in fp64,
    with the same allowances)."""
    n = len(rp) - 1
    steps = [1.0]
    for _ in range(nsteps - 1):
        steps.append(steps[-1] * beta)
    deg = np.diff(rp)
    rows = np.repeat(np.arange(n), deg)
    x_lo, x_hi = -np.log(max_p), -np.log(min_p)
    s_lo, s_hi = np.log(1.0 - max_p), np.log(1.0 - min_p)
    m_lo = 1.0 / (1.0 - max_p) - 1.0
    x = np.einsum("ij,ij->i", F[rows], F[col]) if len(col) else np.zeros(0)
    p = np.clip(np.exp(-x), min_p, max_p)
    w = 1.0 / (1.0 - p)
    S = np.log(1.0 - p)
    m = w - 1.0
    low, high = x <= x_lo, x >= x_hi
    near = (~low) & (~high)
    viol = np.maximum(0.0, s_lo - (S - m * x))                 # tangent of an in-range edge, cut off at x_lo
    g = sps.csr_matrix((w, col, rp), shape=(n, n)) @ F - sumF[None, :] + F
    act = (F > 0) | (g > 0)
    ga = g * act
    gp, gn = np.maximum(ga, 0.0), np.minimum(ga, 0.0) * (F > 0)
    Dp = np.einsum("ij,ij->i", gp[rows], F[col])
    En = np.einsum("ij,ij->i", gn[rows], F[col])
    sDp = np.bincount(rows, weights=Dp * low, minlength=n)
    sEn = np.bincount(rows, weights=En * low, minlength=n)
    Vhigh = np.bincount(rows, weights=np.where(high, m * (x - x_hi), 0.0), minlength=n)
    Qn = (np.minimum(ga, 0.0) ** 2).sum(1) - m_lo * sEn
    Qp = (gp ** 2).sum(1)
    Mp = m_lo * sDp
    G1 = max_f * gp.sum(1)
    t = 2 * np.abs(ga) + np.abs(sumF)[None, :] + 3 * F
    r3 = (F * t * act).sum(1)
    r4 = (np.abs(ga) * t).sum(1)
    gmax = gp.max(1) if F.shape[1] else np.zeros(n)
    kap0 = np.where(gmax > 0, np.maximum(max_f - F.max(1), 0.0) / np.maximum(gmax, 1e-300), np.inf)
    fusf, fufu = F @ sumF, (F * F).sum(1)
    llh = np.bincount(rows, weights=S + x, minlength=n) - fusf + fufu
    G2 = (g * g).sum(1)
    nops = 2.3e-16 * (4 * deg + 3 * act.sum(1) + 16)
    cap = max(-s_lo, s_hi - s_lo)
    base = 2 * cap * deg + np.abs(llh) + 2 * (np.abs(fusf) + fufu) + r3
    c0, c1 = Vhigh + nops * base, nops * (sDp + 2 * r4)
    surv = np.zeros((n, nsteps), bool)
    for j, s in enumerate(steps):
        y = x + s * Dp
        H = np.where(low & (y > x_lo), np.minimum(np.log(np.maximum(y, 1e-300) / (1.0 - max_p)), cap), 0.0)
        H = H + np.where(near & (viol > 0) & (x + s * En < x_lo), viol, 0.0)
        Hs = np.bincount(rows, weights=H, minlength=n)
        kap = np.minimum(1.0, kap0 / s)
        bound = np.minimum(s * Qn, r3) + np.minimum(s * Qp, G1) - s * kap * Mp + c0 + s * c1 + Hs
        surv[:, j] = ~(bound < alpha * s * G2 * (1 - 1e-9))
    surv[deg == 0] = False
    return surv


def _check(oracle, rp, col, F, sumF, k, steps, where):
    P = oracle.make_params(k)
    excluded = total = 0
    for it in range(steps):
        r = oracle.step(rp, col, F, sumF, P)
        acc = np.asarray(r.accepted)
        surv = _survivors(rp, col, F, sumF)
        hit = acc >= 0
        wrong = hit & ~surv[np.arange(len(acc)), np.maximum(acc, 0)]
        assert not wrong.any(), f"{where} step {it}: the bound excludes accepted candidates of nodes {np.nonzero(wrong)[0][:8]}"
        excluded += int((~surv.any(1) & (np.diff(rp) > 0)).sum())
        total += int((np.diff(rp) > 0).sum())
        F, sumF = r.F, r.sumF
    return excluded, total


@pytest.mark.parametrize("k,n,deg,dens,big", [(10, 300, 5, 0.3, 1), (64, 400, 4, 0.1, 900), (200, 500, 5, 0.05, 700), (24, 300, 30, 0.2, 50)])
def test_bound_never_excludes_an_accepted_candidate(oracle, k, n, deg, dens, big):
    rp, col = random_graph(n, deg, seed=100 + k, hub=60 if deg > 20 else 0)
    rng = np.random.default_rng(k)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < dens)
    excluded, total = _check(oracle, rp, col, F0, oracle.colsum(F0) * big, k, 6, f"k={k} big={big}")
    if big >= 700:
        assert excluded > 0.1 * total, f"the bound has no bite here: {excluded} of {total} node-steps excluded"


def test_bound_with_values_next_to_the_clamps(oracle):
    n, k = 240, 24
    rp, col = random_graph(n, 6, seed=7)
    rng = np.random.default_rng(7)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.25)
    F0[:60] *= 0.012          # products around x_lo = 1e-4
    F0[60:120] *= 6.0         # products above x_hi = 9.2
    F0[120:130] = np.where(F0[120:130] > 0, 999.5, 0.0)
    _check(oracle, rp, col, F0, oracle.colsum(F0) * 200, k, 5, "clamps")


def test_bound_on_facebook(oracle, graphs):
    rp, col, _ = graphs.load_npz_graph("facebook_combined")
    n = len(rp) - 1
    F0 = graphs.synthetic_F0(n, 10, seed=1234, density=0.2)
    _check(oracle, rp, col, F0, oracle.colsum(F0) * 80, 10, 4, "facebook")
