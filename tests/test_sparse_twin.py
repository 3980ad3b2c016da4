"""NumPy twin of the sparse-row formulation in csrc/bigclam_sparse.cuh, checked against the oracle on CPU:
the line search evaluates every candidate as  sum_i clamp(fu[idx_i] + s * g[idx_i]) * val_i  over the stored
non-zeros of each neighbour row only, keeps only (ascending idx, value) pairs, and never looks at an inactive
component — the results must be those of the dense restatement."""
import numpy as np

from conftest import random_graph


def _to_sparse(F):
    return [(np.nonzero(r)[0], r[np.nonzero(r)[0]]) for r in F]


def _edge_terms(x, min_p=1e-4, max_p=0.9999):
    p = np.minimum(np.maximum(np.exp(-x), min_p), max_p)
    return np.log(1.0 - p) + x, 1.0 / (1.0 - p)


def sparse_step(rp, col, rows, sumF, steps, alpha=0.05, max_f=1000.0):
    n, K = len(rp) - 1, len(sumF)
    new_rows, accepted, llh_pre = [], np.full(n, -1, dtype=np.int8), 0.0
    D = np.zeros(K)
    for u in range(n):
        idx_u, val_u = rows[u]
        fu_d = np.zeros(K)
        fu_d[idx_u] = val_u
        nb = col[rp[u]:rp[u + 1]]
        x = np.array([rows[v][1] @ fu_d[rows[v][0]] for v in nb])
        t, w = _edge_terms(x) if len(nb) else (np.zeros(0), np.zeros(0))
        llh_u = t.sum() - val_u @ sumF[idx_u] + val_u @ val_u
        llh_pre += llh_u
        if len(nb) == 0:
            new_rows.append(rows[u])
            continue
        g_d = np.zeros(K)
        for v, we in zip(nb, w):                       # neighbour by neighbour, fixed order
            g_d[rows[v][0]] += we * rows[v][1]
        g_d = (g_d - sumF) + fu_d                       # scan: gradient in place (:168)
        G2 = g_d @ g_d
        act = np.nonzero((fu_d > 0) | (g_d > 0))[0]
        jstar = -1
        for j, s in enumerate(steps):                   # all candidates; lowest passing j == largest step
            nf = np.minimum(np.maximum(fu_d + s * g_d, 0.0), max_f)
            assert not nf[np.setdiff1d(np.arange(K), act)].any()        # inactive components clamp to 0
            terms = 0.0
            for v in nb:
                Dv = nf[rows[v][0]] @ rows[v][1]        # only the stored non-zeros of the neighbour row
                terms += _edge_terms(np.array([Dv]))[0][0]
            nfa = nf[act]
            result = (terms - nfa @ ((sumF[act] - fu_d[act]) + nfa)) + nfa @ nfa
            if result >= llh_u + (alpha * s) * G2:
                jstar = j
                break
        accepted[u] = jstar
        if jstar < 0:
            new_rows.append(rows[u])
            continue
        nr = np.minimum(np.maximum(fu_d[act] + steps[jstar] * g_d[act], 0.0), max_f)
        D[act] += fu_d[act] - nr
        keep = nr != 0
        new_rows.append((act[keep], nr[keep]))
    return new_rows, sumF - D, llh_pre, accepted


def test_sparse_formulation_matches_oracle(oracle):
    n, k = 120, 24
    rp, col = random_graph(n, 5, seed=7, hub=40)
    rng = np.random.default_rng(7)
    F = rng.random((n, k)) * (rng.random((n, k)) < 0.2)
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    steps = oracle.step_sizes()
    rows = _to_sparse(F)
    for it in range(3):
        r = oracle.step(rp, col, F, sumF, P)
        rows, sumF_s, llh_pre, acc = sparse_step(rp, col, rows, sumF, steps)
        Fs = np.zeros_like(F)
        for u, (i, v) in enumerate(rows):
            Fs[u, i] = v
            assert (np.diff(i) > 0).all() and (v != 0).all()          # ascending indices, no stored zeros
        same = acc == r.accepted
        assert same.mean() > 0.97
        assert np.allclose(Fs[same], r.F[same], rtol=1e-9, atol=1e-12)
        assert abs(llh_pre - oracle.llh(rp, col, F, sumF, P)) <= 1e-10 * abs(llh_pre)
        if same.all():
            assert np.allclose(sumF_s, r.sumF, rtol=1e-10)
        F, sumF = r.F, r.sumF
        rows = _to_sparse(F)
