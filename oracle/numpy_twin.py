"""Independent NumPy restatement of backtrackingLineSearchs (bigclam4-7.scala:152-223).

TEST INFRASTRUCTURE ONLY.  Written separately from bigclam_oracle.c, straight from the Scala,
so that tests can pin one restatement against the other (the reference itself cannot run here
and ships no golden vectors: PARITY UNPINNED).  Vectorised per node, fp64; sums use NumPy's
pairwise order, so agreement with the C oracle is to fp-reassociation noise, not bit-exact.
Only usable for small graphs (pure Python loop over nodes).
"""
from __future__ import annotations

import numpy as np

MIN_P_, MAX_P_, MIN_F_, MAX_F_ = 0.0001, 0.9999, 0.0, 1000.0   # bigclam4-7.scala:40-43


def list_search(beta=0.1, max_inter=15):
    """bigclam4-7.scala:28-33 (prepends, so the Scala list is reversed; order is irrelevant)."""
    step_size = 1.0
    out = [1.0]
    for _ in range(1, max_inter + 1):
        step_size *= beta
        out.insert(0, step_size)
    return out


def step_fn(fu, s, direction):
    """bigclam4-7.scala:110-113."""
    return np.minimum(np.maximum(fu + s * direction, MIN_F_), MAX_F_)


def _edge_terms(x):
    p = np.minimum(np.maximum(np.exp(-x), MIN_P_), MAX_P_)
    return np.log(1.0 - p) + x, p


def loglikelihood(rowptr, col, F, sumF):
    """bigclamv3-7.scala:106-120 / bigclam4-7.scala:196-219."""
    total = 0.0
    for u in range(len(rowptr) - 1):
        fu = F[u]
        nb = col[rowptr[u]:rowptr[u + 1]]
        acc = 0.0
        if len(nb):
            t, _ = _edge_terms(F[nb] @ fu)
            acc = t.sum()
        total += acc - fu @ sumF + fu @ fu
    return total


def backtracking_line_searchs(rowptr, col, F, sumF, alpha=0.05, beta=0.1, max_inter=15,
                              uset=None):
    """Returns (F_new, sumF_new, LLH, accepted_step_per_node)."""
    n, _ = F.shape
    steps = list_search(beta, max_inter)
    F_new = F.copy()
    accepted = np.full(n, np.nan)
    old_sum = np.zeros_like(sumF)
    new_sum = np.zeros_like(sumF)
    any_upd = False
    for u in range(n):
        if uset is not None and u not in uset:
            continue
        nb = col[rowptr[u]:rowptr[u + 1]]
        if len(nb) == 0:
            continue
        fu = F[u]
        FV = F[nb]
        # PRE (:157-169)
        x = FV @ fu
        t, p = _edge_terms(x)
        grad = (FV * (1.0 / (1.0 - p))[:, None]).sum(axis=0) - sumF + fu
        llh_u = t.sum() - fu @ sumF + fu @ fu
        # LS (:172-182): every candidate is evaluated, the max passing one is kept
        best = None
        for s in steps:
            newfu = step_fn(fu, s, grad)
            sfT = sumF - fu + newfu
            xc = FV @ newfu
            tt, _ = _edge_terms(xc)
            result = tt.sum() - newfu @ sfT + newfu @ newfu
            if result >= llh_u + ((alpha * s * grad) @ grad):
                if best is None or s > best:
                    best = s
        if best is not None:
            row = step_fn(fu, best, grad)          # :183
            F_new[u] = row
            accepted[u] = best
            old_sum += fu
            new_sum += row
            any_upd = True
    sumF_new = sumF - (old_sum - new_sum) if any_upd else sumF.copy()   # :192
    return F_new, sumF_new, loglikelihood(rowptr, col, F_new, sumF_new), accepted


# --------------------------------------------------------------------------------------------
# Init of F, restated from codes/bigclam4-7.scala:58-108 (test infrastructure, like the rest of this file)
def conductance_local_min(rowptr, col):
    """conductanceLocalMin() (:58-73).  Returns (ranked candidate ids, conductance per node).
    Deterministic where Spark is not: ties in the ranking break by node id."""
    n = len(rowptr) - 1
    sigma = int(rowptr[-1])                                  # sum of in+out degrees (:60)
    cond = np.zeros(n)
    for x in range(n):
        y = [x] + list(col[rowptr[x]:rowptr[x + 1]])         # getEgoGraphNodes (:54-56)
        ys = set(y)
        z = [i for u_ in y for i in col[rowptr[u_]:rowptr[u_ + 1]]]
        cut_S = sum(1 for i in z if i not in ys)
        vol_S = len(z) - cut_S
        vol_T = sigma - vol_S - cut_S * 2
        cond[x] = 0.0 if vol_S == 0 else (1.0 if vol_T == 0 else cut_S / min(vol_S, vol_T))
    best = {}
    for x in range(n):
        nb = col[rowptr[x]:rowptr[x + 1]]
        key, val = (min((int(v), cond[v]) for v in nb) if len(nb) else (x, 10.0))   # tuple .min (:70)
        best[key] = min(best.get(key, np.inf), val)          # reduceByKey
    ranked = sorted(best, key=lambda v: (best[v], v))        # sortByKey on the conductance
    return np.array(ranked, dtype=np.int32), cond


def init_neighbor_com_F(rowptr, col, K, ranked, include_self=False):
    """initNeighborComF(K) (:81-108) without the random padding (callers pass K <= len(ranked))."""
    n = len(rowptr) - 1
    S = sorted(int(v) for v in ranked[:K])                   # filter over collectNeighbor + zipWithIndex (:85-86)
    F = np.zeros((n, K))
    for c, s in enumerate(S):
        F[col[rowptr[s]:rowptr[s + 1]], c] = 1.0
        if include_self:
            F[s, c] = 1.0
    return F
