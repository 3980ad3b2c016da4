"""CPU tests of the callers in front of the hot path: conductance seeding and the initial F,
C++ (product, host side) against the NumPy restatement of codes/bigclam4-7.scala:58-108."""
import numpy as np
import pytest

from conftest import random_graph, tiny_graph


def _seeds(rp, col):
    from bigclam_apachespark_b200 import BigClam
    b = BigClam()
    b.set_graph(rp, col)
    return b, b.conductanceLocalMin(on_gpu=False)


@pytest.mark.parametrize("seed,n,deg,hub", [(1, 80, 4, 10), (2, 300, 6, 40), (3, 500, 3, 0)])
def test_conductance_seeds_match_twin(seed, n, deg, hub):
    from oracle import numpy_twin as T
    rp, col = random_graph(n, deg, seed, hub=hub)
    b, seeds = _seeds(rp, col)
    ranked, cond = T.conductance_local_min(rp, col)
    assert np.allclose(b.conductance, cond, rtol=0, atol=0)
    assert np.array_equal(seeds, ranked)
    # every candidate is the min-id neighbour of some node (or an isolated node)
    mins = {int(col[rp[x]:rp[x + 1]].min()) for x in range(n) if rp[x + 1] > rp[x]}
    iso = {x for x in range(n) if rp[x + 1] == rp[x]}
    assert set(seeds.tolist()) == mins | iso


def test_init_neighbor_com_F(graphs):
    from oracle import numpy_twin as T
    import ctypes as C
    from bigclam_apachespark_b200 import _lib
    rp, col = tiny_graph(graphs)
    n = 12
    ranked, _ = T.conductance_local_min(rp, col)
    lib = _lib.load()
    for K, self_ in [(3, False), (4, True)]:
        F = np.empty((n, K))
        assert lib.bigclam_init_neighbor_com_F(n, rp.ctypes.data, col.ctypes.data, K, ranked.ctypes.data, len(ranked),
                                               1 if self_ else 0, C.c_uint64(7), F.ctypes.data) == 0
        assert np.array_equal(F, T.init_neighbor_com_F(rp, col, K, ranked, include_self=self_))
        assert set(np.unique(F)) <= {0.0, 1.0}
    # more communities than candidates: the extra columns are random 0/1, reproducible from the seed
    K = len(ranked) + 3
    F1, F2 = np.empty((n, K)), np.empty((n, K))
    for F in (F1, F2):
        assert lib.bigclam_init_neighbor_com_F(n, rp.ctypes.data, col.ctypes.data, K, ranked.ctypes.data, len(ranked),
                                               0, C.c_uint64(99), F.ctypes.data) == 0
    assert np.array_equal(F1, F2) and set(np.unique(F1[:, len(ranked):])) == {0.0, 1.0}
    assert np.array_equal(F1[:, :len(ranked)], T.init_neighbor_com_F(rp, col, len(ranked), ranked))


def test_facebook_seeds_and_colsums(graphs):
    """Reference-style F0 on facebook: sumF[c] = degree of seed c (:105-106), F0 is 0/1 and very sparse."""
    import ctypes as C
    from bigclam_apachespark_b200 import BigClam, _lib
    rp, col, _ = graphs.load_npz_graph("facebook_combined")
    n = len(rp) - 1
    b, seeds = _seeds(rp, col)
    assert len(seeds) == len(set(seeds.tolist())) and (np.diff(b.conductance[seeds]) >= 0).all()
    K = 10
    F = np.empty((n, K))
    assert _lib.load().bigclam_init_neighbor_com_F(n, rp.ctypes.data, col.ctypes.data, K, seeds.ctypes.data, len(seeds),
                                                   0, C.c_uint64(1), F.ctypes.data) == 0
    S = np.sort(seeds[:K])
    assert np.array_equal(F.sum(axis=0), np.diff(rp)[S].astype(float))


def test_negative_conductance_on_multigraph_is_still_a_candidate():
    """vol_T = sigma - vol_S - 2 cut goes negative on multigraph input (every neighbour listed several times): as coded
    (:64-67) the conductance is then negative and the node ranks FIRST.  Candidates are tracked by a flag, not by the
    sign of the key."""
    from oracle import numpy_twin as T
    from bigclam_apachespark_b200 import graphs as G
    # a small clique whose edges are all listed three times, plus a path hanging off it
    u = np.array([0, 0, 0, 1, 1, 2] * 3 + [3, 4, 5])
    v = np.array([1, 2, 3, 2, 3, 3] * 3 + [4, 5, 6])
    a = np.concatenate([u, v]); b_ = np.concatenate([v, u])
    order = np.lexsort((b_, a))
    a, b_ = a[order], b_[order]
    n = 7
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(a, minlength=n), out=rp[1:])
    col = b_.astype(np.int32)
    ranked, cond = T.conductance_local_min(rp, col)
    assert (cond < 0).any(), "the construction must produce a negative conductance"
    b, seeds = _seeds(rp, col)
    assert np.array_equal(b.conductance, cond)
    assert np.array_equal(seeds, ranked)
    assert cond[seeds[0]] == cond[ranked].min()
