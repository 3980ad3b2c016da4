"""GPU parity tests of the sparse-row mode (BIGCLAM_F_SPARSE_ROWS, csrc/bigclam_sparse.cuh): same checks as the
dense path (tests/test_gpu_parity.py), same oracle, same tolerances — the C ABI stays dense."""
import numpy as np
import pytest

from conftest import random_graph, tiny_graph
from test_gpu_parity import RTOL_TIGHT, _check_step

pytestmark = pytest.mark.gpu


def _solver(rp, col, K, F0, sumF=None, **kw):
    from bigclam_apachespark_b200 import BigClam
    b = BigClam(record_accepted=True, sparse_rows=True, **kw)
    b.set_graph(rp, col).set_K(K).set_F(F0, sumF=sumF)
    return b


@pytest.mark.parametrize("k", [1, 2, 3, 5, 10, 31, 64, 65, 100, 200, 256, 257, 500, 1000])
def test_sparse_single_step_all_k(oracle, k):
    n = 400
    rp, col = random_graph(n, 6, seed=k, hub=60)
    rng = np.random.default_rng(k)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < min(1.0, 8.0 / k + 0.05))
    sumF = oracle.colsum(F0)
    b = _solver(rp, col, k, F0, sumF)
    assert np.array_equal(b.F, F0)                      # dense -> sparse -> dense round trip
    llh = b.backtrackingLineSearchs()
    r = oracle.step(rp, col, F0, sumF, oracle.make_params(k))
    _check_step(b, r, llh, where=f"sparse k={k}", inputs=(rp, col, F0, sumF, oracle.make_params(k)))
    b.close()


def test_sparse_dense_rows_and_long_neighbour_lists(oracle):
    """Rows with every component non-zero (entry buffer holds only 2 rows per chunk) and a hub of 300 edges."""
    n, k = 600, 200
    rp, col = random_graph(n, 8, seed=3, hub=300)
    rng = np.random.default_rng(3)
    F0 = rng.random((n, k)) * 0.2
    F0[::3] *= (rng.random((len(F0[::3]), k)) < 0.05)
    sumF = oracle.colsum(F0)
    b = _solver(rp, col, k, F0, sumF)
    F, s = F0, sumF
    for it in range(3):
        llh = b.backtrackingLineSearchs()
        r = oracle.step(rp, col, F, s, oracle.make_params(k))
        _check_step(b, r, llh, max_flips=2, where=f"sparse dense-rows it{it}", inputs=(rp, col, F, s, oracle.make_params(k)))
        F, s = b.F, b.sumF
    b.close()


def test_sparse_golden_tiny_and_isolated_nodes(oracle, golden, graphs):
    rp, col = tiny_graph(graphs)
    F = golden["tiny_F0"]
    b = _solver(rp, col, 5, F, oracle.colsum(F))
    for it in range(3):
        llh = b.backtrackingLineSearchs()
        assert abs(llh - golden[f"tiny_llh_{it}"]) <= 1e-10 * abs(golden[f"tiny_llh_{it}"])
        assert np.allclose(b.F, golden[f"tiny_F_{it}"], rtol=RTOL_TIGHT, atol=1e-12)
        assert np.array_equal(b.F[10:], F[10:])          # empty neighbour lists: rows never change
    b.close()


def test_sparse_uset_mask_and_loglikelihood(oracle):
    n, k = 500, 12
    rp, col = random_graph(n, 5, seed=11)
    rng = np.random.default_rng(11)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.4)
    sumF = oracle.colsum(F0)
    b = _solver(rp, col, k, F0, sumF)
    assert abs(b.loglikelihood() - oracle.llh(rp, col, F0, sumF, oracle.make_params(k))) <= 1e-10 * abs(b.loglikelihood())
    mask = (rng.random(n) < 0.5).astype(np.uint8)
    llh = b.backtrackingLineSearchs(np.nonzero(mask)[0])
    r = oracle.step(rp, col, F0, sumF, oracle.make_params(k), node_mask=mask)
    _check_step(b, r, llh, where="sparse uset", inputs=(rp, col, F0, sumF, oracle.make_params(k)))
    assert np.array_equal(b.F[mask == 0], F0[mask == 0])
    b.close()


@pytest.mark.parametrize("variant", [2, 3, 4])
def test_sparse_run_loop_matches_dense(oracle, variant):
    """The device-side convergence loop over sparse rows ends where the dense kernels end."""
    from bigclam_apachespark_b200 import BigClam
    n, k = 800, 20
    rp, col = random_graph(n, 6, seed=5)
    rng = np.random.default_rng(5)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.3)
    outs = []
    for sparse in (False, True):
        b = BigClam(sparse_rows=sparse)
        b.set_graph(rp, col).set_K(k).set_F(F0)
        llh = b._run(variant, 1e-4, 200)
        outs.append((llh, b.last_calls, b.F, b.sumF))
        b.close()
    assert outs[0][1] == outs[1][1]
    assert abs(outs[0][0] - outs[1][0]) <= 1e-9 * abs(outs[0][0])
    assert np.allclose(outs[0][2], outs[1][2], rtol=1e-7, atol=1e-9)
    assert np.allclose(outs[0][3], outs[1][3], rtol=1e-9)


def test_sparse_com_amazon_k200(oracle, graphs):
    rp, col, _ = graphs.load_npz_graph("com-amazon")
    n, K = len(rp) - 1, 200
    F0 = graphs.synthetic_F0(n, K, seed=1234, density=0.05)
    sumF = oracle.colsum(F0)
    P = oracle.make_params(K)
    b = _solver(rp, col, K, F0, sumF)
    F, s = F0, sumF
    for it in range(2):
        llh = b.backtrackingLineSearchs()
        r = oracle.step(rp, col, F, s, P)
        _check_step(b, r, llh, max_flips=5, where=f"sparse amazon it{it}", inputs=(rp, col, F, s, P))
        F, s = b.F, b.sumF
    b._run(4, 0.0, 10)
    Fg, sg = b.F, b.sumF
    assert np.abs(sg - Fg.sum(axis=0)).max() <= 1e-9 * np.abs(sg).max()
    assert abs(b.loglikelihood() - b.last_trace[-1]) <= 1e-12 * abs(b.last_trace[-1])
    b.close()


def test_sparse_mode_limits():
    from bigclam_apachespark_b200 import BigClam, _lib
    rp = np.array([0, 1, 2], dtype=np.int64)
    col = np.array([1, 0], dtype=np.int32)
    b = BigClam(sparse_rows=True)
    b.MIN_F_ = 0.5                                       # the sparse layout presumes MIN_F_ == 0 (zeros are not stored)
    b.set_graph(rp, col)
    with pytest.raises(_lib.BigclamError):
        b.set_K(4)


def test_sparse_split_hubs(oracle, monkeypatch):
    """Hubs split into 256-edge segments over warps (phases 1-3 through the global scratch); the threshold is
    lowered through the test knob BIGCLAM_SPARSE_HUB_DEG so that a small graph has split hubs."""
    monkeypatch.setenv("BIGCLAM_SPARSE_HUB_DEG", "200")
    n, k = 4000, 40
    rp, col = random_graph(n, 6, seed=17, hub=1500)
    assert np.diff(rp).max() >= 1500
    rng = np.random.default_rng(17)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.2)
    sumF = oracle.colsum(F0)
    b = _solver(rp, col, k, F0, sumF)
    F, s = F0, sumF
    for it in range(3):
        llh = b.backtrackingLineSearchs()
        r = oracle.step(rp, col, F, s, oracle.make_params(k))
        _check_step(b, r, llh, max_flips=2, where=f"sparse hubs it{it}", inputs=(rp, col, F, s, oracle.make_params(k)))
        F, s = b.F, b.sumF
    b.close()


def test_sparse_rmat_skewed_graph(oracle, graphs):
    rp, col = graphs.rmat_graph(20000, 200000, seed=42)
    n, k = len(rp) - 1, 32
    rng = np.random.default_rng(2)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.15)
    sumF = oracle.colsum(F0)
    b = _solver(rp, col, k, F0, sumF)
    F, s = F0, sumF
    for it in range(2):
        llh = b.backtrackingLineSearchs()
        r = oracle.step(rp, col, F, s, oracle.make_params(k))
        _check_step(b, r, llh, max_flips=3, where=f"sparse rmat it{it}", inputs=(rp, col, F, s, oracle.make_params(k)))
        F, s = b.F, b.sumF
    b.close()


@pytest.mark.parametrize("sparse", [True, False])
def test_csr_entry_points(oracle, sparse):
    """bigclam_set_F_csr / bigclam_get_F_nnz / bigclam_get_F_csr: the reference's RDD[(Long, BSV[Double])] shape."""
    import scipy.sparse as sps
    from bigclam_apachespark_b200 import BigClam
    n, k = 700, 50
    rp, col = random_graph(n, 6, seed=8)
    rng = np.random.default_rng(8)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.1)
    b = BigClam(record_accepted=True, sparse_rows=sparse)
    b.set_graph(rp, col).set_K(k)
    b.set_F(sps.csr_matrix(F0))
    assert np.array_equal(b.F, F0)
    assert np.allclose(b.sumF, F0.sum(axis=0), rtol=1e-13)
    llh = b.backtrackingLineSearchs()
    r = oracle.step(rp, col, F0, oracle.colsum(F0), oracle.make_params(k))
    _check_step(b, r, llh, where="csr", inputs=(rp, col, F0, oracle.colsum(F0), oracle.make_params(k)))
    ip, ix, vl = b.F_csr()
    F1 = b.F
    assert ip[-1] == (F1 != 0).sum()
    G = np.zeros_like(F1)
    for u in range(n):
        G[u, ix[ip[u]:ip[u + 1]]] = vl[ip[u]:ip[u + 1]]
    assert np.array_equal(G, F1)
    b.close()


def test_sparse_pool_exhaustion_is_reported(monkeypatch):
    from bigclam_apachespark_b200 import BigClam, _lib
    monkeypatch.setenv("BIGCLAM_SPARSE_POOL_WORDS", "2000")
    n, k = 600, 40
    rp, col = random_graph(n, 6, seed=2)
    rng = np.random.default_rng(2)
    b = BigClam(sparse_rows=True)
    b.set_graph(rp, col).set_K(k)
    with pytest.raises(_lib.BigclamError):
        b.set_F(rng.random((n, k)))                      # 600 full rows do not fit 2000 words
    b.close()


def test_sparse_reruns_are_bit_identical(oracle):
    """No floating-point atomics on the sparse path: per-node results + fixed-order reduction.  Two fresh contexts give
    the same bits (LLH trace, F, sumF) — with dynamic work distribution, split hubs and tiles that fall back."""
    rp, col = random_graph(3000, 7, seed=17, hub=400)
    k = 60
    rng = np.random.default_rng(17)
    F0 = rng.random((3000, k)) * (rng.random((3000, k)) < 0.15)
    sumF = oracle.colsum(F0)
    out = []
    for _ in range(2):
        b = _solver(rp, col, k, F0, sumF)
        llh = b._run(4, 0.0, 7)
        out.append((llh, b.last_trace.tobytes(), b.F.tobytes(), b.sumF.tobytes()))
        b.close()
    assert out[0] == out[1]


def test_sparse_tiles_are_recut_when_rows_fill_up(oracle):
    """Small K: the rows fill up while the solver runs and the tiles cut for the sparse F0 start to fall back; between its
    batches bigclam_run re-cuts them (smaller tiles / tiles off) — results unchanged."""
    n, k = 4000, 24
    rp, col = random_graph(n, 10, seed=23)
    rng = np.random.default_rng(23)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.1)
    sumF = oracle.colsum(F0)
    b = _solver(rp, col, k, F0, sumF, time_kernels=True)
    t0 = b.tile_stats()
    llh = b._run(4, 0.0, 24)
    t1 = b.tile_stats()
    Fo, so, llh_o, calls_o, _ = oracle.run(rp, col, F0, sumF, oracle.make_params(k), variant=4, rel_tol=0.0, max_outer=24)
    assert b.last_calls == calls_o == 24
    assert abs(llh - llh_o) <= 1e-9 * abs(llh_o)
    assert np.abs(b.F - Fo).max() <= 1e-7 * max(np.abs(Fo).max(), 1e-300)
    assert t0["n_tiles"] > 0
    print("tiles before / after:", t0, t1)
    b.close()
