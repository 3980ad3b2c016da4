"""Kernel LOGIC on the CPU: the source of csrc/bigclam_sparse.cuh compiled for the host against the SIMT
emulation in tests/emu (one OS thread per CUDA thread, warp collectives as barrier rounds) and checked against
the oracle.  Test infrastructure only — nothing here is reachable from the package, which has no CPU path; the
`-m gpu` tests remain the parity tests of the compiled sm_100a code."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import random_graph, tiny_graph

HERE = os.path.dirname(os.path.abspath(__file__))


HAS_SPARSE_SRC = os.path.exists(os.path.join(HERE, "..", "bigclam_apachespark_b200", "csrc", "bigclam_sparse.cuh"))


# "prefetch": the sparse kernel built with its one-node-ahead prefetch (BIGCLAM_SP_PREFETCH=1, see bigclam_sparse.cuh)
@pytest.fixture(scope="module", params=["default", "prefetch"] if HAS_SPARSE_SRC else ["default"])
def emu(request):
    out = "libemu.so" if request.param == "default" else "libemu_prefetch.so"
    env = dict(os.environ, EMU_OUT=out, EMU_DEFS="" if request.param == "default" else "-DBIGCLAM_SP_PREFETCH=1")
    subprocess.run([os.path.join(HERE, "emu", "build.sh")], check=True, env=env)
    lib = C.CDLL(os.path.join(HERE, "emu", out))
    lib.variant = request.param
    step_args = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                 C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu_dense_step.restype = C.c_int
    lib.emu_dense_step.argtypes = step_args
    lib.emu_dense_step_hubs.restype = C.c_int
    lib.emu_dense_step_hubs.argtypes = step_args
    lib.has_sparse = hasattr(lib, "emu_sparse_step")          # csrc/bigclam_sparse.cuh present in this tree
    if lib.has_sparse:
        lib.emu_sparse_step.restype = C.c_int
        lib.emu_sparse_step.argtypes = step_args + [C.POINTER(C.c_int64)]
        lib.emu_pack_roundtrip.restype = C.c_int64
        lib.emu_pack_roundtrip.argtypes = [C.c_int64, C.c_int32] + [C.c_void_p] * 8
        lib.emu_sparse_step_hubs.restype = C.c_int
        lib.emu_sparse_step_hubs.argtypes = step_args
        lib.emu_sparse_step_ranks.restype = C.c_int
        lib.emu_sparse_step_ranks.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def dense_step(lib, rp, col, F, sumF, mask=None, linesearch=True, grid=1):
    n, k = F.shape
    ld = (k + 3) & ~3
    rp = np.ascontiguousarray(rp, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    F = np.ascontiguousarray(F, dtype=np.float64)
    sumF = np.ascontiguousarray(sumF, dtype=np.float64)
    Fo = np.empty_like(F)
    partials = np.zeros(2 * ld + 2)
    acc = np.empty(n, dtype=np.int8)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    rc = lib.emu_dense_step(n, rp.ctypes.data, col.ctypes.data, k, F.ctypes.data, sumF.ctypes.data,
                            None if m is None else m.ctypes.data, 1 if linesearch else 0, 15, 0.05, 0.1, grid,
                            Fo.ctypes.data, partials.ctypes.data, acc.ctypes.data)
    assert rc == 0
    D, llh_pre, nupd = partials[:k], partials[2 * ld], int(round(partials[2 * ld + 1]))
    return Fo, sumF - D if nupd else sumF.copy(), llh_pre, nupd, acc


def sparse_step(lib, rp, col, F, sumF, mask=None, linesearch=True, grid=1):
    if not lib.has_sparse:
        pytest.skip("no sparse-row kernel in this tree")
    n, k = F.shape
    ld = (k + 3) & ~3
    rp = np.ascontiguousarray(rp, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    F = np.ascontiguousarray(F, dtype=np.float64)
    sumF = np.ascontiguousarray(sumF, dtype=np.float64)
    Fo = np.empty_like(F)
    partials = np.zeros(2 * ld + 2)
    acc = np.empty(n, dtype=np.int8)
    words = C.c_int64(0)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    rc = lib.emu_sparse_step(n, rp.ctypes.data, col.ctypes.data, k, F.ctypes.data, sumF.ctypes.data,
                             None if m is None else m.ctypes.data, 1 if linesearch else 0, 15, 0.05, 0.1, grid,
                             Fo.ctypes.data, partials.ctypes.data, acc.ctypes.data, C.byref(words))
    assert rc == 0
    D, llh_pre, nupd = partials[:k], partials[2 * ld], int(round(partials[2 * ld + 1]))
    return Fo, sumF - D if nupd else sumF.copy(), llh_pre, nupd, acc, words.value


def check(F, s, llh_pre, nupd, acc, r, oracle_llh_pre, max_flips=0):
    scale = max(np.abs(r.F).max(), 1e-300)
    row_err = np.abs(F - r.F).max(axis=1)
    flipped = row_err > 1e-9 * scale
    assert int(flipped.sum()) <= max_flips, (int(flipped.sum()), row_err.max())
    d = (acc != r.accepted) & ~flipped
    assert (((acc[d] < 0) | (acc[d] >= 12)) & ((r.accepted[d] < 0) | (r.accepted[d] >= 12))).all()
    assert abs(llh_pre - oracle_llh_pre) <= 1e-10 * abs(oracle_llh_pre)
    if not flipped.any():
        assert np.allclose(s, r.sumF, rtol=1e-11, atol=1e-9)
        if not d.any():
            assert nupd == r.n_updated


@pytest.mark.timeout(600)
@pytest.mark.parametrize("k,grid", [(5, 1), (12, 2), (40, 1), (200, 1), (300, 2), (1000, 1)])
def test_sparse_kernel_source_against_oracle(emu, oracle, k, grid):
    n = 96
    rp, col = random_graph(n, 5, seed=k, hub=40)
    rng = np.random.default_rng(k)
    F = rng.random((n, k)) * (rng.random((n, k)) < min(1.0, 6.0 / k + 0.05))
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    for it in range(2):
        r = oracle.step(rp, col, F, sumF, P)
        Fo, so, llh_pre, nupd, acc, words = sparse_step(emu, rp, col, F, sumF, grid=grid)
        check(Fo, so, llh_pre, nupd, acc, r, oracle.llh(rp, col, F, sumF, P))
        nnz = (Fo != 0).sum(axis=1)
        assert words == int((((nnz + 3) // 4) * 5).sum())          # the pool holds exactly the padded row blocks
        F, sumF = r.F, r.sumF


@pytest.mark.timeout(600)
def test_sparse_kernel_source_mask_llh_only_and_isolated(emu, oracle, graphs):
    rp, col = tiny_graph(graphs)                                 # nodes 10, 11 have no neighbours
    n, k = len(rp) - 1, 5
    rng = np.random.default_rng(1)
    F = rng.random((n, k)) * (rng.random((n, k)) < 0.6)
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    mask = np.array([1, 0, 1, 1, 0, 1, 1, 0, 1, 1, 1, 0], dtype=np.uint8)
    r = oracle.step(rp, col, F, sumF, P, node_mask=mask)
    Fo, so, llh_pre, nupd, acc, _ = sparse_step(emu, rp, col, F, sumF, mask=mask)
    check(Fo, so, llh_pre, nupd, acc, r, oracle.llh(rp, col, F, sumF, P))
    assert np.array_equal(Fo[mask == 0], F[mask == 0]) and np.array_equal(Fo[10:], F[10:])
    Fo, so, llh_pre, nupd, acc, _ = sparse_step(emu, rp, col, F, sumF, linesearch=False)
    assert np.array_equal(Fo, F) and nupd == 0
    assert abs(llh_pre - oracle.llh(rp, col, F, sumF, P)) <= 1e-10 * abs(llh_pre)


@pytest.mark.timeout(900)
def test_sparse_kernel_source_dense_rows_and_chunking(emu, oracle):
    """Full rows (K = 200 non-zeros: two rows per staged chunk) and a hub whose neighbour list spans chunks."""
    n, k = 60, 200
    rp, col = random_graph(n, 4, seed=9, hub=45)
    rng = np.random.default_rng(9)
    F = rng.random((n, k)) * 0.1
    F[::2] *= (rng.random((n // 2, k)) < 0.05)
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    r = oracle.step(rp, col, F, sumF, P)
    Fo, so, llh_pre, nupd, acc, _ = sparse_step(emu, rp, col, F, sumF)
    check(Fo, so, llh_pre, nupd, acc, r, oracle.llh(rp, col, F, sumF, P), max_flips=1)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("k,grid", [(5, 1), (40, 2), (100, 1), (200, 1), (300, 1), (600, 1)])
def test_dense_kernel_source_against_oracle(emu, oracle, k, grid):
    """The GPU-validated dense kernels through the same emulation: pins the emulation itself (shuffle, ballot and
    barrier semantics) as much as the kernel logic (C2 = 1, 1, 2, 4, 8, 16 chunk shapes; dense and pair-list paths)."""
    if emu.variant != "default":
        pytest.skip("dense kernels do not depend on the sparse build knobs")
    n = 80
    rp, col = random_graph(n, 5, seed=100 + k, hub=40)
    rng = np.random.default_rng(k)
    F = rng.random((n, k)) * (rng.random((n, k)) < min(1.0, 6.0 / k + 0.05))
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    r = oracle.step(rp, col, F, sumF, P)
    Fo, so, llh_pre, nupd, acc = dense_step(emu, rp, col, F, sumF, grid=grid)
    check(Fo, so, llh_pre, nupd, acc, r, oracle.llh(rp, col, F, sumF, P))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 3])
def test_sparse_kernel_source_node_partitioned_pushes(emu, oracle, world):
    """sparse_step_kernel<true>: every rank computes its owned rows and writes them into all replicas' output
    pools (disjoint regions, identical offsets); afterwards every replica holds the oracle's new F."""
    if not emu.has_sparse:
        pytest.skip("no sparse-row kernel in this tree")
    n, k = 90, 16
    rp, col = random_graph(n, 5, seed=21, hub=30)
    rng = np.random.default_rng(21)
    F = rng.random((n, k)) * (rng.random((n, k)) < 0.3)
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    r = oracle.step(rp, col, F, sumF, P)
    ld = (k + 3) & ~3
    Fo = np.empty((world, n, k))
    partials = np.zeros(2 * ld + 2)
    acc = np.empty(n, dtype=np.int8)
    rp64, col32 = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int32)
    rc = emu.emu_sparse_step_ranks(n, rp64.ctypes.data, col32.ctypes.data, k, F.ctypes.data, sumF.ctypes.data, 15, 0.05, 0.1,
                                   world, Fo.ctypes.data, partials.ctypes.data, acc.ctypes.data)
    assert rc == 0
    for w in range(1, world):
        assert np.array_equal(Fo[0], Fo[w])                       # identical replicas
    nupd = int(round(partials[2 * ld + 1]))
    check(Fo[0], sumF - partials[:k] if nupd else sumF, partials[2 * ld], nupd, acc, r, oracle.llh(rp, col, F, sumF, P))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("masked,linesearch", [(False, True), (True, True), (False, False)])
def test_sparse_kernel_source_split_hubs(emu, oracle, masked, linesearch):
    """Hubs of 700 and 300 edges split into 256-edge segments over the warps (phases 1-3 through the global
    scratch), the other nodes on the plain path; also with the larger hub outside the uset and PRE-only."""
    if not emu.has_sparse:
        pytest.skip("no sparse-row kernel in this tree")
    from bigclam_apachespark_b200 import graphs as G
    n, k = 900, 12
    rng = np.random.default_rng(33)
    u = np.concatenate([rng.integers(0, n, 1800), np.zeros(700, dtype=np.int64), np.ones(300, dtype=np.int64)])
    v = np.concatenate([rng.integers(0, n, 1800), rng.choice(np.arange(2, n), 700, replace=False),
                        rng.choice(np.arange(2, n), 300, replace=False)])
    keep = u != v
    lo, hi = np.minimum(u[keep], v[keep]), np.maximum(u[keep], v[keep])
    key = np.unique(lo * n + hi)
    rp, col = G.csr_from_undirected(n, key // n, key % n)
    assert np.diff(rp)[0] >= 700 and np.diff(rp)[1] >= 300
    F = rng.random((n, k)) * (rng.random((n, k)) < 0.3)
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    mask = None
    if masked:
        mask = np.ones(n, dtype=np.uint8)
        mask[0] = 0
    ld = (k + 3) & ~3
    Fo = np.empty_like(F)
    partials = np.zeros(2 * ld + 2)
    acc = np.empty(n, dtype=np.int8)
    rp64, col32 = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int32)
    rc = emu.emu_sparse_step_hubs(n, rp64.ctypes.data, col32.ctypes.data, k, F.ctypes.data, sumF.ctypes.data,
                                  None if mask is None else mask.ctypes.data, 1 if linesearch else 0, 15, 0.05, 0.1, 280,
                                  Fo.ctypes.data, partials.ctypes.data, acc.ctypes.data)
    assert rc == 1002                                           # two split hubs
    llh_pre, nupd = partials[2 * ld], int(round(partials[2 * ld + 1]))
    assert abs(llh_pre - oracle.llh(rp, col, F, sumF, P)) <= 1e-10 * abs(llh_pre)
    if not linesearch:
        assert np.array_equal(Fo, F) and nupd == 0
        return
    r = oracle.step(rp, col, F, sumF, P, node_mask=mask)
    check(Fo, sumF - partials[:k] if nupd else sumF, llh_pre, nupd, acc, r, oracle.llh(rp, col, F, sumF, P))
    if masked:
        assert np.array_equal(Fo[0], F[0])


def test_sparse_host_packer_roundtrip(emu):
    """sp_host_pack (bigclam_set_F_csr) -> device layout -> sparse_to_dense_kernel, and sp_host_unpack
    (bigclam_get_F_csr): unsorted input rows, explicit zeros dropped, empty rows."""
    if not emu.has_sparse:
        pytest.skip("no sparse-row kernel in this tree")
    rng = np.random.default_rng(4)
    n, k = 300, 37
    F = rng.random((n, k)) * (rng.random((n, k)) < 0.15)
    F[5] = 0.0
    F[6] = rng.random(k) + 0.1                                   # a full row
    indptr, indices, values = [0], [], []
    for u in range(n):
        nz = np.nonzero(F[u])[0]
        if u % 7 == 0 and k - len(nz) > 2:                       # sprinkle explicit zeros
            nz = np.concatenate([nz, np.setdiff1d(np.arange(k), nz)[:2]])
        nz = rng.permutation(nz)                                 # unsorted inside the row
        indices.extend(nz.tolist())
        values.extend(F[u, nz].tolist())
        indptr.append(len(indices))
    indptr = np.array(indptr, dtype=np.int64)
    indices = np.array(indices, dtype=np.int32)
    values = np.array(values, dtype=np.float64)
    Fo, cs = np.empty((n, k)), np.empty(k)
    ip2, ix2, vl2 = np.empty(n + 1, dtype=np.int64), np.empty(len(indices), dtype=np.int32), np.empty(len(indices))
    used = emu.emu_pack_roundtrip(n, k, indptr.ctypes.data, indices.ctypes.data, values.ctypes.data, Fo.ctypes.data,
                                  cs.ctypes.data, ip2.ctypes.data, ix2.ctypes.data, vl2.ctypes.data)
    nnz_rows = (F != 0).sum(axis=1)
    assert used == int((((nnz_rows + 3) // 4) * 5).sum())
    assert np.array_equal(Fo, F)
    assert np.allclose(cs, F.sum(axis=0), rtol=1e-13)
    assert ip2[-1] == int(nnz_rows.sum()) and np.array_equal(np.diff(ip2), nnz_rows)
    G = np.zeros_like(F)
    for u in range(n):
        sl = slice(ip2[u], ip2[u + 1])
        assert (np.diff(ix2[sl]) > 0).all()
        G[u, ix2[sl]] = vl2[sl]
    assert np.array_equal(G, F)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("k", [12, 200])
def test_dense_kernel_source_hub_phase(emu, oracle, k):
    """step_kernel<C2,R,true,false>: a 700-edge hub as a multi-phase "mega" hub (two 384-edge slices through the
    global scratch), a 300-edge hub done by the whole block (phase 0), the other nodes one warp each."""
    if emu.variant != "default":
        pytest.skip("dense kernels do not depend on the sparse build knobs")
    from bigclam_apachespark_b200 import graphs as G
    n = 900
    rng = np.random.default_rng(44)
    u = np.concatenate([rng.integers(0, n, 1500), np.zeros(700, dtype=np.int64), np.ones(300, dtype=np.int64)])
    v = np.concatenate([rng.integers(0, n, 1500), rng.choice(np.arange(2, n), 700, replace=False),
                        rng.choice(np.arange(2, n), 300, replace=False)])
    keep = u != v
    lo, hi = np.minimum(u[keep], v[keep]), np.maximum(u[keep], v[keep])
    key = np.unique(lo * n + hi)
    rp, col = G.csr_from_undirected(n, key // n, key % n)
    F = rng.random((n, k)) * (rng.random((n, k)) < min(1.0, 4.0 / k + 0.05))
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    ld = (k + 3) & ~3
    Fo = np.empty_like(F)
    partials = np.zeros(2 * ld + 2)
    acc = np.empty(n, dtype=np.int8)
    rp64, col32 = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int32)
    rc = emu.emu_dense_step_hubs(n, rp64.ctypes.data, col32.ctypes.data, k, F.ctypes.data, sumF.ctypes.data, None, 1, 15,
                                 0.05, 0.1, 280, Fo.ctypes.data, partials.ctypes.data, acc.ctypes.data)
    assert rc == 1002
    r = oracle.step(rp, col, F, sumF, P)
    nupd = int(round(partials[2 * ld + 1]))
    check(Fo, sumF - partials[:k] if nupd else sumF, partials[2 * ld], nupd, acc, r, oracle.llh(rp, col, F, sumF, P), max_flips=1)
