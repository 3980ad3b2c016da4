"""Two ranks of the node-partitioned path in ONE process, against the host-emulation build of the C API only
(BIGCLAM_HOSTEMU=1, tests/emu/build_hostemu.sh): there the "IPC handles" are plain pointers, so the peers' replicas
can be opened inside the process and the C-ABI sequence of bigclam_apachespark_b200/dist.py (set_owned_nodes,
set_pool_region, ipc_export / ipc_open_peers, step_local, all-reduce of the partials, finish_local) can be driven
without torch or a GPU.  Skipped everywhere else (a real cudaIpcOpenMemHandle refuses the exporting process)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import random_graph

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("BIGCLAM_HOSTEMU") != "1", reason="host-emulation build only")]


@pytest.mark.parametrize("hub", [40, 150])          # 150: the owner of the hub runs the hub phase (dense) / split hubs (sparse)
@pytest.mark.parametrize("sparse", [False, True])
def test_two_ranks_through_the_c_abi(oracle, sparse, hub, monkeypatch):
    monkeypatch.setenv("BIGCLAM_SPARSE_HUB_DEG", "100")
    from bigclam_apachespark_b200 import BigClam, _lib
    from bigclam_apachespark_b200.dist import deal_all_by_degree
    lib = _lib.load()
    world, n, k = 2, 160, 12
    rp, col = random_graph(n, 5, seed=31, hub=hub)
    rng = np.random.default_rng(31)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.3)
    sumF = oracle.colsum(F0)
    ld = (k + 3) & ~3
    ranks = []
    for r in range(world):
        b = BigClam(record_accepted=True, sparse_rows=sparse)
        b.set_graph(rp, col).set_K(k).set_F(F0, sumF=sumF)
        deal = deal_all_by_degree(rp, world)
        nodes = deal[r]
        _lib.check(lib.bigclam_set_owned_nodes(b._ctx, nodes.ctypes.data, len(nodes)), b._ctx)
        if sparse:
            counts = [len(x) for x in deal]
            row_words = _lib.sparse_node_words(ld)
            _lib.check(lib.bigclam_set_pool_region(b._ctx, sum(counts[:r]) * row_words, counts[r] * row_words), b._ctx)
        ranks.append(b)
    per = 64 * lib.bigclam_ipc_handle_count(ranks[0]._ctx)
    assert per == (256 if sparse else 128)
    allh = (C.c_ubyte * (world * per))()
    for r, b in enumerate(ranks):
        mine = (C.c_ubyte * per)()
        _lib.check(lib.bigclam_ipc_export(b._ctx, mine), b._ctx)
        C.memmove(C.addressof(allh) + r * per, mine, per)
    for r, b in enumerate(ranks):
        _lib.check(lib.bigclam_ipc_open_peers(b._ctx, world, r, allh), b._ctx)

    P = oracle.make_params(k)
    F, s = F0, sumF
    for it in range(3):
        ref = oracle.step(rp, col, F, s, P)
        parts = []
        for b in ranks:                                   # step kernels (each pushes its rows into the other replica)
            p = C.c_void_p()
            _lib.check(lib.bigclam_step_local(b._ctx, C.byref(p)), b._ctx)
            parts.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(2 * ld + 2,)))
        total = parts[0] + parts[1]                       # the all-reduce
        for p in parts:
            p[:] = total
        for b in ranks:
            llh = C.c_double()
            nupd = C.c_int64()
            _lib.check(lib.bigclam_finish_local(b._ctx, C.byref(llh), C.byref(nupd)), b._ctx)
            assert abs(llh.value - oracle.llh(rp, col, F, s, P)) <= 1e-10 * abs(llh.value)
            assert nupd.value == ref.n_updated
        for b in ranks:                                   # every replica holds the new state
            assert np.abs(b.F - ref.F).max() <= 1e-9 * max(np.abs(ref.F).max(), 1e-300)
            assert np.allclose(b.sumF, ref.sumF, rtol=1e-11, atol=1e-9)
        F, s = ref.F, ref.sumF
    for b in ranks:
        b.close()
