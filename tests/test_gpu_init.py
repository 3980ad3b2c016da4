"""GPU parity of the caller in front of the hot path: the ego-net conductance kernel
(csrc/initf_gpu.cu, conductanceLocalMin(), bigclam4-7.scala:58-73) against the NumPy restatement and the host path."""
import time

import numpy as np
import pytest

from conftest import random_graph, tiny_graph

pytestmark = pytest.mark.gpu


def _both(rp, col):
    from bigclam_apachespark_b200 import BigClam
    b = BigClam()
    b.set_graph(rp, col)
    t0 = time.perf_counter()
    sg = b.conductanceLocalMin(on_gpu=True).copy()
    tg = time.perf_counter() - t0
    cg = b.conductance.copy()
    t0 = time.perf_counter()
    sh = b.conductanceLocalMin(on_gpu=False).copy()
    th = time.perf_counter() - t0
    return sg, cg, sh, b.conductance.copy(), tg, th


@pytest.mark.parametrize("seed,n,deg,hub", [(1, 80, 4, 10), (2, 300, 6, 40), (3, 500, 3, 0), (4, 2000, 8, 700)])
def test_gpu_conductance_matches_twin(seed, n, deg, hub):
    from oracle import numpy_twin as T
    rp, col = random_graph(n, deg, seed, hub=hub)
    sg, cg, sh, ch, _, _ = _both(rp, col)
    ranked, cond = T.conductance_local_min(rp, col)
    assert np.array_equal(cg, cond) and np.array_equal(ch, cond)          # integer counts, one division: exact
    assert np.array_equal(sg, ranked) and np.array_equal(sh, ranked)


def test_gpu_conductance_unsorted_lists_and_multiplicity(graphs):
    """Neighbour lists in arbitrary order and with repeated entries (GraphX semantics)."""
    from oracle import numpy_twin as T
    rp, col = tiny_graph(graphs)
    rng = np.random.default_rng(0)
    # double every list and shuffle it
    rp2 = rp * 2
    col2 = np.concatenate([rng.permutation(np.repeat(col[rp[u]:rp[u + 1]], 2)) for u in range(len(rp) - 1)]).astype(np.int32)
    sg, cg, sh, ch, _, _ = _both(rp2, col2)
    ranked, cond = T.conductance_local_min(rp2, col2)
    assert np.array_equal(cg, cond) and np.array_equal(sg, ranked) and np.array_equal(sh, ranked)


def test_gpu_conductance_com_amazon_equals_host_path(graphs):
    rp, col, _ = graphs.load_npz_graph("com-amazon")
    sg, cg, sh, ch, tg, th = _both(rp, col)
    assert np.array_equal(cg, ch) and np.array_equal(sg, sh)
    print(f"conductanceLocalMin on com-amazon: GPU path {tg * 1e3:.1f} ms (incl. H2D/D2H and the host-side ranking), host path {th * 1e3:.1f} ms")
