"""Line search by bounds (csrc/bigclam_tile.cuh, phase H2) against the exhaustive search.

The sparse-row engine does not evaluate a candidate step that a bound on the node's objective proves unable to
pass the Armijo test (bigclam4-7.scala:181).  That is only legitimate if nothing observable changes: these tests run
the same steps on two contexts — default and BIGCLAM_F_LS_EXHAUSTIVE (all MaxInter + 1 candidates of every node, like
the reference's cartesian :172-181) — and require the SAME BITS: accepted step index of every node, F, sumF and the
returned LLH, step after step; and they check against the oracle that no accepted candidate was ever excluded.  The
parity of the default engine with the oracle itself is what tests/test_gpu_sparse.py and tests/test_gpu_parity.py
check (they run with the bounds on)."""
import numpy as np
import pytest

from conftest import random_graph
from test_gpu_parity import _check_step

pytestmark = pytest.mark.gpu


def _pair(rp, col, K, F0, sumF):
    from bigclam_apachespark_b200 import BigClam
    out = []
    for ex in (False, True):
        b = BigClam(record_accepted=True, sparse_rows=True, time_kernels=True, exhaustive_linesearch=ex)
        b.set_graph(rp, col).set_K(K).set_F(F0, sumF=sumF)
        out.append(b)
    return out


def _same_rows(bp, bx, big):
    """F of the two engines, bit for bit (large graphs: through the sparse rows, 40 MB instead of a dense n x K image)."""
    if not big:
        return bp.F.tobytes() == bx.F.tobytes()
    return all(np.array_equal(x, y) for x, y in zip(bp.F_csr(), bx.F_csr()))


def _run_both(rp, col, K, F0, sumF, steps, oracle=None, where=""):
    bp, bx = _pair(rp, col, K, F0, sumF)
    asked = searched = 0
    F, s = F0, sumF
    big = F0.size > (1 << 24)
    for it in range(steps):
        lp, lx = bp.backtrackingLineSearchs(), bx.backtrackingLineSearchs()
        ap, ax = bp.accepted(), bx.accepted()
        assert np.array_equal(ap, ax), f"{where} step {it}: accepted steps differ at nodes {np.nonzero(ap != ax)[0][:8]}"
        assert _same_rows(bp, bx, big), f"{where} step {it}: F differs"
        assert bp.sumF.tobytes() == bx.sumF.tobytes(), f"{where} step {it}: sumF differs"
        # the LLH is a sum over nodes of per-node sums: the same bits as long as every node took the same path in both engines
        # (a tile that overflows the line-search buffers only when ALL its nodes are searched goes node by node through the
        # general path, which adds a node's terms in another order: rounding-level difference of the LLH, same rows)
        if bp.tile_stats()["tiles_fallback"] == bx.tile_stats()["tiles_fallback"]:
            assert lp == lx, f"{where} step {it}: llh {lp!r} vs {lx!r}"
        else:
            assert abs(lp - lx) <= 1e-13 * abs(lx), f"{where} step {it}: llh {lp!r} vs {lx!r}"
        st = bp.ls_stats()
        sx = bx.ls_stats()
        assert sx["nodes_searched"] == sx["nodes_asked"], f"{where}: the exhaustive engine skipped nodes: {sx}"
        asked += st["nodes_asked"]
        searched += st["nodes_searched"]
        if oracle is not None:
            r = oracle.step(rp, col, F, s, oracle.make_params(K))
            # whatever the oracle accepts, the bounds must not have excluded (differences must be proven ties)
            _check_step(bp, r, lp, max_flips=2, where=f"{where} step {it}", inputs=(rp, col, F, s, oracle.make_params(K)))
            F, s = bp.F, bp.sumF
    bp.close()
    bx.close()
    return asked, searched


@pytest.mark.parametrize("k,n,deg,dens,steps,big", [(10, 300, 5, 0.3, 4, 1), (50, 400, 6, 0.12, 4, 1), (200, 500, 5, 0.05, 4, 1),
                                                     (200, 500, 5, 0.05, 6, 700), (64, 400, 4, 0.1, 6, 900)])
def test_bounds_change_nothing_random_graphs(oracle, k, n, deg, dens, steps, big):
    """`big` > 1: sumF as if the graph had `big` times as many nodes (the hot path takes sumF as given, :38,:192) —
    the regime of the large graphs, where most components of a node's gradient are -sumF_c and most candidates can be
    excluded."""
    rp, col = random_graph(n, deg, seed=100 + k)
    rng = np.random.default_rng(k)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < dens)
    asked, searched = _run_both(rp, col, k, F0, oracle.colsum(F0) * big, steps, oracle=oracle, where=f"k={k}")
    assert asked > 0
    print(f"k={k} big={big}: {searched} of {asked} nodes line-searched")
    if big > 1:
        assert searched < asked, f"the bounds excluded nothing ({searched} of {asked} nodes searched)"


@pytest.mark.parametrize("k,n,deg,hub,big", [(64, 260, 40, 0, 900), (200, 300, 12, 250, 700)])
def test_bounds_on_the_general_path(oracle, k, n, deg, hub, big, monkeypatch):
    """Nodes above the tile budget (more than 32 edges: one warp per node, chunks of 32 staged rows) and a split hub.
    The bounds of the general path are off by default (they only pay off on few workloads): BIGCLAM_LS_PRUNE=2."""
    monkeypatch.setenv("BIGCLAM_LS_PRUNE", "2")
    rp, col = random_graph(n, deg, seed=300 + k, hub=hub)
    rng = np.random.default_rng(300 + k)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.08)
    asked, searched = _run_both(rp, col, k, F0, oracle.colsum(F0) * big, 4, oracle=oracle, where=f"general k={k}")
    print(f"general path k={k}: {searched} of {asked} nodes line-searched")
    assert 0 < searched < asked


def test_bounds_with_tiny_and_clamped_values(oracle):
    """Rows with values next to the clamps: x just above / below x_lo = -log(MAX_P_), very large products (p at
    MIN_P_), rows near MAX_F_ (the bounds switch themselves off for a tile that can reach it)."""
    n, k = 240, 24
    rp, col = random_graph(n, 6, seed=7)
    rng = np.random.default_rng(7)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.25)
    F0[:60] *= 0.012          # products around 1e-4
    F0[60:120] *= 6.0         # products above -log(MIN_P_) = 9.2
    F0[120:130] = np.where(F0[120:130] > 0, 999.5, 0.0)
    asked, searched = _run_both(rp, col, k, F0, oracle.colsum(F0), 4, oracle=oracle, where="clamps")
    assert asked > 0


def test_bounds_with_uset_and_drifted_sumF(oracle):
    n, k = 300, 40
    rp, col = random_graph(n, 5, seed=21)
    rng = np.random.default_rng(21)
    F0 = rng.random((n, k)) * (rng.random((n, k)) < 0.15)
    sumF = oracle.colsum(F0) * (1.0 + 1e-9 * rng.standard_normal(k))       # (:192 never recomputes it)
    bp, bx = _pair(rp, col, k, F0, sumF)
    mask = np.nonzero(rng.random(n) < 0.6)[0]
    for it in range(3):
        lp, lx = bp.backtrackingLineSearchs(mask), bx.backtrackingLineSearchs(mask)
        assert abs(lp - lx) <= 1e-13 * abs(lx) and bp.F.tobytes() == bx.F.tobytes() and np.array_equal(bp.accepted(), bx.accepted()), f"step {it}"
    bp.close()
    bx.close()


def _baseline_graph_case(graphs, name, k, steps):
    rp, col, _ = graphs.load_npz_graph(name)
    n = len(rp) - 1
    F0 = graphs.synthetic_F0(n, k, seed=1234, density=0.05 if k >= 100 else 0.2)
    from oracle import oracle as O
    return _run_both(rp, col, k, F0, O.colsum(F0), steps, where=name)


@pytest.mark.parametrize("name,k,steps", [("facebook_combined", 10, 3), ("email-enron", 50, 3)])
def test_bounds_change_nothing_on_the_baseline_graphs(graphs, name, k, steps):
    """BASELINE configs 1-2, several steps from the synthetic F0 of the bench (config 3, com-amazon K=200, is the last test
    of the suite: tests/test_gpu_zz_amazon_prune.py)."""
    asked, searched = _baseline_graph_case(graphs, name, k, steps)
    assert searched <= asked
