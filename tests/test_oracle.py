"""CPU tests of the oracle itself: golden vectors, the NumPy twin, and the semantic traps of
SURVEY.md §8c.  The oracle is test infrastructure (oracle/); nothing here touches the GPU."""
import numpy as np
import pytest

from conftest import random_graph, tiny_graph


def test_step_sizes_repeated_multiplication(oracle):
    # T3: bigclam4-7.scala:28-33 builds the list by `stepSize *= beta`, not pow(10, -k)
    s = oracle.step_sizes(0.1, 15)
    assert len(s) == 16 and s[0] == 1.0 and s[1] == 0.1
    ref = [1.0]
    for _ in range(15):
        ref.append(ref[-1] * 0.1)
    assert np.array_equal(s, np.array(ref))
    assert s[2] == 0.010000000000000002 and s[2] != 1e-2          # the trap
    assert s[6] == 1.0000000000000004e-06


def test_kset_known_answer():
    # the only known-answer vector in the reference: bigclam4-7.scala:268
    from bigclam_apachespark_b200.driver import Kset
    assert Kset(50, 200, 15) == [50, 54, 59, 64, 70, 76, 83, 91, 99, 108, 118, 129, 141, 154, 168, 184, 200]


def test_oracle_reproduces_golden(oracle, golden, graphs):
    rp, col = tiny_graph(graphs)
    F = golden["tiny_F0"]
    P = oracle.make_params(5)
    sumF = oracle.colsum(F)
    for it in range(3):
        r = oracle.step(rp, col, F, sumF, P, early_exit=False)
        assert r.llh == golden[f"tiny_llh_{it}"]
        assert np.array_equal(r.F, golden[f"tiny_F_{it}"])
        assert np.array_equal(r.accepted, golden[f"tiny_accepted_{it}"])
        assert np.array_equal(r.sumF, golden[f"tiny_sumF_{it}"])
        F, sumF = r.F, r.sumF
    rp, col, _ = graphs.load_npz_graph("facebook_combined")
    F = graphs.synthetic_F0(len(rp) - 1, 10, seed=1234, density=0.3)
    P = oracle.make_params(10)
    sumF = oracle.colsum(F)
    for it in range(3):
        r = oracle.step(rp, col, F, sumF, P, early_exit=True)
        assert r.llh == golden[f"facebook_llh_{it}"]
        assert r.n_updated == golden[f"facebook_nupd_{it}"]
        assert np.array_equal(r.accepted, golden[f"facebook_accepted_{it}"])
        assert np.array_equal(r.sumF, golden[f"facebook_sumF_{it}"])
        assert np.array_equal(r.F[0], golden[f"facebook_Frow0_{it}"])
        F, sumF = r.F, r.sumF


@pytest.mark.parametrize("seed,n,deg,k,dens", [(1, 60, 4, 3, 0.6), (2, 200, 6, 8, 0.3), (3, 150, 3, 17, 0.1)])
def test_oracle_matches_numpy_twin(oracle, seed, n, deg, k, dens):
    from oracle import numpy_twin as T
    rp, col = random_graph(n, deg, seed, hub=20)
    rng = np.random.default_rng(seed)
    F = rng.random((n, k)) * (rng.random((n, k)) < dens)
    sumF = oracle.colsum(F)
    P = oracle.make_params(k)
    steps = oracle.step_sizes()
    for _ in range(2):
        r = oracle.step(rp, col, F, sumF, P, early_exit=False)
        Fn, sn, llh, acc = T.backtracking_line_searchs(rp, col, F, sumF)
        acc_o = np.where(r.accepted >= 0, steps[np.maximum(r.accepted, 0)], np.nan)
        same = np.nan_to_num(acc_o, nan=-1) == np.nan_to_num(acc, nan=-1)
        assert same.mean() > 0.99        # summation order differs; marginal Armijo ties may flip
        assert np.allclose(Fn[same], r.F[same], rtol=1e-12, atol=1e-14)
        if same.all():
            assert np.allclose(sn, r.sumF, rtol=1e-12)
            assert abs(llh - r.llh) <= 1e-10 * abs(r.llh)
        F, sumF = r.F, r.sumF


def test_early_exit_equals_max_passing(oracle):
    # T4: evaluating all 16 and keeping the max passing step == first pass in descending order
    rp, col = random_graph(300, 5, 11, hub=40)
    rng = np.random.default_rng(5)
    F = rng.random((300, 6)) * (rng.random((300, 6)) < 0.4)
    P = oracle.make_params(6)
    a = oracle.step(rp, col, F, oracle.colsum(F), P, early_exit=False)
    b = oracle.step(rp, col, F, oracle.colsum(F), P, early_exit=True)
    assert np.array_equal(a.F, b.F) and np.array_equal(a.accepted, b.accepted) and a.llh == b.llh
    assert (a.trials[np.diff(rp) > 0] == 16).all()


def test_llh_of_step_equals_next_pre_sum(oracle):
    # property (i): LLH returned by call t == sum_u llh_u of call t+1's PRE block
    rp, col = random_graph(250, 6, 3)
    rng = np.random.default_rng(9)
    F = rng.random((250, 7)) * (rng.random((250, 7)) < 0.3)
    P = oracle.make_params(7)
    r1 = oracle.step(rp, col, F, oracle.colsum(F), P)
    r2 = oracle.step(rp, col, r1.F, r1.sumF, P, want_pre=True)
    assert abs(r2.llh_u.sum() - r1.llh) <= 1e-12 * abs(r1.llh)
    v, per = oracle.llh(rp, col, r1.F, r1.sumF, P, per_node=True)
    assert v == r1.llh and np.array_equal(per, r2.llh_u)


def test_jacobi_and_mask_and_isolated(oracle, graphs):
    rp, col = tiny_graph(graphs)          # nodes 10, 11 have empty neighbour lists
    n, k = 12, 4
    rng = np.random.default_rng(2)
    F = rng.random((n, k))
    P = oracle.make_params(k)
    sumF = oracle.colsum(F)
    full = oracle.step(rp, col, F, sumF, P)
    assert (full.accepted[10:] == -1).all() and np.array_equal(full.F[10:], F[10:])
    # T2 Jacobi: updating a subset gives exactly the rows of the full update for that subset
    mask = np.zeros(n, dtype=np.uint8)
    mask[[0, 3, 7]] = 1
    part = oracle.step(rp, col, F, sumF, P, node_mask=mask)
    assert np.array_equal(part.F[[0, 3, 7]], full.F[[0, 3, 7]])
    others = [i for i in range(n) if i not in (0, 3, 7)]
    assert np.array_equal(part.F[others], F[others])
    # T6: sumF is incremental
    upd = part.accepted >= 0
    expect = sumF - (F[upd].sum(axis=0) - part.F[upd].sum(axis=0))
    assert np.allclose(part.sumF, expect, rtol=1e-14)


def test_all_zero_rows_never_update(oracle):
    # property (v): an all-zero node whose neighbours are all zero keeps its row (grad = -sumF <= 0)
    rp, col = random_graph(80, 4, 4)
    F = np.zeros((80, 3))
    F[:5] = 0.5                                   # only nodes 0..4 are non-zero
    P = oracle.make_params(3)
    r = oracle.step(rp, col, F, oracle.colsum(F), P)
    nb_nonzero = np.array([F[col[rp[u]:rp[u + 1]]].any() for u in range(80)])
    quiet = (~nb_nonzero) & (~F.any(axis=1))
    assert quiet.any() and np.array_equal(r.F[quiet], F[quiet])


def test_run_variants(oracle):
    rp, col = random_graph(120, 5, 8)
    rng = np.random.default_rng(1)
    F = rng.random((120, 4)) * (rng.random((120, 4)) < 0.5)
    P = oracle.make_params(4)
    sumF = oracle.colsum(F)
    F4, s4, llh4, calls4, tr4 = oracle.run(rp, col, F, sumF, P, variant=4)
    F3, s3, llh3, calls3, tr3 = oracle.run(rp, col, F, sumF, P, variant=3)
    F2, s2, llh2, calls2, tr2 = oracle.run(rp, col, F, sumF, P, variant=2)
    # T8: v4 spends one call for LLHold and v3 starts from 0.0: both stop at the same call
    assert calls4 == calls3 and np.array_equal(F4, F3) and np.array_equal(tr4, tr3)
    assert llh4 == tr4[-2] and llh3 == tr3[-1]          # v4 returns LLHold (:242)
    assert abs(1 - tr4[-1] / tr4[-2]) < 1e-4
    assert calls2 <= calls4 and np.array_equal(tr2, tr4[:calls2])
    Fm, sm, llhm, callsm, trm = oracle.run(rp, col, F, sumF, P, variant=4, max_outer=3)
    assert callsm == 3 and np.array_equal(trm, tr4[:3]) and llhm == trm[-1]
