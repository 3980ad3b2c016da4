"""community extraction (Bigclamv2.scala:223-230) and the Avg-F1 scorer."""
import math

import numpy as np
import pytest

from conftest import random_graph


def test_delta_threshold_as_coded():
    from bigclam_apachespark_b200.communities import delta_threshold
    # e = 2*count/(N(N-1)); delta = sqrt(-log(1-e))   (Bigclamv2.scala:223-224)
    n, count = 4039, 4039
    e = 2.0 * count / (n * (n - 1.0))
    assert delta_threshold(n, count) == math.sqrt(-math.log(1.0 - e))
    assert delta_threshold(n, 88234) > delta_threshold(n, count)      # thesis def. 9 uses |E|: a larger threshold


def test_avg_f1_and_ground_truth_fixture():
    from bigclam_apachespark_b200 import communities as Cm
    gt = Cm.load_ground_truth()
    assert len(gt) == 75149 and min(len(g) for g in gt) >= 3 and max(len(g) for g in gt) == 53551
    assert all((np.diff(g) > 0).all() for g in gt[:200])
    n = 334863
    sub = gt[:500]
    assert Cm.avg_f1(sub, sub, n) == 1.0
    half = [g[: max(1, len(g) // 2)] for g in sub]
    f = Cm.avg_f1(half, sub, n)
    assert 0.5 < f < 0.95
    assert Cm.avg_f1([], sub, n) == 0.0
    # hand-checked: C = {0,1,2}, G = {1,2,3}: overlap 2 -> F1 = 2*2/(3+3)
    assert abs(Cm.avg_f1([np.array([0, 1, 2])], [np.array([1, 2, 3])], 10) - 2 / 3) < 1e-15


@pytest.mark.gpu
def test_extract_matches_numpy():
    from bigclam_apachespark_b200 import BigClam
    from bigclam_apachespark_b200.communities import delta_threshold, extract
    n, k = 700, 37
    rp, col = random_graph(n, 6, seed=12)
    rng = np.random.default_rng(12)
    F = rng.random((n, k)) * (rng.random((n, k)) < 0.2)
    F[5] = 0.0                                   # all-zero row: every entry equals the maximum
    F[6, :] = 0.0; F[6, [3, 9]] = 1e-4           # below delta with a tie for the maximum
    b = BigClam()
    b.set_graph(rp, col).set_K(k).set_F(F)
    delta = delta_threshold(n, n)
    comms, cids, member = extract(b, delta)
    fmax = F.max(axis=1, keepdims=True)
    expect = np.where(fmax < delta, F == fmax, F >= delta)     # :225-228
    assert np.array_equal(member.astype(bool), expect)
    assert member[5].all() and list(np.flatnonzero(member[6])) == [3, 9]
    for c, idx in zip(cids, comms):
        assert np.array_equal(idx, np.flatnonzero(expect[:, c]))
    b.close()
