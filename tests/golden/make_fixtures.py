"""Regenerates tests/golden/: compact copies of the SNAP graphs the reference ships and golden
input/output vectors of the CPU oracle.

Run in the build container (needs /root/reference/data, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py
The reference has no runnable implementation here (no JVM/Spark), so the golden OUTPUTS come from
oracle/bigclam_oracle.c (cross-checked against oracle/numpy_twin.py by tests/test_oracle.py);
they pin the oracle against drift and give the GPU tests reference values that travel.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import oracle as O                      # noqa: E402
from bigclam_apachespark_b200 import graphs as G    # noqa: E402

REF_DATA = "/root/reference/data"


def graphs():
    for src, name in [("facebook_combined.txt", "facebook_combined"), ("Email-Enron.txt", "email-enron"),
                      ("com-amazon.ungraph.txt", "com-amazon")]:
        (rp, col), ids = O.read_edge_list(os.path.join(REF_DATA, src), "dedup")
        out = os.path.join(G.FIXTURE_DIR, name + ".npz")
        G.save_npz_graph(out, rp, col, ids)
        rp2, col2, ids2 = G.load_npz_graph(out)
        assert np.array_equal(rp, rp2) and np.array_equal(col, col2) and np.array_equal(ids, ids2)
        print(name, len(rp) - 1, len(col), os.path.getsize(out))


def golden_steps():
    """Three consecutive hot-path calls on facebook K=10 and on a tiny ragged graph."""
    out = {}
    # tiny ragged graph: a hub, a path, an isolated node (empty neighbour list), duplicate-free
    edges = [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (1, 2), (2, 3), (6, 7), (7, 8), (8, 9), (9, 6), (5, 6)]
    n = 12   # nodes 10, 11 isolated
    u, v = np.array(edges).T
    rp, col = G.csr_from_undirected(n, u, v)
    K = 5
    F = G.synthetic_F0(n, K, seed=7, density=0.5)
    cases = [("tiny", rp, col, K, F)]
    rp, col, _ = G.load_npz_graph("facebook_combined")
    cases.append(("facebook", rp, col, 10, G.synthetic_F0(len(rp) - 1, 10, seed=1234, density=0.3)))
    for name, rp, col, K, F in cases:
        P = O.make_params(K)
        sumF = O.colsum(F)
        out[name + "_F0"] = F if name == "tiny" else np.zeros(0)
        for it in range(3):
            r = O.step(rp, col, F, sumF, P, early_exit=False, want_pre=(it == 0))
            out[f"{name}_llh_{it}"] = np.float64(r.llh)
            out[f"{name}_nupd_{it}"] = np.int64(r.n_updated)
            out[f"{name}_accepted_{it}"] = r.accepted
            out[f"{name}_sumF_{it}"] = r.sumF
            out[f"{name}_Fsum_{it}"] = np.float64(r.F.sum())
            out[f"{name}_Frow0_{it}"] = r.F[0].copy()
            if name == "tiny":
                out[f"{name}_F_{it}"] = r.F
            if it == 0:
                out[f"{name}_llh_u_0"] = r.llh_u
                out[f"{name}_gradnorm_0"] = np.linalg.norm(r.grad, axis=1)
            F, sumF = r.F, r.sumF
    np.savez_compressed(os.path.join(HERE, "oracle_steps.npz"), **out)
    print("oracle_steps.npz", os.path.getsize(os.path.join(HERE, "oracle_steps.npz")))


def ground_truth():
    """com-amazon.all.dedup.cmty.txt -> dense vertex indices, sorted + delta-encoded per community."""
    (rp, col), ids = O.read_edge_list(os.path.join(REF_DATA, "com-amazon.ungraph.txt"), "dedup")
    idmap = {int(v): i for i, v in enumerate(ids)}
    sizes, deltas = [], []
    for line in open(os.path.join(REF_DATA, "com-amazon.all.dedup.cmty.txt"), "rb"):
        xs = np.sort(np.array([idmap[int(t)] for t in line.split() if int(t) in idmap], dtype=np.int64))
        sizes.append(len(xs))
        deltas.append(np.diff(xs, prepend=0))
    out = os.path.join(G.FIXTURE_DIR, "com-amazon.cmty.npz")
    np.savez_compressed(out, sizes=np.array(sizes, dtype=np.int32), deltas=np.concatenate(deltas).astype(np.int32))
    print("com-amazon.cmty", len(sizes), os.path.getsize(out))


if __name__ == "__main__":
    graphs()
    golden_steps()
    ground_truth()
