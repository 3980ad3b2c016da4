"""Bounds vs exhaustive line search + oracle parity on a BFS ball of com-amazon (K=200, synthetic F0, sumF scaled as if the
graph were `big` times larger) under the HOST EMULATION of the kernels (test infrastructure, tests/emu): the regime of
BASELINE config 3 — tiles, tile fallbacks, nodes on the general path, split hubs — at a size the emulation finishes.

    python tests/emu/amazon_subgraph.py <nodes> <steps> <big>        # 1200 6 250: 1.5 min; 8000 10 40: 16 min (8 cores)

Needs tests/emu/libbigclam_hostemu.so (tests/emu/build_hostemu.sh)."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, 'tests')); sys.path.insert(0, REPO)
os.environ["BIGCLAM_HOSTEMU"] = "1"; os.environ["BIGCLAM_HOSTEMU_NOBUILD"] = "1"
import conftest
import test_gpu_prune as T
from oracle import oracle as O
from bigclam_apachespark_b200 import graphs as G
O.build()
rp, col, _ = G.load_npz_graph("com-amazon")
n = len(rp) - 1
# BFS ball
N = int(sys.argv[1]); steps = int(sys.argv[2]); big = float(sys.argv[3])
seen = -np.ones(n, dtype=np.int64); order = []
q = [12345]; seen[12345] = 0
while q and len(order) < N:
    u = q.pop(0); order.append(u)
    for v in col[rp[u]:rp[u+1]]:
        if seen[v] < 0:
            seen[v] = 1; q.append(int(v))
order = np.array(order[:N]); idx = -np.ones(n, dtype=np.int64); idx[order] = np.arange(len(order))
us, vs = [], []
for u in order:
    for v in col[rp[u]:rp[u+1]]:
        if idx[v] >= 0 and idx[u] < idx[v]:
            us.append(idx[u]); vs.append(idx[v])
srp, scol = G.csr_from_undirected(len(order), np.array(us), np.array(vs))
K = 200
F0 = G.synthetic_F0(len(order), K, seed=1234, density=0.05)
print("sub graph", len(order), len(scol), "max deg", np.diff(srp).max(), flush=True)
t0 = time.time()
asked, searched = T._run_both(srp, scol, K, F0, O.colsum(F0) * big, steps, oracle=O, where="amazon-sub")
print("ok", asked, searched, "sec", time.time() - t0, flush=True)
