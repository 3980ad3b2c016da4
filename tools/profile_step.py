"""Small driver for ncu: com-amazon K=200, W warm-up steps then S steps through bigclam_run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigclam_apachespark_b200 import BigClam, _lib, graphs as G  # noqa: E402

if os.environ.get("BIGCLAM_AB_LIB"):      # dev only: profile a variant built by tools/build_variant.sh
    _lib.LIB_PATH = os.path.abspath(os.environ["BIGCLAM_AB_LIB"])

K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2
name = sys.argv[4] if len(sys.argv) > 4 else "com-amazon"
if name.startswith("rmat:"):          # rmat:<nodes>:<edges>
    _, nn, mm = name.split(":")
    rp, col = G.rmat_graph(int(nn), int(mm), seed=42)
else:
    rp, col, _ = G.load_npz_graph(name)
n = len(rp) - 1
b = BigClam(device=0, time_kernels=True, sparse_rows=os.environ.get("BIGCLAM_AB_SPARSE") == "1")
b.set_graph(rp, col).set_K(K).set_F(G.synthetic_F0(n, K, seed=1234, density=0.05))
b._run(4, 0.0, W)
b._run(4, 0.0, S)
ms, nk, nall = b.kernel_time()
print(f"{name} K={K}: step kernel avg {ms / max(nk, 1):.3f} ms over {nk} launches, llh {b.last_trace[-1]:.6e}")
