"""Debug: per-node cycle counts of the step kernel (needs the v10dbg kernel variant + BIGCLAM_DEBUG_CYCLES=1)."""
import ctypes as C, os, sys
import numpy as np
os.environ["BIGCLAM_DEBUG_CYCLES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigclam_apachespark_b200 import BigClam, graphs as G, _lib
rp, col, _ = G.load_npz_graph("com-amazon")
n, K = len(rp) - 1, 200
b = BigClam(device=0, time_kernels=True)
b.set_graph(rp, col).set_K(K).set_F(G.synthetic_F0(n, K))
lib = C.CDLL(_lib.LIB_PATH)
raw = np.zeros(4 * n + 16, dtype=np.int64)
b._run(4, 0.0, 11)
assert lib.bigclam_debug_cycles(b._ctx, C.c_void_p(raw.ctypes.data)) == 0
b._run(4, 0.0, 1)
assert lib.bigclam_debug_cycles(b._ctx, C.c_void_p(raw.ctypes.data)) == 0
out = raw[:4 * n].reshape(n, 4)
ph = raw[4 * n:]
names = ["setup/pipeline", "PRE", "grad+active", "pair build", "consume", "own+decision(+LS glue)", "swap"]
print("phase warp-cycles (one step):")
print(f"   PRE sub-phases per node: loads+dots {ph[7]/n/1e3:.2f}k  reduce {ph[8]/n/1e3:.2f}k  eval {ph[9]/n/1e3:.2f}k  (rest = axpy/loop)")
nb = np.ceil(np.diff(rp)/4).sum()
print(f"   per batch of 4: loads+dots {ph[7]/nb:.0f}  reduce {ph[8]/nb:.0f}  eval {ph[9]/nb:.0f}  PRE total {ph[1]/nb:.0f} cycles; batches/node {nb/n:.2f}")
for nm, v in zip(names, ph):
    print(f"   {nm:24s} {v/1e9:7.3f} G  {100*v/ph[:7].sum():5.1f}%   per node {v/n/1e3:6.2f}k")
ms, nk, _ = b.kernel_time()
deg = np.diff(rp)
t0, cyc, sm, m = out[:, 0], out[:, 1], out[:, 2], out[:, 3]
print(f"kernel avg {ms/nk:.3f} ms; last kernel: sum node cycles {cyc.sum()/1e9:.3f} G; per-warp-slot avg {cyc.sum()/2368/1e6:.2f} Mcyc")
order = np.argsort(-cyc)[:12]
for u in order:
    print(f"  node {u} deg {deg[u]} m {m[u]} cycles {cyc[u]/1e3:.0f}k  sm {sm[u]}")
for lo, hi in [(1, 2), (2, 4), (4, 8), (8, 16), (16, 32), (32, 64), (64, 128), (128, 1000)]:
    sel = (deg >= lo) & (deg < hi)
    if sel.any():
        print(f"deg [{lo},{hi}): nodes {sel.sum():7d} mean cycles {cyc[sel].mean()/1e3:8.1f}k  cycles/edge {cyc[sel].sum()/deg[sel].sum()/1e3:6.2f}k  share of total {100*cyc[sel].sum()/cyc.sum():5.1f}%  mean m {m[sel].mean():.1f}")
# per-SM span
for name, arr in [("span", None)]:
    spans = []
    for s_ in np.unique(sm):
        sel = sm == s_
        spans.append(((t0[sel] + cyc[sel]).max() - t0[sel].min()) / 1e6)
    spans = np.array(spans)
    print(f"per-SM span Mcycles: min {spans.min():.2f} mean {spans.mean():.2f} max {spans.max():.2f}")
