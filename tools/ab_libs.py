"""A/B prebuilt kernel variants in ONE process on the GPU box (tools/build_variant.sh makes tools/ab/lib_<v>.so):
    python tools/ab_libs.py v29 v31 ...          timing on com-amazon K=200 + one-step parity against the oracle
TEST/DEV TOOL — uses oracle/ as the checker, never ships."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bigclam_apachespark_b200 import _lib, graphs as G  # noqa: E402
from bigclam_apachespark_b200.driver import BigClam  # noqa: E402
from oracle import oracle  # noqa: E402

K = 200
rp, col, _ = G.load_npz_graph("com-amazon")
n = len(rp) - 1
F0 = G.synthetic_F0(n, K, seed=1234, density=0.05)
sumF = oracle.colsum(F0)
t0 = time.time()
ref = oracle.step(rp, col, F0, sumF, oracle.make_params(K))
print(f"oracle step: {time.time() - t0:.2f} s, llh {ref.llh:.12e}", flush=True)
scale = np.abs(ref.F).max()

for v in sys.argv[1:]:
    opts = v.split(":")[1:]                    # e.g. "product:sparse": the library's sparse-row mode; ":ex": exhaustive line search
    sparse = "sparse" in opts
    ex = "ex" in opts
    tag = v
    v = v.split(":")[0]
    _lib._lib = None
    _lib.LIB_PATH = os.path.join(ROOT, "tools", "ab", f"lib_{v}.so") if v != "product" else os.path.join(ROOT, "bigclam_apachespark_b200", "libbigclam_b200.so")
    try:
        b = BigClam(device=0, time_kernels=True, record_accepted=True, sparse_rows=sparse, exhaustive_linesearch=ex)
        b.set_graph(rp, col).set_K(K).set_F(F0, sumF=sumF)
        llh = b.backtrackingLineSearchs()
        F = b.F
        row_err = np.abs(F - ref.F).max(axis=1) / scale
        acc = b.accepted()
        par = (f"1-step: rows>1e-12 {int((row_err > 1e-12).sum())} max {row_err.max():.2e} "
               f"llh_rel {abs(llh - ref.llh) / abs(ref.llh):.1e} idx_diff {int((acc != ref.accepted).sum())}")
        b.close()
        b = BigClam(device=0, time_kernels=True, sparse_rows=sparse, exhaustive_linesearch=ex)
        b.set_graph(rp, col).set_K(K).set_F(F0, sumF=sumF)
        b._run(4, 0.0, 10)
        res = []
        for _ in range(2):
            b._run(4, 0.0, 40)
            ms, nk, _ = b.kernel_time()
            res.append(ms / max(nk, 1))
        st = b.tile_stats() if sparse else {}
        if sparse and hasattr(b, "ls_stats"):
            try:
                st.update(b.ls_stats())
            except Exception:  # noqa: BLE001  (variant libraries built before the entry point existed)
                pass
        llh90 = b.last_trace[-1]
        b.close()
        # determinism: the same 6 steps on two fresh contexts must give the same bits
        bits = []
        for _ in range(2):
            b = BigClam(device=0, sparse_rows=sparse, exhaustive_linesearch=ex)
            b.set_graph(rp, col).set_K(K).set_F(F0, sumF=sumF)
            b._run(4, 0.0, 6)
            bits.append((b.last_trace[-1], b.F.tobytes(), b.sumF.tobytes()))
            b.close()
        det = bits[0] == bits[1]
        print(f"== {tag}: kernel {res[0]:.4f} / {res[1]:.4f} ms  llh@90 {llh90:.12e}  {par}  bit-identical reruns {det}  {st}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"== {v}: FAILED {e!r}", flush=True)
