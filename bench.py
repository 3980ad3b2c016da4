#!/usr/bin/env python
"""bench.py — BASELINE.json metric: "edges/sec in F-gradient step at 1/2/4/8 B200; iters/sec on
com-amazon K=200".

A step = one call of the hot path (backtrackingLineSearchs, codes/bigclam4-7.scala:152-223: PRE +
16-candidate line search + row swap + sumF update + LLH) over the whole graph.  Workload: the
com-amazon topology (SNAP, 334,863 nodes / 925,872 edges, committed as tests/golden/graphs/
com-amazon.npz) with K=200 and the synthetic F0 of BASELINE.md (U[0,1) with probability 0.05,
seed 1234), fp64.  F is 536 MB (> the 126 MB L2), so no L2 flush is needed between steps.

  value   directed neighbour-list entries processed per second, F resident in HBM, K steps run by
          the device-side loop (bigclam_run), timed with CUDA events on the launching stream
  e2e     the same metric through per-call bigclam_step() with host buffers (uset mask H2D from
          pinned memory, LLH/n_updated D2H every step)
  roofline  algorithmic bytes of one step kernel / its average duration (CUDA events in the library)
  cpu_baseline  the CPU restatement of the reference (oracle/, NOT Spark) on the host cores
`--impl reference` times that CPU restatement alone (the reference needs a JVM + Spark: absent).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOAD = "com-amazon K=200, synthetic F0 (p=0.05 U[0,1), seed 1234), fp64"
K = 200


_GRAPH = "com-amazon"


def load_workload():
    """Default: the com-amazon topology fixture.  `--graph rmat:<nodes>:<edges>` generates BASELINE config 5's
    R-MAT family instead ((a,b,c,d) = (.57,.19,.19,.05), seed 42) — not the headline workload."""
    from bigclam_apachespark_b200 import graphs as G
    if _GRAPH.startswith("rmat:"):
        _, nn, mm = _GRAPH.split(":")
        rp, col = G.rmat_graph(int(nn), int(mm), seed=42)
    else:
        rp, col, _ = G.load_npz_graph(_GRAPH)
    n = len(rp) - 1
    if n * K > (1 << 31):
        # n x K does not exist densely at this size: CSR rows (same distribution), needs --layout sparse
        import scipy.sparse as sps
        ip, ix, vl = G.synthetic_F0_csr(n, K, seed=1234, density=0.05)
        return rp, col, sps.csr_matrix((vl, ix, ip), shape=(n, K))
    F0 = G.synthetic_F0(n, K, seed=1234, density=0.05)
    return rp, col, F0


def alg_bytes(n, nnz, k, s=8):
    """nnz*(K*s+4) + N*(2*K*s+8) + K*s."""
    return nnz * (k * s + 4) + n * (2 * k * s + 8) + k * s


def hbm_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints), every ~5 ms from a thread."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            self.h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
        except Exception as exc:            # noqa: BLE001
            self.N = None
            self.err = repr(exc)
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        N = self.N
        bits = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:               # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        if self.N is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self._t.join(timeout=1)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def time_oracle(rp, col, F0, steps, warmup):
    """Faithful CPU restatement (all 16 candidates per node, like the reference); all host threads."""
    from oracle import oracle as O
    O.build()
    P = O.make_params(K)
    F, s = F0, O.colsum(F0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = O.step(rp, col, F, s, P, early_exit=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        F, s = r.F, r.sumF
    return float(np.mean(times)), O.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rp, col, F0 = load_workload()
    n, nnz = len(rp) - 1, len(col)
    sec, cores = time_oracle(rp, col, F0, args.steps, args.warmup)
    val = nnz / sec
    sample = f"{args.steps} full steps of the workload (all 16 candidates per node), {args.warmup} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": "edges/sec in F-gradient step", "value": val, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "iters_per_sec": 1.0 / sec, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "com-amazon topology (SNAP fixture) + synthetic F0",
        "config": {"workload": WORKLOAD, "n": n, "nnz_directed": nnz, "k": K,
                   "note": "CPU restatement of the reference (oracle/, C + OpenMP), NOT Spark: no JVM in the image"},
        "cpu_baseline": {"value": val, "unit": "edges/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_single(args):
    import torch
    from bigclam_apachespark_b200 import BigClam

    torch.cuda.set_device(0)
    rp, col, F0 = load_workload()
    n, nnz = len(rp) - 1, len(col)
    sparse = args.layout == "sparse"
    b = BigClam(device=0, time_kernels=True, sparse_rows=sparse)
    b.set_graph(rp, col).set_K(K)
    stream = torch.cuda.current_stream()
    b.set_stream(stream.cuda_stream)
    b.set_F(F0)

    # ---- value: device-resident loop ----
    b._run(4, 0.0, args.warmup)                      # W untimed warm-up steps
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    b._run(4, 0.0, args.steps)                       # exactly K steps (rel_tol 0: never converges early)
    e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    assert b.last_calls == args.steps
    total_ms = e0.elapsed_time(e1)
    kern_ms, n_step_kernels, n_all = b.kernel_time()
    ms_per_step = total_ms / args.steps
    value = nnz / (ms_per_step * 1e-3)
    llh_end = float(b.last_trace[-1])

    # ---- roofline of the dominant kernel (step_kernel) ----
    peak, peak_src = hbm_peak()
    balg = alg_bytes(n, nnz, K)
    kavg_ms = kern_ms / max(n_step_kernels, 1)
    achieved = balg / (kavg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("step_kernel_dram_bytes_per_launch")

    layout_bytes = None
    if sparse:
        # bytes the sparse layout actually has to move per launch (row blocks: 10 B per padded entry + 8 B header):
        # every neighbour row once per edge, every own row read and written once
        cnt = np.diff(b.F_csr()[0])
        blk = 10 * ((cnt + 3) // 4 * 4) + 8
        layout_bytes = int((blk[col].sum() + 4 * nnz) + 2 * blk.sum() + 16 * n)

    # ---- e2e: per-call C ABI with host buffers ----
    mask = torch.ones(n, dtype=torch.uint8).pin_memory()
    llh = C.c_double(); nupd = C.c_int64()
    from bigclam_apachespark_b200 import _lib
    lib = _lib.load()
    for _ in range(3):
        _lib.check(lib.bigclam_step(b._ctx, mask.data_ptr(), C.byref(llh), C.byref(nupd)), b._ctx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _lib.check(lib.bigclam_step(b._ctx, mask.data_ptr(), C.byref(llh), C.byref(nupd)), b._ctx)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e = {"value": nnz / (e2e_ms * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": int(n),
           "d2h_bytes_per_step": 72, "ms_per_step": e2e_ms,
           "note": "bigclam_step() per step: uset mask H2D (pinned) + sumF commit + the next call's step kernel launched speculatively (its PRE is this call's LLH) + LLH/n_updated D2H, synchronous; F stays resident like the reference's cached RDD"}

    # ---- workload A of SURVEY §8d: the reference's own init (conductance seeds, 0/1 indicator columns) ----
    extra_a = None
    if not args.no_init_a and not hasattr(F0, "tocsr"):
        t0 = time.perf_counter()
        b.initNeighborComF(K)
        init_s = time.perf_counter() - t0
        b._run(4, 0.0, args.warmup)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        b._run(4, 0.0, args.steps)
        eb.record(stream)
        torch.cuda.synchronize()
        ms_a = ea.elapsed_time(eb) / args.steps
        kms_a, nk_a, _ = b.kernel_time()
        extra_a = {"workload": "com-amazon K=200, F0 = initNeighborComF(200) (bigclam4-7.scala:81-108: 0/1 indicator columns of the 200 best-conductance seeds)",
                   "value": nnz / (ms_a * 1e-3), "unit": "edges/s", "ms_per_step": ms_a, "step_kernel_ms": kms_a / max(nk_a, 1),
                   "roofline_frac": balg / (kms_a / max(nk_a, 1) * 1e-3) / 1e9 / peak, "host_init_seconds": init_s,
                   "llh_end": float(b.last_trace[-1])}

    # ---- CPU baseline beside it (bounded: 2 faithful steps after 1 warm-up) ----
    cpu = None
    if not args.no_cpu and not hasattr(F0, "tocsr"):
        sec, cores = time_oracle(rp, col, F0, 2, 1)
        cpu = {"value": nnz / sec, "unit": "edges/s", "cores": cores, "kind": "port",
               "sample": "2 full steps of the same workload (all 16 candidates per node) after 1 warm-up; CPU restatement of the reference, not Spark",
               "ms_per_step": sec * 1e3}

    print(json.dumps({
        "metric": "edges/sec in F-gradient step", "value": value, "unit": "edges/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "iters_per_sec": 1e3 / ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "com-amazon topology (SNAP fixture) + synthetic F0",
        "config": {"workload": WORKLOAD, "n": n, "nnz_directed": nnz, "k": K, "parallelism": "1 GPU",
                   "l2": "inputs (F 536 MB x2 buffers) larger than L2, no flush", "llh_end": llh_end},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(n_all),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None if sparse else traffic, "kernel": "sparse_step_kernel" if sparse else "step_kernel<4>",
                     "kernel_ms": kavg_ms, "alg_bytes_per_launch": balg, "peak_source": peak_src,
                     "f_layout": args.layout, "layout_bytes_per_launch": layout_bytes},
        "cpu_baseline": cpu, "reference_init_workload": extra_a,
    }))
    b.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-init-a", action="store_true", help="skip the reference-style-init extra workload")
    ap.add_argument("--k", type=int, default=200, help="number of communities (default: the headline K = 200)")
    ap.add_argument("--layout", default="dense", choices=["dense", "sparse"],
                    help="device layout of F: dense n x K rows, or sparse rows like the reference's BSV[Double]")
    ap.add_argument("--graph", default="com-amazon", help="fixture name or rmat:<nodes>:<edges> (default: the headline workload)")
    args = ap.parse_args()
    global _GRAPH, WORKLOAD, K
    _GRAPH = args.graph
    K = args.k
    if _GRAPH != "com-amazon" or K != 200:
        WORKLOAD = f"{_GRAPH} K={K}, synthetic F0 (p=0.05 U[0,1), seed 1234), fp64"
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        args.steps = min(args.steps, 20)     # bounded: each step is ~1-4 s of all-core CPU work
        args.warmup = min(args.warmup, 1)
        return run_reference(args)
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if args.layout == "sparse":
            os.environ["BIGCLAM_SPARSE"] = "1"
        from bigclam_apachespark_b200 import dist
        return dist.bench_main(args, load_workload, alg_bytes, hbm_peak, ClockSampler, WORKLOAD, K)
    return run_single(args)


if __name__ == "__main__":
    main()
