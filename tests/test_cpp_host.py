"""include/bigclam_b200.hpp — the compiled (C++17, header-only) host-side mirror of the reference's spark-shell surface —
through a driver that reads like the script (tests/cpp_host/script_mirror.cpp).

The reference's host code is JVM code and the image has no JVM, so the compiled mirror above the C ABI is C++; the Python
mirror is bigclam_apachespark_b200/driver.py (same names).  CPU: the header compiles with -Wall -Wextra -Werror and links
with the product library; Kset() reproduces the pasted REPL value of bigclam4-7.scala:268; without a CUDA device set_K()
throws (exit 1, code -2: no CPU path).  Linked with the host-emulation build of the C API (test infrastructure) every
entry of the surface is checked against the oracle: backtrackingLineSearchs call by call (:152-223), SGDFindC (:225-243),
MBSGD of v3 and v2, numGPUs=2, and the K sweep (:244-266) against an oracle-driven sweep.  `-m gpu`:
tests/test_gpu_zy_c_host.py runs the same driver against the product library."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import REPO, random_graph
from test_c_host import EMU_DIR, PRODUCT_DIR, _case, _check_against_oracle, _read_out, _write_edgelist, cached_build

SRC = os.path.join(REPO, "tests", "cpp_host", "script_mirror.cpp")


def build_mirror(out, libdir, libname):
    return cached_build(("cpp", libdir), out, lambda o: _build_now(o, libdir, libname))


def _build_now(out, libdir, libname):
    cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else shutil.which("g++")
    cmd = [cxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(REPO, "include"), SRC, "-o", out,
           "-L", libdir, f"-l:{libname}", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def _emu_exe(tmp_path):
    if not os.path.exists(os.path.join(EMU_DIR, "libbigclam_hostemu.so")):
        subprocess.run([os.path.join(EMU_DIR, "build_hostemu.sh")], check=True)
    return build_mirror(str(tmp_path / "script_mirror_emu"), EMU_DIR, "libbigclam_hostemu.so")


def _run(exe, *args, timeout=800):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def oracle_sweep(oracle, rp, col, ks, max_outer):
    """The script's sweep (:244-266) by the CPU restatement: reference-style F0 per K (NumPy twin of initNeighborComF),
    SGDFindC, stop at the first K with `1 - LLH_K / LLH_prev < 0.001` (LLHKold starts at 0.0 as coded)."""
    from oracle import numpy_twin as T
    ranked, _ = T.conductance_local_min(rp, col)
    old, kfor, hist = 0.0, 0, []
    for K in ks:
        F0 = T.init_neighbor_com_F(rp, col, K, ranked)
        llh = oracle.run(rp, col, F0, oracle.colsum(F0), oracle.make_params(K), variant=4, rel_tol=1e-4, max_outer=max_outer)[2]
        hist.append((K, llh))
        with np.errstate(divide="ignore", invalid="ignore"):
            gain = 1.0 - np.float64(llh) / np.float64(old)
        if gain < 0.001:
            kfor = K
            break
        old = llh
    return kfor, hist


def check_sweep_output(oracle, out, rp, col, ks, max_outer):
    lines = out.strip().splitlines()
    hist = [(int(l.split()[0]), float(l.split()[2])) for l in lines if " LLH: " in l]
    kfor = int(lines[-1].split()[1])
    k_o, hist_o = oracle_sweep(oracle, rp, col, ks, max_outer)
    assert kfor == k_o and [k for k, _ in hist] == [k for k, _ in hist_o], (kfor, k_o, hist, hist_o)
    for (_, l1), (_, l2) in zip(hist, hist_o):
        assert abs(l1 - l2) <= 1e-8 * abs(l2), (l1, l2)


# ------------------------------------------------------------------------------------------------
def test_mirror_compiles_links_and_kset_known_answer(tmp_path):
    exe = build_mirror(str(tmp_path / "script_mirror"), PRODUCT_DIR, "libbigclam_b200.so")
    # bigclam4-7.scala:268 (minCom=50, maxCom=200, divCom=15), the only known-answer vector the reference holds
    assert _run(exe, "kset", 50, 200, 15).split() == "50 54 59 64 70 76 83 91 99 108 118 129 141 154 168 184 200".split()


def test_mirror_throws_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rp, col, F0, edges, f0 = _case(tmp_path)
    exe = build_mirror(str(tmp_path / "script_mirror"), PRODUCT_DIR, "libbigclam_b200.so")
    for world in (1, 2):
        r = subprocess.run([exe, "sgd", edges, "8", "3", f0, str(tmp_path / "o.bin"), str(world)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and r.stderr.startswith("bigclam_b200 error -2: "), (r.returncode, r.stderr)


@pytest.mark.timeout(900)
def test_mirror_hot_path_call_by_call(oracle, tmp_path):
    rp, col, F0, edges, f0 = _case(tmp_path, n=120, deg=4, k=6, seed=21)
    out = str(tmp_path / "o.bin")
    _run(_emu_exe(tmp_path), "steps", edges, 6, 3, f0, out, 1)
    calls, llh, trace, sumF, F = _read_out(out)
    P, Fo, so = oracle.make_params(6), F0, oracle.colsum(F0)
    for it in range(3):
        r = oracle.step(rp, col, Fo, so, P)
        assert abs(trace[it] - r.llh) <= 1e-9 * abs(r.llh), it
        Fo, so = r.F, r.sumF
    assert np.abs(F - Fo).max() <= 1e-9 * np.abs(Fo).max() and np.allclose(sumF, so, rtol=1e-8)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [1, 2])
def test_mirror_sgdfindc(oracle, tmp_path, world):
    rp, col, F0, edges, f0 = _case(tmp_path, n=120, deg=4, k=6, seed=22)
    out = str(tmp_path / "o.bin")
    stdout = _run(_emu_exe(tmp_path), "sgd", edges, 6, 4, f0, out, world)
    assert stdout.startswith("-------Inter: 1 LLH: ")                      # the script's progress line (:236)
    _check_against_oracle(oracle, rp, col, F0, 6, out, 4)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("version", [3, 2])
def test_mirror_mbsgd(oracle, tmp_path, version):
    rp, col, F0, edges, f0 = _case(tmp_path, n=120, deg=4, k=6, seed=23)
    out = str(tmp_path / "o.bin")
    _run(_emu_exe(tmp_path), "mbsgd", edges, 6, 4, f0, out, 1, version)
    calls, llh, trace, sumF, F = _read_out(out)
    Fo, so, llho, callso, tro = oracle.run(rp, col, F0, oracle.colsum(F0), oracle.make_params(6), variant=version, max_outer=4)
    assert calls == callso and np.allclose(trace, tro, rtol=1e-9)
    assert np.abs(F - Fo).max() <= 1e-9 * np.abs(Fo).max() and np.allclose(sumF, so, rtol=1e-8)


@pytest.mark.timeout(900)
def test_mirror_k_sweep(oracle, tmp_path):
    from bigclam_apachespark_b200.driver import Kset
    rp, col = _case(tmp_path, n=150, deg=5, k=4, seed=24)[:2]
    edges = str(tmp_path / "sweep.txt")
    _write_edgelist(edges, rp, col)
    ks = Kset(4, 8, 2)
    out = _run(_emu_exe(tmp_path), "sweep", edges, 4, 8, 2, 5)
    check_sweep_output(oracle, out, rp, col, ks, 5)


@pytest.mark.timeout(900)
def test_mirror_extraction(oracle, tmp_path):
    """Bigclamv2.scala:223-230 from C++ (delta_threshold + extract) against the rule restated in NumPy on the oracle's F."""
    import math
    rp, col, F0, edges, f0 = _case(tmp_path, n=120, deg=4, k=6, seed=25, dens=0.5)
    n = len(rp) - 1
    out = _run(_emu_exe(tmp_path), "extract", edges, 6, 2, f0, n)
    lines = out.strip().splitlines()
    e = 2.0 * n / (n * (n - 1.0))
    delta = math.sqrt(-math.log(1.0 - e))
    assert abs(float(lines[0].split()[1]) - delta) <= 1e-15
    P, Fo, so = oracle.make_params(6), F0, oracle.colsum(F0)
    for _ in range(2):
        r = oracle.step(rp, col, Fo, so, P)
        Fo, so = r.F, r.sumF
    fmax = Fo.max(axis=1, keepdims=True)
    member = np.where(fmax >= delta, Fo >= delta, Fo == fmax)                 # :227 (ties included)
    margin = np.abs(Fo - delta).min()
    assert margin > 1e-9, "an entry sits on the threshold: pick another seed"
    got = {int(l.split(":")[0]): [int(x) for x in l.split(":")[1].split()] for l in lines[1:]}
    want = {c: np.flatnonzero(member[:, c]).tolist() for c in range(6) if member[:, c].any()}
    assert got == want
