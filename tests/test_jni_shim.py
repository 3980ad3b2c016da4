"""The JNI shim of INTEGRATION.md §2 as real code (integration/jni/bigclam_b200_jni.c), as far as it can be checked
without a JDK: compiled strict C99 against a stand-in <jni.h> (tests/jni_stub/jni.h: the JNI types and JNIEnv functions it
uses, with the specification's names and signatures) and driven by a fake JVM written in C (tests/jni_stub/fake_jvm.c:
copying array semantics, recorded exceptions) that plays the changed lines of the reference's driver script.

CPU: the shim type-checks against the header and links with the product library; without a CUDA device its `create`
throws (RuntimeException with the library's message — no CPU path); linked with the host-emulation build of the C API
(test infrastructure) the whole sequence create -> setF -> run / step -> getF agrees with the oracle, on one handle for
one device and for two.  `-m gpu`: tests/test_gpu_zy_c_host.py runs the same harness against the product library.
None of this involves a JVM: what is checked is the forwarding code (argument order, array release modes, error
propagation), not the JNI binding itself."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO
from test_c_host import EMU_DIR, PRODUCT_DIR, _case, _cc, _check_against_oracle, _read_out, cached_build

SHIM = os.path.join(REPO, "integration", "jni", "bigclam_b200_jni.c")
FAKE = os.path.join(REPO, "tests", "jni_stub", "fake_jvm.c")


def build_fake_jvm(out, libdir, libname):
    return cached_build(("jni", libdir), out, lambda o: _build_now(o, libdir, libname))


def _build_now(out, libdir, libname):
    cmd = [_cc(), "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(REPO, "tests", "jni_stub"),
           "-I", os.path.join(REPO, "include"), SHIM, FAKE, "-o", out, "-L", libdir, f"-l:{libname}", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def _emu():
    emu = os.path.join(EMU_DIR, "libbigclam_hostemu.so")
    if not os.path.exists(emu):
        subprocess.run([os.path.join(EMU_DIR, "build_hostemu.sh")], check=True)
    return EMU_DIR, "libbigclam_hostemu.so"


def test_shim_compiles_and_throws_without_a_device(tmp_path):
    import torch
    exe = build_fake_jvm(str(tmp_path / "fake_jvm"), PRODUCT_DIR, "libbigclam_b200.so")
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rp, col, F0, edges, f0 = _case(tmp_path)
    for world in ("1", "2"):
        r = subprocess.run([exe, edges, "8", "3", f0, str(tmp_path / "out.bin"), world, "run"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 3 and r.stderr.startswith("EXCEPTION java/lang/RuntimeException: "), (r.returncode, r.stderr)
        assert "CUDA" in r.stderr or "device" in r.stderr, r.stderr


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world,mode", [(1, "run"), (2, "run"), (1, "csr"), (2, "csr")])
def test_shim_run_under_host_emulation(oracle, tmp_path, world, mode):
    """mode csr: F travels as the reference keeps it — rows of (index, data), RDD[(Long, BSV[Double])] — through setFCsr and
    getFNnz + getFCsr instead of a dense n x K array."""
    rp, col, F0, edges, f0 = _case(tmp_path, n=120, deg=4, k=6, seed=11)
    exe = build_fake_jvm(str(tmp_path / "fake_jvm"), *_emu())
    out = str(tmp_path / "out.bin")
    r = subprocess.run([exe, edges, "6", "4", f0, out, str(world), mode], capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stderr[-2000:]
    calls, llh, trace, sumF, F = _read_out(out)
    Fo, so, llho, callso, tro = oracle.run(rp, col, F0, oracle.colsum(F0), oracle.make_params(6), variant=4, max_outer=4)
    assert calls == callso and abs(llh - llho) <= 1e-9 * abs(llho)
    assert np.abs(F - Fo).max() <= 1e-9 * np.abs(Fo).max() and np.allclose(sumF, so, rtol=1e-8)


@pytest.mark.timeout(900)
def test_shim_steps_under_host_emulation(oracle, tmp_path):
    """`def backtrackingLineSearchs(uset) = BigclamNative.step(ctx, null)` call by call: the LLH of every call and the number
    of rows it updated against the oracle's step."""
    rp, col, F0, edges, f0 = _case(tmp_path, n=120, deg=4, k=6, seed=12)
    exe = build_fake_jvm(str(tmp_path / "fake_jvm"), *_emu())
    out = str(tmp_path / "out.bin")
    r = subprocess.run([exe, edges, "6", "3", f0, out, "1", "step"], capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stderr[-2000:]
    calls, llh, trace, sumF, F = _read_out(out)
    P = oracle.make_params(6)
    Fo, so = F0, oracle.colsum(F0)
    lines = [l for l in r.stdout.splitlines() if l.startswith("step ")]
    for it in range(3):
        res = oracle.step(rp, col, Fo, so, P)
        assert abs(trace[it] - res.llh) <= 1e-9 * abs(res.llh)
        # Sx.size (:186) through nUpdated: equal up to Armijo ties at the smallest candidates (s = 1e-14, 1e-15: margins of
        # 1e-14 on |llh_u| ~ 20, decided by rounding — tests/test_gpu_parity.py::_check_step proves such ties one by one)
        nupd = int(lines[it].split(", ")[1].split()[0])
        assert abs(nupd - res.n_updated) <= 2, (lines[it], res.n_updated)
        Fo, so = res.F, res.sumF
    assert calls == 3 and np.abs(F - Fo).max() <= 1e-9 * np.abs(Fo).max() and np.allclose(sumF, so, rtol=1e-8)
