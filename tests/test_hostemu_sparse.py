"""Kernel LOGIC of the sparse-row engine on the CPU (no GPU in this container): a selection of the `-m gpu` tests of
tests/test_gpu_sparse.py is run, in a child pytest, against the HOST-EMULATION build of the whole C API
(tests/emu/build_hostemu.sh: csrc/bigclam_capi.cu + the kernel sources compiled for the host against the SIMT
emulation of tests/emu/include/cuda_emu.h — one OS thread per CUDA thread, warp collectives as barrier rounds, so a
collective reached by part of a warp hangs instead of passing).  The tile path, the general path, split hubs, the
fixed-order reduction, the pool / CSR entry points and the device-side loop bookkeeping all run here against the
oracle.  Test infrastructure only: the product library has no CPU path and the `-m gpu` tests on the B200 remain
the parity tests of the compiled sm_100a code."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SELECTION = ("golden_tiny or uset_mask or csr_entry or pool_exhaustion or split_hubs or mode_limits or "
             "(all_k and (k1- or 3 or 31 or 65))")


def _child(test_file, selection, nobuild):
    env = dict(os.environ, BIGCLAM_HOSTEMU="1")
    if nobuild:
        env["BIGCLAM_HOSTEMU_NOBUILD"] = "1"
    else:
        env.pop("BIGCLAM_HOSTEMU_NOBUILD", None)
    cmd = [sys.executable, "-m", "pytest", os.path.join(REPO, "tests", test_file), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"]
    if selection:
        cmd += ["-k", selection]
    r = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=1400)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


@pytest.mark.timeout(1500)
def test_sparse_engine_under_host_emulation():
    _child("test_gpu_sparse.py", SELECTION, nobuild=False)


@pytest.mark.timeout(1500)
def test_multi_gpu_c_abi_under_host_emulation():
    """bigclam_multi_* on 2 and 3 emulated devices: node deal, pool regions, peer pushes, the fused collective (publish,
    flags, rank-ordered sums), the device-side loop on every rank — against the oracle, replicas bit-identical."""
    _child("test_gpu_multi.py", None, nobuild=True)
    _child("test_hostemu_multirank.py", None, nobuild=True)


@pytest.mark.timeout(900)
def test_line_search_bounds_under_host_emulation():
    """Line search by bounds (csrc/bigclam_tile.cuh H2, bigclam_sparse.cuh bound_mask) against the exhaustive search and the
    oracle: the same accepted steps, F, sumF and LLH bits — a selection of tests/test_gpu_prune.py (large-graph regime through
    an injected sumF, values next to the clamps)."""
    _child("test_gpu_prune.py", "(random_graphs and 900) or clamped", nobuild=True)


@pytest.mark.timeout(600)
def test_conductance_kernel_under_host_emulation():
    _child("test_gpu_init.py", "twin and (80 or 300) or unsorted", nobuild=True)
