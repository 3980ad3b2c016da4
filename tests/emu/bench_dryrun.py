"""Dry run of bench.py's single-GPU arm on a machine WITHOUT a GPU (test infrastructure, like the rest of tests/emu):
the ctypes binding is pointed at the host-emulation build of the library and torch.cuda's stream / event / pinning calls
are replaced by host stand-ins, so that the control flow of bench.run_single — every leg, the JSON line, the watchdog —
runs on a tiny graph.  Nothing it prints is a measurement.

    python tests/emu/bench_dryrun.py [bench.py flags ...]
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from bigclam_apachespark_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(REPO, "tests", "emu", "libbigclam_hostemu.so")


class _Stream:
    cuda_stream = 0


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-6)


torch.cuda.set_device = lambda *_a, **_k: None
torch.cuda.current_stream = lambda *_a, **_k: _Stream()
torch.cuda.synchronize = lambda *_a, **_k: None
torch.cuda.Event = _Event
torch.Tensor.pin_memory = lambda self, *_a, **_k: self

import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    bench.main()
