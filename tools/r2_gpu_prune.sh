#!/bin/bash
# GPU check of the line search by bounds: parity tests, then A/B timing default vs exhaustive (tools/ab_libs.py)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prune.py tests/test_gpu_sparse.py -m gpu -x -q -s > gpurun_out/r2p_pytest.log 2>&1; tail -5 gpurun_out/r2p_pytest.log
timeout 900 python tools/ab_libs.py "$@" > gpurun_out/r2p_ab.log 2>&1; cut -c1-400 gpurun_out/r2p_ab.log | tail -12
