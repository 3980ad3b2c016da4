#!/bin/bash
# Multi-GPU validation (run with gpurun --gpus N): C-ABI multi tests, torch.distributed tests, then the N-GPU bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
echo "== 1. bigclam_multi_* through ctypes (no torch.distributed)"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2m_pytest_multi_$N.log 2>&1; tail -3 gpurun_out/r2m_pytest_multi_$N.log
echo "== 2. one process per GPU (torch.distributed plumbing)"; timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r2m_pytest_dist_$N.log 2>&1; tail -3 gpurun_out/r2m_pytest_dist_$N.log
echo "== 3. bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2m_bench_$N.json 2> gpurun_out/r2m_bench_$N.err; tail -c 1500 gpurun_out/r2m_bench_$N.json; tail -3 gpurun_out/r2m_bench_$N.err

