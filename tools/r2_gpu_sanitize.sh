#!/bin/bash
# compute-sanitizer over a selection of the parity tests; summaries -> gpurun_out/r2s_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL_SPARSE='golden_tiny or uset_mask or split_hubs or pool_exhaustion or (all_k and (k5- or 31 or 200 or 1000))'
SEL_DENSE='golden_tiny or mask or (all_k and (k5- or 200))'
for tool in memcheck racecheck; do
  echo "== $tool: sparse-row engine"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "$SEL_SPARSE" > gpurun_out/r2s_${tool}_sparse.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2s_${tool}_sparse.log | tail -3
  echo "== $tool: dense kernels"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_DENSE" > gpurun_out/r2s_${tool}_dense.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2s_${tool}_dense.log | tail -3
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== memcheck: two GPUs through bigclam_multi_*"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "world2 or 2-" > gpurun_out/r2s_memcheck_multi.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2s_memcheck_multi.log | tail -3
fi
