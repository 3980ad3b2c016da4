#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the sparse-row engine with the line search by bounds: a selection of the
# sparse parity tests and of the bounds-vs-exhaustive tests (tile path, general path, split hub).  Logs: gpurun_out/r2s2_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL_SPARSE='golden_tiny or uset_mask or split_hubs or (all_k and (k5- or 200 or 1000))'
SEL_PRUNE='(random_graphs and (900 or 50-400)) or clamped or uset or (general_path and 200)'
for tool in memcheck racecheck; do
  echo "== $tool: sparse-row engine"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x -k "$SEL_SPARSE" > gpurun_out/r2s2_${tool}_sparse.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2s2_${tool}_sparse.log | tail -3
  echo "== $tool: line search by bounds vs exhaustive"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_prune.py -m gpu -q -x -k "$SEL_PRUNE" > gpurun_out/r2s2_${tool}_prune.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2s2_${tool}_prune.log | tail -3
done
