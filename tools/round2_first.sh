#!/bin/bash
# First GPU call of the next round, everything in one go (logs under gpurun_out/r2_*):
#   locally first:   tools/round2_first.sh prebuild      (variant libraries for the A/B; they travel with the snapshot)
#   on the GPU box:  /usr/local/graft/bin/gpurun --timeout 900 -- 'tools/round2_first.sh gpu'
set -u
cd "$(dirname "$0")/.."
if [ "${1:-}" = "prebuild" ]; then
  make -C bigclam_apachespark_b200/csrc && make -C oracle
  for c in "b2p0 2 0" "b2p1 2 1" "b3p0 3 0" "b3p1 3 1"; do
    set -- $c
    tools/build_variant.sh sp_$1 -DBIGCLAM_SP_BLOCKS=$2 -DBIGCLAM_SP_PREFETCH=$3
  done
  exit 0
fi
mkdir -p gpurun_out
echo "== 1. dense GPU tests (hardened C API)"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest_dense.log 2>&1; tail -2 gpurun_out/r2_pytest_dense.log
echo "== 2. sparse GPU tests (first time on hardware)"; timeout 400 python -m pytest tests/test_gpu_sparse.py -m gpu -q > gpurun_out/r2_pytest_sparse.log 2>&1; tail -4 gpurun_out/r2_pytest_sparse.log
echo "== 3. step kernel A/B: dense vs sparse, sparse build knobs"; timeout 200 python tools/ab_libs.py product product:sparse sp_b2p0:sparse sp_b2p1:sparse sp_b3p0:sparse sp_b3p1:sparse > gpurun_out/r2_ab.log 2>&1; tail -8 gpurun_out/r2_ab.log
echo "== 4. bench, sparse layout"; timeout 200 python bench.py --layout sparse --no-cpu > gpurun_out/r2_bench_sparse.json 2> gpurun_out/r2_bench_sparse.err; tail -c 600 gpurun_out/r2_bench_sparse.json
echo "== 5. ncu of the sparse step kernel"; BIGCLAM_AB_SPARSE=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:sparse_step_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2_prof_sparse python tools/profile_step.py 200 6 2 > gpurun_out/r2_ncu.log 2>&1; tail -3 gpurun_out/r2_ncu.log
