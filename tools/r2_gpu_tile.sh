#!/bin/bash
# GPU validation + A/B of the tile kernel variants (tools/build_variant.sh), logs under gpurun_out/r2t_*
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 1. sparse GPU tests"; timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/r2t_pytest_sparse.log 2>&1; tail -4 gpurun_out/r2t_pytest_sparse.log
echo "== 2. A/B"; timeout 600 python tools/ab_libs.py "$@" > gpurun_out/r2t_ab.log 2>&1; tail -12 gpurun_out/r2t_ab.log
echo "== 3. ncu of the tile kernel"; BIGCLAM_AB_SPARSE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_step_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/r2t_prof_tile python tools/profile_step.py 200 6 2 > gpurun_out/r2t_ncu.log 2>&1; tail -3 gpurun_out/r2t_ncu.log
