#!/bin/bash
# A/B timing of kernel variants on the GPU box: tools/ab.sh v3 v4 v5  (files tools/ab/<name>.cuh)
set -e
cd "$(dirname "$0")/.."
cp bigclam_apachespark_b200/csrc/bigclam_kernels.cuh /tmp/kernels_keep.cuh
for v in "$@"; do
  cp tools/ab/$v.cuh bigclam_apachespark_b200/csrc/bigclam_kernels.cuh
  make -C bigclam_apachespark_b200/csrc -B > /tmp/ab_build.log 2>&1 || { tail -5 /tmp/ab_build.log; continue; }
  echo "== $v: $(python tools/profile_step.py 200 10 40 | tail -1)"
done
cp /tmp/kernels_keep.cuh bigclam_apachespark_b200/csrc/bigclam_kernels.cuh
make -C bigclam_apachespark_b200/csrc -B > /tmp/ab_build.log 2>&1
