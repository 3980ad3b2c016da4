#!/bin/bash
# step-kernel time of one workload under the three line-search modes: tools/r2_gpu_cfgab.sh <K> <graph> [steps]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K=$1; G=$2; S=${3:-20}
for mode in "BIGCLAM_LS_EXHAUSTIVE=1" "BIGCLAM_LS_PRUNE=1" "BIGCLAM_LS_PRUNE=2"; do
  echo "== $mode: $(env $mode BIGCLAM_AB_SPARSE=1 timeout 900 python tools/profile_step.py $K 12 $S $G 2>&1 | tail -1)"
done
