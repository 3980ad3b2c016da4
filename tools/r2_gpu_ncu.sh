#!/bin/bash
# one ncu --set full capture of the tile kernel in its steady state (launch 15 of a com-amazon K=200 run) -> gpurun_out/<tag>.ncu-rep
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2p_prof}
BIGCLAM_AB_SPARSE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_step_kernel --launch-skip ${2:-14} --launch-count 1 -f -o gpurun_out/$tag python tools/profile_step.py 200 12 4 > gpurun_out/$tag.log 2>&1; tail -3 gpurun_out/$tag.log; ls -la gpurun_out/$tag.ncu-rep
