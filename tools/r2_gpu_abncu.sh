#!/bin/bash
# A/B timing (tools/ab_libs.py "$@") followed by one ncu capture of the product's tile kernel in its steady state
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
timeout 900 python tools/ab_libs.py "$@" > gpurun_out/${tag}_ab.log 2>&1; cut -c1-420 gpurun_out/${tag}_ab.log | tail -8
bash tools/r2_gpu_ncu.sh ${tag}_prof 14
