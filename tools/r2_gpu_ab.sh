#!/bin/bash
# A/B only (tools/ab_libs.py) of prebuilt variants; log gpurun_out/r2_ab_<tag>.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
timeout 900 python tools/ab_libs.py "$@" > gpurun_out/r2_ab_$tag.log 2>&1; cut -c1-170 gpurun_out/r2_ab_$tag.log | tail -12
