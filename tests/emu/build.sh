#!/bin/bash
# Builds tests/emu/libemu.so: the kernel sources of csrc/ compiled for the host against cuda_emu.h.
# TEST INFRASTRUCTURE ONLY (the product library is built by csrc/Makefile with nvcc and has no CPU path).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/../../bigclam_apachespark_b200/csrc"
gen="$here/_gen"
mkdir -p "$gen"
defs=""
# (the sparse-row kernels are exercised through the host build of the whole C API: build_hostemu.sh)
for f in bigclam_kernels.cuh; do
  test -f "$src/$f" || continue
  sed -e 's/extern __shared__ __align__(16) unsigned char smem_raw\[\];/unsigned char *smem_raw = emu::dyn_smem();/' \
      -e 's/^\( *\)__shared__ /\1static /' "$src/$f" > "$gen/$f"
done
CXX=/usr/bin/g++; test -x $CXX || CXX=g++
$CXX -O1 -g -std=c++20 -ffp-contract=off -DBIGCLAM_EMU $defs $EMU_DEFS -x c++ -I "$here/include" -I "$gen" -I "$here/../../include" \
    -pthread -shared -fPIC -Wno-unknown-pragmas -o "$here/${EMU_OUT:-libemu.so}" "$here/emu_driver.cpp"
