#!/bin/bash
# Builds tests/emu/libbigclam_hostemu.so: csrc/bigclam_capi.cu and the kernels compiled for the host against the
# SIMT emulation and the synchronous runtime stand-in of cuda_emu.h.  DEVELOPMENT TOOL: lets the `-m gpu` tests
# drive the host logic of the C API on a machine without a GPU,
#     BIGCLAM_HOSTEMU=1 python -m pytest tests/test_gpu_parity.py -m gpu -k "not amazon and not enron ..."
# (tests/conftest.py points the ctypes binding at this file only when that variable is set).  It is not the
# product, is not built by __graft_entry__.build() and is never loaded by the package on its own.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/../../bigclam_apachespark_b200/csrc"
gen="$here/_gen"
mkdir -p "$gen"
defs=""
for f in bigclam_kernels.cuh bigclam_sparse.cuh bigclam_tile.cuh; do
  test -f "$src/$f" || continue
  sed -e 's/extern __shared__ __align__(16) unsigned char smem_raw\[\];/unsigned char *smem_raw = emu::dyn_smem();/' \
      -e 's/^\( *\)__shared__ /\1static /' "$src/$f" > "$gen/$f"
done
python3 "$here/rewrite_launches.py" "$src/bigclam_capi.cu" | sed -e 's#"../../include/bigclam_b200.h"#"bigclam_b200.h"#' > "$gen/bigclam_capi_emu.cpp"
python3 "$here/rewrite_launches.py" "$src/initf_gpu.cu" | sed -e 's#"../../include/bigclam_b200.h"#"bigclam_b200.h"#' > "$gen/initf_gpu_emu.cpp"
cat > "$gen/emu_globals.cpp" <<'EOT'
#include "cuda_emu.h"
thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
namespace emu {
thread_local WarpCtx *warp = nullptr;
thread_local BlockCtx *block = nullptr;
thread_local int lane = 0;
thread_local int xpar = 0;
unsigned char *g_dyn_smem = nullptr;
}
EOT
CXX=/usr/bin/g++; test -x $CXX || CXX=g++
$CXX -O1 -g -std=c++20 -ffp-contract=off -DBIGCLAM_EMU -DBIGCLAM_EMU_HOST $EMU_DEFS -I "$here/include" -I "$gen" -I "$here/../../include" \
    -pthread -shared -fPIC -Wno-unknown-pragmas -o "$here/libbigclam_hostemu.so" \
    "$gen/bigclam_capi_emu.cpp" "$gen/initf_gpu_emu.cpp" "$gen/emu_globals.cpp" "$src/edgelist.cpp" "$src/initf.cpp"
