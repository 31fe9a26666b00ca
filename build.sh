#!/bin/bash
# Build libglcb200.so (the C-ABI shared library) in-tree for sm_100a.
set -e
cd "$(dirname "$0")"
SRC=galacticus_b200/csrc
OUT=${GLC_OUT:-galacticus_b200/libglcb200.so}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-mfma,-ffp-contract=off -shared ${GLC_FMAD:--fmad=false} ${GLC_NVCC_EXTRA} \
  -o $OUT $SRC/glc_api.cu $SRC/glc_params.cpp $(ls $SRC/host/*.cpp 2>/dev/null)
echo "built $OUT"
