#!/usr/bin/env bash
# Builds libspcl_b200.so (C ABI, include/spcl.h) for sm_100a.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${SPCL_OUT:-${HERE}/../libspcl_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
  -Xcompiler -fPIC -shared --cudart=static ${SPCL_NVCC_EXTRA:-} \
  -o "${OUT}" "${HERE}/aux_kernels.cu" "${HERE}/dense_frontend.cu" "${HERE}/supcon_simt.cu" "${HERE}/supcon_tc.cu"
echo "built ${OUT}"
