#!/usr/bin/env bash
# Development tool: SASS of tc::bwd_kernel (or $1) from a cubin-only compile of supcon_tc.cu (no GPU needed).
#   tools/sass_bwd.sh [kernel-substring] [extra nvcc flags...]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
K="${1:-10bwd_kernelE}"; shift || true
/usr/local/cuda/bin/nvcc -cubin -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo "$@" \
  -o /tmp/spcl_tc.cubin "${HERE}/self-paced-contrastive-learning_b200/csrc/supcon_tc.cu"
FN=$(cuobjdump -sass /tmp/spcl_tc.cubin | grep "Function :" | grep "$K" | head -1 | awk '{print $3}')
cuobjdump -sass -fun "$FN" /tmp/spcl_tc.cubin > /tmp/spcl_sass.txt
echo "$FN: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' /tmp/spcl_sass.txt) instructions -> /tmp/spcl_sass.txt"
