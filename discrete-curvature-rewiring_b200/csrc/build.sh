#!/usr/bin/env bash
# Build libdcr.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libdcr.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared)
if [[ "${DCR_PTXAS_V:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
"${NVCC}" "${FLAGS[@]}" -o "${OUT}" "${HERE}"/*.cu
echo "built ${OUT}"
