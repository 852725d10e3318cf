#!/usr/bin/env bash
# build_variants.sh name1:"-DFLAG=.. -DFLAG2=.." name2:"..."  -> build/libdcr_<name>.so (tuning experiments; not shipped)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
CSRC="${HERE}/../../discrete-curvature-rewiring_b200/csrc"
OUT="${HERE}/../../build"
mkdir -p "${OUT}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# objects that do not depend on the tuning macros are compiled once
if [[ ! -f "${OUT}/rest.a" || -n "$(find "${CSRC}" -newer "${OUT}/rest.a" -name '*.cu*' ! -name 'dcr_bfc_paper.cu')" ]]; then
  objs=()
  for f in "${CSRC}"/*.cu; do
    b=$(basename "$f" .cu)
    [[ "$b" == "dcr_bfc_paper" ]] && continue
    "${NVCC}" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -c "$f" -o "${OUT}/${b}.o" &
    objs+=("${OUT}/${b}.o")
  done
  wait
  touch "${OUT}/rest.a"
fi
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( "${NVCC}" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${flags} -c "${PAPER_SRC:-${CSRC}/dcr_bfc_paper.cu}" -I"${CSRC}" -o "${OUT}/paper_${name}.o" &&
    "${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${OUT}/libdcr_${name}.so" "${OUT}/paper_${name}.o" $(ls "${OUT}"/dcr_*.o) &&
    echo "built libdcr_${name}.so (${flags})" ) &
done
wait
