#!/usr/bin/env bash
# Build libmpn_b200.so for sm_100a (B200) in-tree.  No torch dependency: plain CUDA runtime, C ABI.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libmpn_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 ${MPN_NVCC_EXTRA:-})
OBJ="${HERE}/_obj"
mkdir -p "${OBJ}"
# objects built with other flags are stale
if [ "$(cat "${OBJ}/.flags" 2>/dev/null)" != "${FLAGS[*]}" ]; then rm -f "${OBJ}"/*.o; echo "${FLAGS[*]}" > "${OBJ}/.flags"; fi
pids=()
for f in graph gemm_simt gemm_tc gemm_api gram_ef edge_features mpn_forward postproc split_exact eval; do
  if [ ! -f "${OBJ}/${f}.o" ] || [ "${HERE}/${f}.cu" -nt "${OBJ}/${f}.o" ] || [ "${HERE}/common.cuh" -nt "${OBJ}/${f}.o" ] || \
     [ "${HERE}/kernels.h" -nt "${OBJ}/${f}.o" ] || [ "${HERE}/tcgen05.cuh" -nt "${OBJ}/${f}.o" ] || [ "${HERE}/edge_feature_gather.inc" -nt "${OBJ}/${f}.o" ] || [ "${HERE}/../../include/mpn_b200.h" -nt "${OBJ}/${f}.o" ]; then
    "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${OBJ}/${f}.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${OUT}" "${OBJ}"/*.o -lcudart -lcuda
echo "built ${OUT}"
