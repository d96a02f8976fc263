#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libgsdf_ref.so: the UNMODIFIED
# reference CUDA rasterizer + simple-knn compiled for sm_100a from the sources
# where they lie under /root/reference (never copied into this repo), plus the
# C-ABI wrapper oracle/ref_capi.cu.  glm (un-vendored third-party header lib)
# is replaced by oracle/glm_standin.  No --use_fast_math, default -fmad=true:
# the reference's CMakeLists.txt sets no CUDA flags (CMakeLists.txt:59-90).
# Runs only where /root/reference exists (the build container); the GPU box
# uses the prebuilt .so that travels with the repo snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GSDF_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF/cuda_rasterizer" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt $OUT/libgsdf_ref.so (if any)"
  exit 0
fi
NVCC="${NVCC:-nvcc}"
FLAGS=(-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo
       -Xcompiler -fPIC -include cfloat -w
       -I"$HERE/glm_standin" -I"$REF/cuda_rasterizer" -I"$REF/submodules/simple-knn")
SRCS=("$REF/cuda_rasterizer/forward.cu" "$REF/cuda_rasterizer/backward.cu"
      "$REF/cuda_rasterizer/rasterizer_impl.cu"
      "$REF/submodules/simple-knn/simple_knn.cu" "$HERE/ref_capi.cu")
STAMP="$OUT/.stamp"
NEWSTAMP="$(cat "${SRCS[@]}" "$HERE/glm_standin/glm/glm.hpp" "$HERE/build_ref.sh" | sha256sum | cut -d' ' -f1)"
if [ -f "$OUT/libgsdf_ref.so" ] && [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$NEWSTAMP" ]; then
  echo "build_ref.sh: up to date"
  exit 0
fi
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  o="$OUT/$(basename "${s%.cu}").o"
  OBJS+=("$o")
  "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$OUT/libgsdf_ref.so" "${OBJS[@]}" -gencode arch=compute_100a,code=sm_100a
rm -f "${OBJS[@]}"
echo "$NEWSTAMP" > "$STAMP"
echo "build_ref.sh: built $OUT/libgsdf_ref.so"
