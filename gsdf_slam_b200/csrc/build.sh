#!/usr/bin/env bash
# Builds libgsr.so (the C-ABI product library) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-nvcc}"
OUT="${GSR_OUT:-$HERE/../libgsr.so}"
OBJDIR="${GSR_OBJDIR:-$HERE/obj}"
FLAGS=(-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
       -I"$ROOT/include" -I"$HERE" ${GSR_EXTRA_NVCC_FLAGS:-})
SRCS=(gsr_capi.cu densify.cu ingest.cu preprocess.cu radix_sort.cu binning.cu blend_fwd.cu blend_bwd.cu preprocess_bwd.cu adam.cu peer_adam.cu loss.cu knn.cu mapper.cu)
mkdir -p "$OBJDIR"
pids=()
for s in "${SRCS[@]}"; do
  o="$OBJDIR/${s%.cu}.o"
  if [ ! -f "$o" ] || [ "$HERE/$s" -nt "$o" ] || [ -n "$(find "$HERE" "$ROOT/include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" 2>/dev/null | head -1)" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=(); for s in "${SRCS[@]}"; do OBJS+=("$OBJDIR/${s%.cu}.o"); done
"$NVCC" -shared -o "$OUT.tmp" "${OBJS[@]}" -gencode arch=compute_100a,code=sm_100a
mv -f "$OUT.tmp" "$OUT"   # atomic: a concurrent snapshot never sees a partial library
echo "built $OUT"
