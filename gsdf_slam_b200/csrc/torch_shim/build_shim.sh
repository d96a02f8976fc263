#!/usr/bin/env bash
# Builds libgsr_torch_shim.so: the LibTorch drop-in shim (same C++ symbols as the reference's
# src/rasterize_points.cu + submodules/simple-knn/spatial.cu) on top of libgsr.so.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../../.." && pwd)"
PKG="$ROOT/gsdf_slam_b200"
OUT="$PKG/libgsr_torch_shim.so"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/rasterize_points.cpp" ] && [ "$OUT" -nt "$HERE/rasterize_points.h" ] && [ "$OUT" -nt "$ROOT/include/gsr_capi.h" ]; then
  echo "build_shim.sh: up to date"; exit 0
fi
PY="${PYTHON:-python}"
TORCH_DIR="$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
CXX11="$($PY -c 'import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')"
g++ -std=c++17 -O2 -fPIC -shared -D_GLIBCXX_USE_CXX11_ABI=$CXX11 \
  -I"$ROOT/include" -I"$HERE" -I"$TORCH_DIR/include" -I"$TORCH_DIR/include/torch/csrc/api/include" \
  -I/usr/local/cuda/include -I"$PYINC" \
  "$HERE/rasterize_points.cpp" -o "$OUT" \
  -L"$PKG" -l:libgsr.so -Wl,-rpath,'$ORIGIN' \
  -L"$TORCH_DIR/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -Wl,-rpath,"$TORCH_DIR/lib"
echo "build_shim.sh: built $OUT"
