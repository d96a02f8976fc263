#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/ply_ref from oracle/ref_ply.cpp and the reference's own
# vendored tinyply.h (included from where it lies under /root/reference; never copied).  No-op when the
# reference tree is absent.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GSDF_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -f "$REF/submodules/utils/ply_utils/tinyply.h" ]; then
  echo "build_ref_ply.sh: $REF not present; keeping prebuilt $OUT/ply_ref (if any)"; exit 0
fi
if [ -f "$OUT/ply_ref" ] && [ "$OUT/ply_ref" -nt "$HERE/ref_ply.cpp" ]; then echo "build_ref_ply.sh: up to date"; exit 0; fi
g++ -std=c++17 -O2 -w -I"$REF/submodules/utils/ply_utils" "$HERE/ref_ply.cpp" -o "$OUT/ply_ref"
echo "build_ref_ply.sh: built $OUT/ply_ref"
