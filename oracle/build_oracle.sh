#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/liboracle_cpu.so from oracle/cpu_ref.cpp (g++, OpenMP).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/liboracle_cpu.so"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/cpu_ref.cpp" ] && [ "$OUT" -nt "$HERE/build_oracle.sh" ]; then
  echo "build_oracle.sh: up to date"; exit 0
fi
# -march=x86-64-v3 (AVX2+FMA) rather than native: the .so is built in the build container
# and travels to the GPU box, whose CPU may differ.
g++ -std=c++17 -O3 -march=x86-64-v3 -ffp-contract=fast -fopenmp -fPIC -shared "$HERE/cpu_ref.cpp" -o "$OUT"
echo "build_oracle.sh: built $OUT"
