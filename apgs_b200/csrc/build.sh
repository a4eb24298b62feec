#!/usr/bin/env bash
# Builds apgs_b200/lib/libapgs_b200.so for sm_100a.  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=("$here"/*.cu)
"$NVCC" -std=c++17 -O3 -lineinfo -shared -Xcompiler -fPIC,-fvisibility=hidden \
    -gencode arch=compute_100a,code=sm_100a \
    -Xptxas -v ${APGS_NVCC_EXTRA:-} \
    -o "$out/libapgs_b200.so" "${srcs[@]}" 2> "$out/ptxas.log" || { cat "$out/ptxas.log" >&2; exit 1; }
grep -E "error|warning" "$out/ptxas.log" | grep -v "^ptxas info" >&2 || true
echo "built $out/libapgs_b200.so"
