#!/usr/bin/env bash
# Builds apgs_b200/lib/libapgs_b200.so for sm_100a.  nvcc cross-compiles without a GPU.
# One object per source, compiled in parallel; a source is recompiled when it or any header is newer than its object.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
obj="$here/../../gpurun_out/.build_obj"   # scratch: git-ignored and never sent to the GPU box
mkdir -p "$out" "$obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
flags=(-std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -gencode arch=compute_100a,code=sm_100a -Xptxas -v ${APGS_NVCC_EXTRA:-})
newest_hdr="$(ls -t "$here"/*.cuh "$here"/../../include/*.h "$here"/build.sh | head -1)"
echo "${flags[*]}" > "$obj/flags.new"
if ! cmp -s "$obj/flags.new" "$obj/flags" 2>/dev/null; then rm -f "$obj"/*.o; mv "$obj/flags.new" "$obj/flags"; fi
pids=()
names=()
for src in "$here"/*.cu; do
    n="$(basename "$src" .cu)"
    if [[ ! -f "$obj/$n.o" || "$src" -nt "$obj/$n.o" || "$newest_hdr" -nt "$obj/$n.o" ]]; then
        ( "$NVCC" "${flags[@]}" -c "$src" -o "$obj/$n.o" 2> "$obj/$n.log" ) &
        pids+=($!)
        names+=("$n")
    fi
done
fail=0
for i in "${!pids[@]}"; do
    if ! wait "${pids[$i]}"; then cat "$obj/${names[$i]}.log" >&2; rm -f "$obj/${names[$i]}.o"; fail=1; fi
done
[[ $fail == 0 ]] || exit 1
cat "$obj"/*.log > "$out/ptxas.log"
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$out/libapgs_b200.so" "$obj"/*.o
grep -E "error|warning" "$out/ptxas.log" | grep -v "^ptxas info" >&2 || true
echo "built $out/libapgs_b200.so"
