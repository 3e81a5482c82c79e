#!/usr/bin/env bash
# Builds libb2s.so (sm_100a only) next to the Python package. Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out" "$here/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v "$@")
pids=()
for f in api conv_tc conv2_tc wgrad2_tc elementwise vnet_ops attn_ops; do
  ( "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$here/obj/$f.o" > "$here/obj/$f.log" 2>&1 || { cat "$here/obj/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libb2s.so" "$here/obj/api.o" "$here/obj/conv_tc.o" "$here/obj/conv2_tc.o" "$here/obj/wgrad2_tc.o" "$here/obj/elementwise.o" "$here/obj/vnet_ops.o" "$here/obj/attn_ops.o" -cudart static
echo "built $out/libb2s.so"
