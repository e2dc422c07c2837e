#!/usr/bin/env bash
# Builds graphical-gan_b200/lib/libgg_b200.so (sm_100a only) from csrc/*.cu.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/lib"
OBJ="$HERE/build"
mkdir -p "$OUT" "$OBJ"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --use_fast_math=false)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC)
[ "${GG_PTXAS_V:-0}" = "1" ] && FLAGS+=(-Xptxas -v)
[ -n "${GG_EXTRA_FLAGS:-}" ] && FLAGS+=(${GG_EXTRA_FLAGS})
pids=()
objs=()
for src in "$HERE"/csrc/*.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || [ -n "$(find "$HERE/csrc" "$HERE/../include" -name '*.cuh' -newer "$o" -o -name '*.h' -newer "$o")" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a "${objs[@]}" -o "$OUT/libgg_b200.so" -cudart static
echo "built $OUT/libgg_b200.so"
if [ "${GG_BUILD_TIMELINE:-0}" = "1" ]; then
  # instrumented twin for tools/timeline_conv.py (GG_LIB=.../libgg_b200_tl.so): per-stage globaltimer stamps in the conv kernel
  "$NVCC" "${FLAGS[@]}" -DGG_TIMELINE -c "$HERE/csrc/gg_conv_tc.cu" -o "$OBJ/gg_conv_tc_tl.o"
  tl_objs=()
  for o in "${objs[@]}"; do
    if [ "$(basename "$o")" = "gg_conv_tc.o" ]; then tl_objs+=("$OBJ/gg_conv_tc_tl.o"); else tl_objs+=("$o"); fi
  done
  "$NVCC" -shared -gencode arch=compute_100a,code=sm_100a "${tl_objs[@]}" -o "$OUT/libgg_b200_tl.so" -cudart static
  echo "built $OUT/libgg_b200_tl.so"
fi
