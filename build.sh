#!/usr/bin/env bash
# Builds eilev_b200/libvideoblip_b200.so for sm_100a (in-tree, travels with gpurun).
# One nvcc per translation unit, in parallel; objects are rebuilt only when a source is newer.
set -euo pipefail
cd "$(dirname "$0")"
SRC=eilev_b200/csrc
OBJ=build/obj
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC ${NVCC_EXTRA:-}"
UNITS="api gemm_tcgen05 gemm_tcgen05_2cta gemm_generic attention attention_tcgen05 attention_flash_tcgen05 layernorm elementwise decode t5 frames"
newest_hdr=$(ls -t $SRC/*.cuh $SRC/*.h include/*.h build.sh | head -1)
pids=()
for u in $UNITS; do
  o="$OBJ/$u.o"
  if [ ! -f "$o" ] || [ "$SRC/$u.cu" -nt "$o" ] || [ "$newest_hdr" -nt "$o" ]; then
    nvcc $FLAGS -c "$SRC/$u.cu" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=""
for u in $UNITS; do OBJS="$OBJS $OBJ/$u.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a $OBJS -o eilev_b200/libvideoblip_b200.so
