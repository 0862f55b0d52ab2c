#!/usr/bin/env bash
# Builds build/libvideoblip_b200_fatrace.so: the product library with the flash attention compiled with
# -DVB_FA_TRACE (clock64 stamps of CTA 0's MMA issuer and one warp per group).  Used by scripts/micro/attn_trace.py.
set -euo pipefail
cd "$(dirname "$0")/../.."
./build.sh
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC -DVB_FA_TRACE \
  -c eilev_b200/csrc/attention_flash_tcgen05.cu -o build/obj/attention_flash_tcgen05_trace.o
OBJS=""
for u in api gemm_tcgen05 gemm_tcgen05_2cta gemm_generic attention attention_tcgen05 layernorm elementwise decode t5 frames; do OBJS="$OBJS build/obj/$u.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a $OBJS build/obj/attention_flash_tcgen05_trace.o -o build/libvideoblip_b200_fatrace.so
