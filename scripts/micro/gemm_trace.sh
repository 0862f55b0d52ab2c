#!/usr/bin/env bash
# Builds build/libvideoblip_b200_trace.so: the product library with the CTA-pair GEMM compiled with
# -DVB_GEMM_TRACE (per-CTA clock64 stamps of the kernel phases).  Used by scripts/micro/gemm_trace.py.
set -euo pipefail
cd "$(dirname "$0")/../.."
./build.sh
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC -DVB_GEMM_TRACE \
  -c eilev_b200/csrc/gemm_tcgen05_2cta.cu -o build/obj/gemm_tcgen05_2cta_trace.o
OBJS=""
for u in api gemm_tcgen05 gemm_generic attention attention_tcgen05 attention_flash_tcgen05 layernorm elementwise decode t5 frames; do OBJS="$OBJS build/obj/$u.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a $OBJS build/obj/gemm_tcgen05_2cta_trace.o -o build/libvideoblip_b200_trace.so
