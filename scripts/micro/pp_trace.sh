#!/usr/bin/env bash
# Builds build/libvideoblip_b200_pptrace.so: the product library with the ViT attention compiled with
# -DVB_PP_TRACE (clock64 stamps of CTA 0's MMA issuer and one warp per softmax group).  Used by scripts/micro/pp_trace.py.
set -euo pipefail
cd "$(dirname "$0")/../.."
./build.sh
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC -DVB_PP_TRACE \
  -c eilev_b200/csrc/attention_tcgen05.cu -o build/obj/attention_tcgen05_trace.o
OBJS=""
for u in api gemm_tcgen05 gemm_tcgen05_2cta gemm_generic attention attention_flash_tcgen05 layernorm elementwise decode t5 frames; do OBJS="$OBJS build/obj/$u.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a $OBJS build/obj/attention_tcgen05_trace.o -o build/libvideoblip_b200_pptrace.so
