#!/usr/bin/env bash
# Builds eilev_b200/libvideoblip_b200.so for sm_100a (in-tree, travels with gpurun).
set -euo pipefail
cd "$(dirname "$0")"
SRC=eilev_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --use_fast_math \
  -Xcompiler -fPIC -shared ${NVCC_EXTRA:-} \
  $SRC/api.cu $SRC/gemm_tcgen05.cu $SRC/gemm_tcgen05_2cta.cu $SRC/gemm_generic.cu $SRC/attention.cu $SRC/attention_tcgen05.cu $SRC/layernorm.cu \
  $SRC/elementwise.cu $SRC/decode.cu -o eilev_b200/libvideoblip_b200.so
