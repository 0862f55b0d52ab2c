#!/usr/bin/env bash
# Round-2 GPU call AQ: the cross-K|V weight gradient (K = 34 952 tokens) with MN-major operands vs two transposes.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
for v in 8192 100000 8192 100000; do VB_GEMM_TN_MAXK=$v timeout 300 $B > gpurun_out/aq_$v.log 2>&1; echo "maxk=$v $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/aq_$v.log | head -1)"; done
VB_GEMM_TN_MAXK=100000 timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "backward or trainer or real_dims" 2>&1 | tail -2
