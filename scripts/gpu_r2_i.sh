#!/usr/bin/env bash
# Round-2 GPU call I: phase offset between the two softmax warp groups (VB_ATTN_STAGGER sweep).
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run i_attn 300 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -k "attention"
for st in 0 2000 3500 4500 6000 8000; do
  VB_ATTN_STAGGER=$st run i_bench_attn_$st 120 python scripts/bench_attn.py
  echo "stagger $st: $(head -1 gpurun_out/i_bench_attn_$st.log)"
done
run i_models 900 python -m pytest tests/test_model_gpu.py -q -x
run i_ncu_attn 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05_pp -s 2 -c 1 -o gpurun_out/r02i_attn_pp -f python scripts/bench_attn.py
