#!/usr/bin/env bash
# Round-2 GPU call K: activation backward in the dgrad GEMM epilogue; K/V prefetch in the keep-dQ attention backward.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run k_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -k "gemm or attention"
run k_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q -x
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run k_bench 300 $B
VB_ACT_BWD_FUSED=0 run k_bench_unfused 300 $B
VB_ATTN_BWD_KT=1 run k_bench_kt1 300 $B
run k_bench2 300 $B
for f in k_bench k_bench_unfused k_bench_kt1 k_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
run k_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02k_launches_step.csv python bench.py --profile --no-decode
