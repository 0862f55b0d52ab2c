#!/usr/bin/env bash
# Round-2 GPU call R: f64 LayerNorm-fold statistics (order-independent), flash forward with exact row maximum.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run r_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -x
grep -E "passed|failed|^E  " gpurun_out/r_kernels.log | head -10
timeout 120 python scripts/micro/vision_repeat.py 2>&1 | grep "^call" | tee gpurun_out/r_vision_repeat.log
timeout 200 python scripts/micro/trainer_vs_autograd.py 2>&1 | grep -E "^eager|^trainer" | cut -c1-120 | tee gpurun_out/r_trainer.log
timeout 120 python scripts/micro/attn_accuracy.py 2>&1 | tee gpurun_out/r_accuracy.log
run r_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/r_models.log | head -10
run r_gemm 300 python scripts/bench_gemm.py vit
grep name gpurun_out/r_gemm.log | cut -c1-200
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run r_bench 300 $B
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run r_bench_old 300 $B
run r_bench2 300 $B
for f in r_bench r_bench_old r_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
