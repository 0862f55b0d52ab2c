#!/usr/bin/env bash
# Round-2 GPU call U: single-thread TMA / MMA issue under elect.sync (no waterfall loops) in every tcgen05 kernel.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run u_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -x
grep -E "passed|failed|^E  " gpurun_out/u_kernels.log | head -10
run u_gemm 300 python scripts/bench_gemm.py
grep name gpurun_out/u_gemm.log | cut -c1-200
run u_attn_vit 120 python scripts/bench_attn.py
cat gpurun_out/u_attn_vit.log | head -3
run u_attn_bench 120 python scripts/bench_attn_bwd.py
cat gpurun_out/u_attn_bench.log | head -6
run u_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/u_models.log | head -10
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run u_bench 300 $B
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run u_bench_old 300 $B
run u_bench2 300 $B
for f in u_bench u_bench_old u_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
bash scripts/micro/attn_trace.sh > /dev/null 2>&1; timeout 120 python scripts/micro/attn_trace.py > gpurun_out/u_attn_trace.log 2>&1
