#!/usr/bin/env bash
# Round-2 GPU call F: attention (QK-first polling, batched loads), GEMM narrow last column block.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run f_kernels 900 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -x -k "gemm or attention"
run f_bench_attn 120 python scripts/bench_attn.py
VB_ATTN_POLL=0 run f_bench_attn_fixed 120 python scripts/bench_attn.py
head -1 gpurun_out/f_bench_attn.log gpurun_out/f_bench_attn_fixed.log
run f_bench_gemm 300 python scripts/bench_gemm.py vit qf.crosskv
grep tflops gpurun_out/f_bench_gemm.log | cut -c1-200
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run f_bench 300 $B
grep -o '"ms_per_step": [0-9.]*' gpurun_out/f_bench.log | head -1
run f_ncu_attn 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05_pp -s 2 -c 1 -o gpurun_out/r02f_attn_pp -f python scripts/bench_attn.py
run f_models 900 python -m pytest tests/test_model_gpu.py -q -x
