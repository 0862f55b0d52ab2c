#!/usr/bin/env bash
# Round-2 GPU call C: 16-warp GEMM epilogue + MMA-thread polling in the ping-pong attention.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run c_tmem 120 ./build/tmem_bench
cat gpurun_out/c_tmem.log | head -12
run c_kernels 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or attention_tcgen05"
run c_full 600 python -m pytest tests/test_za_fullsize_gpu.py -q
run c_bench_attn 120 python scripts/bench_attn.py
run c_bench_gemm 300 python scripts/bench_gemm.py vit qf.crosskv
cat gpurun_out/c_bench_attn.log gpurun_out/c_bench_gemm.log | grep -v "^$" | cut -c1-220
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run c_bench 300 $B
grep -o '"ms_per_step": [0-9.]*' gpurun_out/c_bench.log | head -1
run c_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches_step.csv python bench.py --profile --no-decode
run c_models 900 python -m pytest tests/test_model_gpu.py -q -x
