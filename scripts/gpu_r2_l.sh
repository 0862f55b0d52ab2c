#!/usr/bin/env bash
# Round-2 GPU call L: run-time tile widths of the CTA-pair GEMM (wave-filling widths for the OPT shapes).
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run l_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -k "gemm"
run l_sweep 600 python scripts/bench_gemm.py --sweep --opt
cat gpurun_out/l_sweep.log | head -20
run l_gemm 300 python scripts/bench_gemm.py
grep name gpurun_out/l_gemm.log | cut -c1-220
run l_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q -x
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run l_bench 300 $B
run l_bench2 300 $B
for f in l_bench l_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
run l_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02l_launches_step.csv python bench.py --profile --no-decode
