#!/usr/bin/env bash
# Round-2 GPU call G: programmatic dependent launch across the training step (A/B with VB_PDL=0), whole suite.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run g_pytest_gpu 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_zc_fulldepth_gpu.py
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run g_bench_pdl 300 $B
VB_PDL=0 run g_bench_nopdl 300 $B
for f in g_bench_pdl g_bench_nopdl; do grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; done
run g_bench_attn 120 python scripts/bench_attn.py
head -1 gpurun_out/g_bench_attn.log
run g_bench_gemm 300 python scripts/bench_gemm.py vit qf.crosskv
grep tflops gpurun_out/g_bench_gemm.log | cut -c1-200
run g_fulldepth 1200 python -m pytest tests/test_zc_fulldepth_gpu.py -q
