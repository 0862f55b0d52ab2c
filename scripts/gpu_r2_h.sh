#!/usr/bin/env bash
# Round-2 GPU call H: 257th query row on warps 2-3 (A/B VB_ATTN_ROW256), attention maps, step A/B.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run h_attn 300 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -k "attention"
run h_bench_attn 120 python scripts/bench_attn.py
VB_ATTN_ROW256=0 run h_bench_attn_3tiles 120 python scripts/bench_attn.py
head -1 gpurun_out/h_bench_attn.log gpurun_out/h_bench_attn_3tiles.log
run h_models 900 python -m pytest tests/test_model_gpu.py -q -x
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run h_bench 300 $B
VB_ATTN_ROW256=0 run h_bench_3tiles 300 $B
VB_VIT_LN_FOLD=0 run h_bench_nofold 300 $B
for f in h_bench h_bench_3tiles h_bench_nofold; do grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; done
run h_ncu_attn 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05_pp -s 2 -c 1 -o gpurun_out/r02h_attn_pp -f python scripts/bench_attn.py
