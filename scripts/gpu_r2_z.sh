#!/usr/bin/env bash
# Round-2 GPU call Z: dropout / relative-bias attention back on mma.sync; step A/B.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run z_attn 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
grep -E "passed|failed|^E  " gpurun_out/z_attn.log | head
run z_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/z_models.log | head
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run z_bench 300 $B
VB_ATTN_TC_SLOW=1 run z_bench_slow 300 $B
run z_bench2 300 $B
run z_t5 300 $B --lm t5
for f in z_bench z_bench_slow z_bench2 z_t5; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
