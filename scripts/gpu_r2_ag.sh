#!/usr/bin/env bash
# Round-2 GPU call AG: branch-free dropout / relative-bias variants of the flash attention.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run ag_attn 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
grep -E "passed|failed|^E  " gpurun_out/ag_attn.log | head
run ag_bench_attn 120 python scripts/bench_attn_bwd.py
grep -E "dropout" gpurun_out/ag_bench_attn.log
VB_ATTN_TC_SLOW=0 run ag_bench_attn_mma 120 python scripts/bench_attn_bwd.py
grep -E "dropout" gpurun_out/ag_bench_attn_mma.log
run ag_models 1200 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py tests/test_zc_fulldepth_gpu.py tests/test_zz_t5_relu_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/ag_models.log | head
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run ag_1 300 $B
VB_ATTN_TC_SLOW=0 run ag_2 300 $B
run ag_3 300 $B
run ag_t5 300 $B --lm t5
VB_ATTN_TC_SLOW=0 run ag_t5_mma 300 $B --lm t5
for f in ag_1 ag_2 ag_3 ag_t5 ag_t5_mma; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
