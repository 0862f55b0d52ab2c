#!/usr/bin/env bash
# Round-2 GPU call T: flash attention with branch-free elementwise variants.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run t_attn 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
grep -E "passed|failed|^E  " gpurun_out/t_attn.log | head -10
run t_determinism 120 python scripts/micro/attn_determinism.py 5
grep -c "mismatches 0, lse 0, bwd 0" gpurun_out/t_determinism.log
timeout 120 python scripts/micro/attn_accuracy.py 2>&1 | head -4 | tee gpurun_out/t_accuracy.log
run t_attn_bench 120 python scripts/bench_attn_bwd.py
cat gpurun_out/t_attn_bench.log | head -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_attn_launches.csv python scripts/micro/attn_once.py > gpurun_out/t_attn_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/t_attn_launches.csv | head -8
bash scripts/micro/attn_trace.sh > /dev/null 2>&1; timeout 120 python scripts/micro/attn_trace.py > gpurun_out/t_attn_trace.log 2>&1
run t_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/t_models.log | head -10
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run t_bench 300 $B
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run t_bench_old 300 $B
run t_bench2 300 $B
for f in t_bench t_bench_old t_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
