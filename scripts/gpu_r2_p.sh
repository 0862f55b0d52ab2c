#!/usr/bin/env bash
# Round-2 GPU call P: flash attention with 64-column items and a 4-deep block ring.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run p_attn 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
grep -E "passed|failed|Error|error|assert" gpurun_out/p_attn.log | head -20
run p_determinism 120 python scripts/micro/attn_determinism.py 5
cat gpurun_out/p_determinism.log | head -12
for e in "X=1" "VB_PDL=0" "VB_ATTN_FWD_TC=0" "VB_VIT_LN_FOLD=0"; do echo "== vision repeat $e"; env $e timeout 120 python scripts/micro/vision_repeat.py 2>&1 | grep "^call" ; done 2>&1 | tee gpurun_out/p_vision_repeat.log
run p_attn_bench 120 python scripts/bench_attn_bwd.py
cat gpurun_out/p_attn_bench.log | head -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_attn_launches.csv python scripts/bench_attn_bwd.py > gpurun_out/p_attn_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/p_attn_launches.csv | head -14
run p_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run p_bench 300 $B
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run p_bench_old 300 $B
run p_bench2 300 $B
for f in p_bench p_bench_old p_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
