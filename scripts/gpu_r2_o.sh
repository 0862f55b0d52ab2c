#!/usr/bin/env bash
# Round-2 GPU call O: tcgen05 flash attention, forward + backward, two 64-column halves per tile.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run o_attn 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
grep -E "passed|failed|Error|error|assert" gpurun_out/o_attn.log | head -20
run o_attn_bench 120 python scripts/bench_attn_bwd.py
cat gpurun_out/o_attn_bench.log
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run o_attn_bench_old 120 python scripts/bench_attn_bwd.py
cat gpurun_out/o_attn_bench_old.log
run o_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q -x
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run o_bench 300 $B
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run o_bench_old 300 $B
run o_bench2 300 $B
for f in o_bench o_bench_old o_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
