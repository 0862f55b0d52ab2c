#!/usr/bin/env bash
# Round-2 GPU call AC: A/B of the fused LayerNorm-backward dropout output.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; }
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run ac_1 300 $B
VB_OPT_FUSED_LN_DROP=0 run ac_2 300 $B
run ac_3 300 $B
VB_OPT_FUSED_LN_DROP=0 run ac_4 300 $B
for f in ac_1 ac_2 ac_3 ac_4; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
