#!/usr/bin/env bash
# Round-2 GPU call Y: T5 step with / without the tcgen05 flash attention (relative bias = per-element path).
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
B="python bench.py --lm t5 --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run y_t5 300 $B
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run y_t5_old 300 $B
VB_ATTN_FWD_TC=0 run y_t5_bwdonly 300 $B
run y_t5_2 300 $B
for f in y_t5 y_t5_old y_t5_bwdonly y_t5_2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
