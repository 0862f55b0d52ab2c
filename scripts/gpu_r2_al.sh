#!/usr/bin/env bash
# Round-2 GPU call AL: ViT attention with software-pipelined softmax passes and a staged TMA store of O.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run al_vit 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "vit_class or probs_maps"
grep -E "passed|failed|^E  " gpurun_out/al_vit.log | head
run al_full 300 python -m pytest tests/test_za_fullsize_gpu.py -q -x -k "vit_attention"
for f in 1 3 7; do VB_ATTN_PP_FLAGS=$f run al_bench_f$f 120 python scripts/bench_attn.py; done
run al_trace 120 python scripts/micro/pp_trace.py
sed -n 100,175p gpurun_out/al_trace.log
