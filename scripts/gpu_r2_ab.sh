#!/usr/bin/env bash
# Round-2 GPU call AB: residual-dropout mask written by the LayerNorm backward (OPT), step A/B.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run ab_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm or dropout"
grep -E "passed|failed|^E  " gpurun_out/ab_kernels.log | head
run ab_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py tests/test_zc_fulldepth_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/ab_models.log | head
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run ab_bench 300 $B
run ab_bench2 300 $B
for f in ab_bench ab_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/$f.log)"; done
