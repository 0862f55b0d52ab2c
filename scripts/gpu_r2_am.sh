#!/usr/bin/env bash
# Round-2 GPU call AM: ViT attention with rolled / pipelined softmax passes, 208-register softmax warps and a staged
# bulk tensor store of O: kernel + model tests, bench.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run am_kernels 900 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -x
run am_models 1200 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q -x
run am_attn 120 python scripts/bench_attn.py
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run am_1 300 $B
run am_2 300 $B
for f in am_1 am_2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
grep tcgen05: gpurun_out/am_attn.log
