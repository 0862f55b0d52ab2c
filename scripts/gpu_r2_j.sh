#!/usr/bin/env bash
# Round-2 GPU call J: attention backward with several key tiles per CTA (Q-Former cross-attention), step A/B.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run j_attn 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention"
run j_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q -x
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run j_bench 300 $B
VB_ATTN_BWD_KT=1 run j_bench_kt1 300 $B
VB_ATTN_BWD_KT=4 run j_bench_kt4 300 $B
VB_ATTN_BWD_KT=16 run j_bench_kt16 300 $B
for f in j_bench j_bench_kt1 j_bench_kt4 j_bench_kt16; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
run j_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02j_launches_step.csv python bench.py --profile --no-decode
