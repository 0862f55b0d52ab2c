#!/usr/bin/env bash
# Round-2 GPU call E: attention with L2 prefetch (fixed order vs polling), frame transform kernel tests.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run e_attn 300 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -k "attention"
VB_ATTN_POLL=1 run e_attn_poll 300 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -k "attention"
run e_bench_attn 120 python scripts/bench_attn.py
VB_ATTN_POLL=1 run e_bench_attn_poll 120 python scripts/bench_attn.py
head -1 gpurun_out/e_bench_attn.log gpurun_out/e_bench_attn_poll.log
run e_frames 300 python -m pytest tests/test_ze_preprocess_gpu.py tests/test_zz_frames_gpu.py -q
run e_ncu_attn 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05_pp -s 2 -c 1 -o gpurun_out/r02e_attn_pp -f python scripts/bench_attn.py
