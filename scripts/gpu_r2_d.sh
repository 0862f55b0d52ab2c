#!/usr/bin/env bash
# Round-2 GPU call D: attention polling with yield, T5 dropout test, whole suite, bench.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run d_attn 300 python -m pytest tests/test_kernels_gpu.py tests/test_za_fullsize_gpu.py -q -k "attention"
run d_bench_attn 120 python scripts/bench_attn.py
head -2 gpurun_out/d_bench_attn.log
run d_pytest_gpu 1500 python -m pytest tests -q -m gpu --deselect tests/test_zc_fulldepth_gpu.py
run d_fulldepth 1200 python -m pytest tests/test_zc_fulldepth_gpu.py -q
run d_ncu_attn 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05_pp -s 2 -c 1 -o gpurun_out/r02d_attn_pp -f python scripts/bench_attn.py
run d_bench 900 python bench.py --steps 20 --warmup 5
grep "^{" gpurun_out/d_bench.log | cut -c1-400
run d_bench_t5 900 python bench.py --steps 10 --warmup 3 --lm t5 --no-cpu-baseline
grep "^{" gpurun_out/d_bench_t5.log | cut -c1-400
