#!/usr/bin/env bash
# Round-2 GPU call B: TMEM / MUFU micro-benchmark, per-kernel launch list of the new step, A/B of the
# round-2 switches, ncu --set full of the GEMM shapes (roofline.traffic) and of the ping-pong attention.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run b_tmem 120 ./build/tmem_bench
cat gpurun_out/b_tmem.log
run b_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python bench.py --profile --no-decode
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run b_bench_all 300 $B
VB_VIT_LN_FOLD=0 run b_bench_nofold 300 $B
VB_ATTN_PP=0 run b_bench_nopp 300 $B
for f in b_bench_all b_bench_nofold b_bench_nopp; do grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1; done
run b_ncu_gemm 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -o gpurun_out/r02_gemm_shapes -f python scripts/profile_gemm_shapes.py
run b_ncu_attn 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05_pp -s 2 -c 1 -o gpurun_out/r02_attn_pp -f python scripts/bench_attn.py
ls -la gpurun_out/*.ncu-rep
