#!/usr/bin/env bash
# Round-2 GPU call V: OPT-shape width sweep after the issue fix; step launch list.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run v_sweep 600 python scripts/bench_gemm.py --sweep --opt
cat gpurun_out/v_sweep.log | head -8
run v_sweep_vit 600 python scripts/bench_gemm.py --sweep
cat gpurun_out/v_sweep_vit.log | head -8
run v_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02v_launches_step.csv python bench.py --profile --no-decode
python scripts/summarize_launches.py gpurun_out/r02v_launches_step.csv | head -40
