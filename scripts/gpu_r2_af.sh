#!/usr/bin/env bash
# Round-2 GPU call AF: key-padding mask blocks that are all ones take the term-free variants.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run af_attn 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
grep -E "passed|failed|^E  " gpurun_out/af_attn.log | head
run af_determinism 120 python scripts/micro/attn_determinism.py 3
grep -c "mismatches 0, lse 0, bwd 0" gpurun_out/af_determinism.log
run af_bench_attn 120 python scripts/bench_attn_bwd.py
grep -E "opt self" gpurun_out/af_bench_attn.log
run af_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/af_models.log | head
run af_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02af_launches_step.csv python bench.py --profile --no-decode
python scripts/summarize_launches.py gpurun_out/r02af_launches_step.csv 40 > gpurun_out/r02af_launches_summary.txt; head -12 gpurun_out/r02af_launches_summary.txt
