#!/usr/bin/env bash
# Round-2 GPU call AK: weight gradients straight from the activations (MN-major operands), delta loads batched.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run ak_kernels 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or attention"
grep -E "passed|failed|^E  " gpurun_out/ak_kernels.log | head
run ak_models 1200 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/ak_models.log | head
run ak_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02ak_launches_step.csv python bench.py --profile --no-decode
python scripts/summarize_launches.py gpurun_out/r02ak_launches_step.csv 40 > gpurun_out/r02ak_launches_summary.txt; head -20 gpurun_out/r02ak_launches_summary.txt
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run ak_1 300 $B
VB_GEMM_TN=0 run ak_2 300 $B
run ak_3 300 $B
for f in ak_1 ak_2 ak_3; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
