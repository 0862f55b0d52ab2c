#!/usr/bin/env bash
# Round-2 GPU call AJ: delta inside the dQ pass; wider column-sum / LayerNorm parameter-gradient kernels.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run aj_kernels 900 python -m pytest tests/test_kernels_gpu.py -q -x
grep -E "passed|failed|^E  " gpurun_out/aj_kernels.log | head
run aj_determinism 120 python scripts/micro/attn_determinism.py 3
grep -c "mismatches 0, lse 0, bwd 0" gpurun_out/aj_determinism.log
timeout 120 python scripts/micro/attn_accuracy.py 2>&1 | head -4
run aj_bench_attn 120 python scripts/bench_attn_bwd.py
grep -v "^real\|^user\|^sys\|^$\|warm-up" gpurun_out/aj_bench_attn.log | head -6
run aj_models 1200 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py tests/test_zc_fulldepth_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/aj_models.log | head
run aj_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02aj_launches_step.csv python bench.py --profile --no-decode
python scripts/summarize_launches.py gpurun_out/r02aj_launches_step.csv 40 > gpurun_out/r02aj_launches_summary.txt; head -24 gpurun_out/r02aj_launches_summary.txt
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run aj_1 300 $B
run aj_2 300 $B
for f in aj_1 aj_2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
