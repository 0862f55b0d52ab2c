#!/usr/bin/env bash
# Round-2 GPU call W: width-144 pair tiles for long-K narrow GEMMs, masked forward statistics without branches.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run w_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -x
grep -E "passed|failed|^E  " gpurun_out/w_kernels.log | head -10
run w_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/w_models.log | head -10
run w_gemm 300 python scripts/bench_gemm.py opt.
grep name gpurun_out/w_gemm.log | cut -c1-200
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run w_bench 300 $B
VB_ATTN_BWD_TC=0 VB_ATTN_FWD_TC=0 run w_bench_old 300 $B
run w_bench2 300 $B
for f in w_bench w_bench_old w_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
run w_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02w_launches_step.csv python bench.py --profile --no-decode
python scripts/summarize_launches.py gpurun_out/r02w_launches_step.csv | head -12
