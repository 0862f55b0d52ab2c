#!/usr/bin/env bash
# Round-2 GPU call AH: branch-free causal-diagonal variants in the flash attention.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run ah_attn 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
grep -E "passed|failed|^E  " gpurun_out/ah_attn.log | head
run ah_determinism 120 python scripts/micro/attn_determinism.py 3
grep -c "mismatches 0, lse 0, bwd 0" gpurun_out/ah_determinism.log
timeout 120 python scripts/micro/attn_accuracy.py 2>&1 | tee gpurun_out/ah_accuracy.log
run ah_bench_attn 120 python scripts/bench_attn_bwd.py
grep -v "^real\|^user\|^sys\|^$" gpurun_out/ah_bench_attn.log
run ah_models 1200 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py tests/test_zc_fulldepth_gpu.py tests/test_zz_t5_relu_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/ah_models.log | head
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run ah_1 300 $B
run ah_2 300 $B
run ah_t5 300 $B --lm t5
for f in ah_1 ah_2 ah_t5; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
