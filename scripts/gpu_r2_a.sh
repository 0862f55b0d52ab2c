#!/usr/bin/env bash
# Round-2 GPU call A: first hardware run of the ping-pong ViT attention, the folded LayerNorm, the new GELU
# epilogue, then the whole GPU suite, the full-depth parity tests and the bench.  Logs in gpurun_out/.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, timeout, command...
  local name=$1 t=$2; shift 2
  ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-400)"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
# 1. does the P.V instruction take N = round_up(d, 16) with the MN-major 128B-swizzled V tile (old kernel)?
VB_ATTN_PP=0 VB_ATTN_PV_NPAD=1 run a_pv_npad 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_vit_class"
# 2. the new kernels, each family in its own process (a trap poisons the context)
run a_attn_pp 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention_tcgen05_vit_class"
run a_attn_full 300 python -m pytest tests/test_za_fullsize_gpu.py -q -k "attention"
run a_gemm_new 600 python -m pytest tests/test_kernels_gpu.py -q -k "layernorm_fold or row_statistics or gelu or epilogues"
# 3. micro-benchmarks
VB_ATTN_PP=0 run a_bench_attn_old 120 python scripts/bench_attn.py
run a_bench_attn_pp 120 python scripts/bench_attn.py
run a_bench_gemm 300 python scripts/bench_gemm.py vit
# 4. everything
run a_pytest_gpu 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_zc_fulldepth_gpu.py
run a_fulldepth 1200 python -m pytest tests/test_zc_fulldepth_gpu.py -q
# 5. bench
run a_bench 900 python bench.py --steps 20 --warmup 5
tail -n 1 gpurun_out/a_bench.log | cut -c1-3000
