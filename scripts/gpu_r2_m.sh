#!/usr/bin/env bash
# Round-2 GPU call M: CTA-pair GEMM with run-time smem plan (deeper ring for narrow tiles) and L2 prefetch.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run m_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -k "gemm"
run m_trace 300 python scripts/micro/gemm_trace.py
cat gpurun_out/m_trace.log
for a in 0 6 12 20 32; do
  VB_GEMM_L2_AHEAD=$a run m_gemm_a$a 300 python scripts/bench_gemm.py opt.
  echo "ahead=$a"; grep name gpurun_out/m_gemm_a$a.log | cut -c1-160
done
run m_sweep 600 python scripts/bench_gemm.py --sweep --opt
cat gpurun_out/m_sweep.log | head -8
run m_gemm 300 python scripts/bench_gemm.py vit qf
grep name gpurun_out/m_gemm.log | cut -c1-220
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run m_bench 300 $B
VB_GEMM_L2_AHEAD=0 run m_bench_a0 300 $B
run m_bench2 300 $B
for f in m_bench m_bench_a0 m_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
