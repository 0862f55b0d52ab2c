#!/usr/bin/env bash
# First GPU call of the next round: runs the GPU tests that were written after round 1's
# GPU budget was spent, then the frame-path micro-benchmark and the uint8 bench variant.
#   gpurun --timeout 1500 -- 'bash scripts/verify_pending.sh'
# Everything lands in gpurun_out/ (merged back by gpurun).
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in test_za_fullsize_gpu test_zb_v1_real_dims_gpu test_zz_decode_rows_gpu test_zz_frames_gpu test_zz_t5_relu_gpu; do
  timeout 600 python -m pytest "tests/$t.py" -q 2>&1 | tail -25 > "gpurun_out/pending_$t.log"
  tail -3 "gpurun_out/pending_$t.log"
done
timeout 300 python scripts/bench_frames.py > gpurun_out/bench_frames.json 2> gpurun_out/bench_frames.err
cat gpurun_out/bench_frames.json
# launch list of the frame kernels (per-launch times under ncu are cold-cache: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
  --log-file gpurun_out/frames_launches.csv python scripts/bench_frames.py > /dev/null 2> gpurun_out/frames_ncu.err
timeout 900 python bench.py --u8-frames --no-decode --no-cpu-baseline > gpurun_out/bench_u8_frames.json 2> gpurun_out/bench_u8_frames.err
cut -c1-600 gpurun_out/bench_u8_frames.json
