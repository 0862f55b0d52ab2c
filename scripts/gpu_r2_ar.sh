#!/usr/bin/env bash
# Round-2 GPU call AR: no zero fill of the cross dK|dV buffer: gradient tests, full-depth parity, bench.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_zc_fulldepth_gpu.py tests/test_v1_gpu.py -q -x 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
for i in 1 2; do timeout 300 $B > gpurun_out/ar_$i.log 2>&1; echo "run $i $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ar_$i.log | head -1)"; done
