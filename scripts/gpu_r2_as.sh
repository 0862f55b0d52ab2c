#!/usr/bin/env bash
# Round-2 GPU call AS: q / k / v weight and bias gradients accumulated straight into adjacent sink views.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -x -k "trainer or backward or ddp or dropout" 2>&1 | tail -2
B="python bench.py --steps 16 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
for i in 1 2; do timeout 300 $B > gpurun_out/as_$i.log 2>&1; echo "run $i $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/as_$i.log | head -1) $(grep -o '"gpu_launches": [0-9]*' gpurun_out/as_$i.log | head -1)"; done
