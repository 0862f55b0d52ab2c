#!/usr/bin/env bash
# Round-2 GPU call AT: last check (warning test, model tests) and the headline bench line with the 3-step roofline.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | grep -E "passed|failed|Error" | head -3
timeout 900 python bench.py > gpurun_out/at_bench.log 2>&1
grep -o '{"metric.*' gpurun_out/at_bench.log > gpurun_out/r02_bench_n1.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"],2), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), "roofline", round(d["roofline"]["frac"],3), d["roofline"]["launches"], "lib", round(d["library_bar"]["value"],2), "cpu", round(d["cpu_baseline"]["value"],3))
PY
timeout 300 python bench.py --lm t5 --no-decode > gpurun_out/at_bench_t5.log 2>&1; grep -o '"achieved": [0-9.]*' gpurun_out/at_bench_t5.log | head -1
