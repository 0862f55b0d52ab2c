#!/usr/bin/env bash
# Round-2 GPU call AX: the full GPU suite and a bench line on the final state.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/ax_pytest.log 2>&1; grep -E "passed|failed|^E  |^FAILED" gpurun_out/ax_pytest.log | head -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu-baseline --no-library-bar --no-decode > gpurun_out/ax_bench.log 2>&1
grep -o '{"metric.*' gpurun_out/ax_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'roofline', round(d['roofline']['frac'],3), d['clocks'])"
