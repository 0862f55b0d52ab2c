#!/usr/bin/env bash
# Round-2 GPU call AO: per-device launch state + tensor-map cache: full GPU suite, bench, eager generate() timing.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run ao_pytest 1500 python -m pytest tests -m gpu -q -x
grep -E "passed|failed|^E  " gpurun_out/ao_pytest.log | head
run ao_smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run ao_bench 600 python bench.py --no-cpu-baseline --no-library-bar
grep -o '"ms_per_step": [0-9.]*' gpurun_out/ao_bench.log | head -2
grep -o '"decode": {[^}]*}' gpurun_out/ao_bench.log | cut -c1-200
