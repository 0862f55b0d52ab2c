#!/usr/bin/env bash
# Round-2 GPU call AD: greedy generate() with the token bookkeeping inside the decode graph.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run ad_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py tests/test_zz_decode_rows_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/ad_models.log | head
run ad_bench 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-bar
python - <<'PY'
import json,re
s=open('gpurun_out/ad_bench.log').read()
m=re.search(r'\{"metric.*', s)
d=json.loads(m.group(0))
for k in ('decode','decode_batch8'):
    x=d[k]; print(k, 'graph loop', round(x['value'],1), 'tok/s; generate()', round(x['e2e']['value'],1), 'tok/s', 'roofline', round(x['roofline']['frac'],3))
PY
