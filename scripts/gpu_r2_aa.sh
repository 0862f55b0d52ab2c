#!/usr/bin/env bash
# Round-2 GPU call AA: grouped tile order of the CTA-pair GEMM (fc2 / proj A re-reads); optimizer-step cost.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run aa_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm"
grep -E "passed|failed|^E  " gpurun_out/aa_kernels.log | head
run aa_gemm 300 python scripts/bench_gemm.py vit
grep name gpurun_out/aa_gemm.log | cut -c1-200
VB_GEMM_GROUP_N=0 run aa_gemm_plain 300 python scripts/bench_gemm.py vit
grep name gpurun_out/aa_gemm_plain.log | cut -c1-200
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gemm_tcgen05 --csv --log-file gpurun_out/aa_traffic.csv python scripts/profile_gemm_shapes.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/aa_traffic.csv') if not l.startswith('=='))]
h=rows[0]; i_n=h.index('Metric Name'); i_v=h.index('Metric Value'); i_id=h.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault(r[i_id],{})[r[i_n]]=r[i_v]
for k,v in cur.items(): print(k, v)
PY
B="python bench.py --steps 10 --warmup 3 --no-decode --no-cpu-baseline --no-library-bar"
run aa_bench 300 $B
VB_GEMM_GROUP_N=0 run aa_bench_plain 300 $B
run aa_bench2 300 $B
for f in aa_bench aa_bench_plain aa_bench2; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -1)"; done
run aa_opt_cost 600 python scripts/micro/optimizer_cost.py
grep -E "micro-step|mean" gpurun_out/aa_opt_cost.log | tail -36
