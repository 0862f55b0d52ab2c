#!/usr/bin/env bash
# Round-2 GPU call AN: the final artefacts of the round (full GPU suite, bench lines, ncu summaries, launch list).
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/art
A=gpurun_out/art
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "$A/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 $A/$name.log | tr '\n' ' ' | cut -c1-300)"; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
grep -E "passed|failed|^E  " $A/pytest_gpu.log | head -10
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_n1 900 python bench.py
grep -o '{"metric.*' $A/bench_n1.log > $A/r02_bench_n1.json
run bench_t5 600 python bench.py --lm t5
grep -o '{"metric.*' $A/bench_t5.log > $A/r02_bench_t5_n1.json
run bench_ref 900 python bench.py --impl reference --steps 1 --warmup 0
grep -o '{".*' $A/bench_ref.log | tail -1 > $A/r02_bench_reference_arm.json
# ncu: GEMM shapes (traffic), ViT attention, flash attention
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -o $A/r02_gemm_shapes python scripts/profile_gemm_shapes.py > $A/ncu_gemm.log 2>&1
python scripts/gemm_traffic.py $A/r02_gemm_shapes.ncu-rep $A/r02_ncu_gemm_traffic.json > $A/gemm_traffic.log 2>&1
python scripts/ncu_summary.py $A/r02_gemm_shapes.ncu-rep > $A/r02_ncu_gemm_shapes.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05_pp -s 3 -c 1 -o $A/r02_attn_pp python scripts/bench_attn.py > $A/ncu_attn_pp.log 2>&1
python scripts/ncu_summary.py $A/r02_attn_pp.ncu-rep > $A/r02_ncu_attn_tcgen05_vit.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_flash -o $A/r02_attn_flash python scripts/micro/attn_once.py > $A/ncu_attn_flash.log 2>&1
python scripts/ncu_summary.py $A/r02_attn_flash.ncu-rep > $A/r02_ncu_attn_flash.txt 2>&1
run launches 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $A/r02_launches_step.csv python bench.py --profile --no-decode
python scripts/summarize_launches.py $A/r02_launches_step.csv 60 > $A/r02_launches_step_summary.txt 2>&1
python scripts/ncu_source_summary.py $A/r02_attn_pp.ncu-rep 2000 > $A/r02_ncu_attn_pp_source.txt 2>&1
VB_LIB_PATH=build/libvideoblip_b200_pptrace.so timeout 120 python scripts/micro/pp_trace.py > $A/r02_attn_pp_trace.txt 2>&1
timeout 60 ./build/pipe_bench > $A/r02_pipe_bench.txt 2>&1
run attn_bench 120 python scripts/bench_attn_bwd.py
run attn_vit 120 python scripts/bench_attn.py
run gemm_bench 300 python scripts/bench_gemm.py
cp gpurun_out/parity_report*.json $A/ 2>/dev/null
rm -f $A/*.ncu-rep.tmp; ls -la $A | head -50
