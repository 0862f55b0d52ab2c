#!/usr/bin/env bash
# Round-2 GPU call AP (2 GPUs): the 2-rank NCCL gradient-equality test (x3) and the bench line at N = 2.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "adamw or sumsq or optimizer" ) > gpurun_out/ap_sumsq.log 2>&1; tail -3 gpurun_out/ap_sumsq.log
for i in 1 2 3; do ( time timeout 600 python -m pytest tests/test_zd_nccl_gpu.py -q -x ) > gpurun_out/ap_nccl_$i.log 2>&1; grep -E "passed|failed|^E  " gpurun_out/ap_nccl_$i.log | head -5; done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 --no-decode ) > gpurun_out/ap_bench_n2.log 2>&1
grep -o '{"metric.*' gpurun_out/ap_bench_n2.log > gpurun_out/r02_bench_n2.json; cut -c1-200 gpurun_out/r02_bench_n2.json
