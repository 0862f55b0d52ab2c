#!/usr/bin/env bash
# Round-2 GPU call Q: which attention kernel moves the trainer-graph and beam-search tests; accuracy of both.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="tests/test_model_gpu.py::test_trainer_graphs_and_inplace_grad_accumulation_match_plain_autograd tests/test_model_gpu.py::test_beam_search_and_repetition_penalty_match_reference_golden"
for e in "X=1" "VB_ATTN_FWD_TC=0" "VB_ATTN_BWD_TC=0" "VB_ATTN_FWD_TC=0 VB_ATTN_BWD_TC=0"; do
  echo "== $e"; env $e timeout 300 python -m pytest $T -q 2>&1 | grep -E "^E   |passed|failed" | cut -c1-200 | head -8
done 2>&1 | tee gpurun_out/q_tests.log
for e in "X=1" "VB_ATTN_FWD_TC=0 VB_ATTN_BWD_TC=0"; do env $e timeout 120 python scripts/micro/attn_accuracy.py; done 2>&1 | tee gpurun_out/q_accuracy.log
