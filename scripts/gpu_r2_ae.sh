#!/usr/bin/env bash
# Round-2 GPU call AE: splice index scans, beam search on the decode graph, dropout attention per shape.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; ( time timeout "$t" "$@" ) > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? : $(tail -n 4 gpurun_out/$name.log | tr '\n' ' ' | cut -c1-200)"; }
run ae_kernels 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "splice or embed"
grep -E "passed|failed|^E  " gpurun_out/ae_kernels.log | head
run ae_models 900 python -m pytest tests/test_model_gpu.py tests/test_v1_gpu.py tests/test_zz_decode_rows_gpu.py -q
grep -E "passed|failed|^E  " gpurun_out/ae_models.log | head
run ae_attn 120 python scripts/bench_attn_bwd.py
tail -4 gpurun_out/ae_attn.log | head -3
VB_ATTN_TC_SLOW=1 run ae_attn_slow 120 python scripts/bench_attn_bwd.py
tail -4 gpurun_out/ae_attn_slow.log | head -3
python - <<'PY'
import time, torch, sys
sys.path.insert(0,'.')
import bench
from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
dev=torch.device('cuda',0)
cfg=bench.full_config(0.0,'opt'); model=bench.build_gpu_model(cfg,dev).eval()
one=bench.synthetic_batch(7); n=int(one['attention_mask'].sum())-bench.TARGET_TOKENS
kw=dict(pixel_values=one['pixel_values'].to(dev), input_ids=one['input_ids'][:, :n].to(dev), attention_mask=one['attention_mask'][:, :n].to(dev), video_input_mask=one['video_input_mask'][:, :n].to(dev))
for label, extra in (('graph', {}), ('eager', {'use_cuda_graph': False})):
    for _ in range(2):
        model.generate(**kw, num_beams=5, max_new_tokens=32, min_new_tokens=32, **extra)
    torch.cuda.synchronize(); t0=time.perf_counter()
    model.generate(**kw, num_beams=5, max_new_tokens=32, min_new_tokens=32, **extra)
    torch.cuda.synchronize(); print('beam search (5 beams, 32 new tokens, full-size opt-2.7b, 16-ctx prompt)', label, round((time.perf_counter()-t0)*1e3,1), 'ms')
PY
