#!/usr/bin/env bash
# Round-2 GPU call AW: LayerNorm with one block per row for wide rows: kernel tests, OPT / T5 model tests, timing.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm" 2>&1 | grep -E "passed|failed|Error" | head -3
timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "real_dims or t5 or dropout" 2>&1 | grep -E "passed|failed|Error" | head -3
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from eilev_b200 import ops
def t(fn, n=50):
    for _ in range(5): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
for rows, cols in ((976, 2560), (976, 2048)):
    x = torch.randn(rows, cols, device="cuda").bfloat16(); r = torch.randn_like(x.float()).bfloat16()
    g = torch.randn(cols, device="cuda"); b = torch.randn(cols, device="cuda")
    y, mean, rstd = ops.layernorm(x, g, b, 1e-5, residual=r, save_stats=True)
    dy = torch.randn_like(x.float()).bfloat16(); xin = (x.float() + r.float()).bfloat16()
    dg, db = torch.zeros(cols, device="cuda"), torch.zeros(cols, device="cuda")
    print(rows, cols, "fwd %.1f us" % t(lambda: ops.layernorm(x, g, b, 1e-5, residual=r, save_stats=True)),
          "bwd %.1f us" % t(lambda: ops.layernorm_bwd(dy, xin, g, mean, rstd, dgamma=dg, dbeta=db)))
PY
