"""One LayerNorm forward + backward on the OPT shape (976 x 2560) for an ncu launch list:
ncu --metrics gpu__time_duration.sum -k regex:ln_ python scripts/micro/ln_once.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from eilev_b200 import ops  # noqa: E402

rows, cols = 976, 2560
x = torch.randn(rows, cols, device="cuda").bfloat16()
r = torch.randn(rows, cols, device="cuda").bfloat16()
g, b = torch.randn(cols, device="cuda"), torch.randn(cols, device="cuda")
dy = torch.randn(rows, cols, device="cuda").bfloat16()
dg, db = torch.zeros(cols, device="cuda"), torch.zeros(cols, device="cuda")
for _ in range(4):
    y, mean, rstd = ops.layernorm(x, g, b, 1e-5, residual=r, save_stats=True)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dgamma=dg, dbeta=db)
torch.cuda.synchronize()
