"""One forward + backward of the OPT self-attention shape (for ncu captures of the flash kernels)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from eilev_b200 import ops  # noqa: E402

b, heads, d, sq, skv, causal = 1, 32, 80, 976, 976, True
hd = heads * d
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(b, sq, 3 * hd, device="cuda", generator=g).to(torch.bfloat16)
q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
d_o = torch.randn(b, sq, hd, device="cuda", generator=g).to(torch.bfloat16)
for _ in range(2):
    o, lse = ops.attention(q, k, v, heads, d ** -0.5, causal=causal, need_lse=True)
    ops.attention_bwd(q, k, v, o, lse, d_o, heads, d ** -0.5, causal=causal)
torch.cuda.synchronize()
torch.cuda.profiler.start()
o, lse = ops.attention(q, k, v, heads, d ** -0.5, causal=causal, need_lse=True)
ops.attention_bwd(q, k, v, o, lse, d_o, heads, d ** -0.5, causal=causal)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
