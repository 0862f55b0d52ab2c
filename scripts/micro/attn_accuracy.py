"""Forward / backward error of the attention kernels against an fp32 evaluation on the same bf16 inputs
(global relative L2), for the kernel choice given by VB_ATTN_FWD_TC / VB_ATTN_BWD_TC."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from eilev_b200 import ops  # noqa: E402


def ref(q, k, v, heads, scale, causal, key_mask):
    b, sq, hd = q.shape
    skv, d = k.shape[1], hd // heads
    qh, kh, vh = (x.view(b, -1, heads, d).transpose(1, 2) for x in (q, k, v))
    s = qh @ kh.transpose(-1, -2) * scale
    if causal:
        i = torch.arange(sq, device=q.device)[:, None]
        j = torch.arange(skv, device=q.device)[None, :]
        s = s.masked_fill(j > i + (skv - sq), float("-inf"))
    if key_mask is not None:
        s = s.masked_fill(key_mask[:, None, None, :] == 0, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.nan_to_num(p @ vh).transpose(1, 2).reshape(b, sq, hd)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


print("VB_ATTN_FWD_TC =", os.environ.get("VB_ATTN_FWD_TC", "on"), " VB_ATTN_BWD_TC =", os.environ.get("VB_ATTN_BWD_TC", "on"))
for name, b, heads, d, sq, skv, causal, masked, qs in [
        ("opt 976 causal", 1, 32, 80, 976, 976, True, False, 1.0), ("opt 976 causal, peaked", 1, 32, 80, 976, 976, True, False, 4.0),
        ("opt 2x200 causal masked", 2, 5, 16, 200, 200, True, True, 2.0), ("cross 32x2056", 17, 12, 64, 32, 2056, False, False, 1.0),
        ("small 2x40 d16 causal masked", 2, 5, 16, 40, 40, True, True, 3.0)]:
    hd = heads * d
    g = torch.Generator(device="cuda").manual_seed(3)
    q = (torch.randn(b, sq, hd, device="cuda", generator=g) * qs).to(torch.bfloat16)
    k = torch.randn(b, skv, hd, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(b, skv, hd, device="cuda", generator=g).to(torch.bfloat16)
    d_o = torch.randn(b, sq, hd, device="cuda", generator=g).to(torch.bfloat16)
    km = None
    valid = torch.ones(b, sq, dtype=torch.bool, device="cuda")
    if masked:
        km = torch.ones(b, skv, dtype=torch.uint8, device="cuda")
        km[0, :7] = 0
        if causal:
            valid[0, :7] = False
    d_o = d_o * valid[:, :, None]
    scale = d ** -0.5
    o, lse = ops.attention(q, k, v, heads, scale, causal=causal, key_mask=km, need_lse=True)
    qf, kf, vf = (x.float().detach().requires_grad_(True) for x in (q, k, v))
    r = ref(qf, kf, vf, heads, scale, causal, km)
    r.backward(d_o.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, causal=causal, key_mask=km)
    print(f"{name}: fwd {rel(o[valid], r[valid]):.3e}  dq {rel(dq[valid], qf.grad[valid]):.3e}  dk {rel(dk, kf.grad):.3e}  dv {rel(dv, vf.grad):.3e}", flush=True)
