"""Attention backward micro-benchmark on the step's shapes: OPT causal self-attention (1 x 32 heads x 976 x d 80)
and the Q-Former cross-attention (17 x 12 heads, 32 queries over 2 056 keys, d 64).  Run once per kernel choice
(VB_ATTN_BWD_TC=0 keeps the mma.sync kernel)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eilev_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


def case(name, b, heads, d, sq, skv, causal, drop=False, mask=False):
    hd = heads * d
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(b, sq, hd, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(b, skv, hd, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(b, skv, hd, device="cuda", generator=g).to(torch.bfloat16)
    d_o = torch.randn(b, sq, hd, device="cuda", generator=g).to(torch.bfloat16)
    scale = d ** -0.5
    dr = (0.1, torch.tensor([7], dtype=torch.int64, device="cuda"), 3) if drop else None
    km = None
    if mask:  # a key-padding mask as the collator makes it: a few padded positions, the rest ones
        km = torch.ones(b, skv, dtype=torch.uint8, device="cuda")
        km[:, :6] = 0
    o, lse = ops.attention(q, k, v, heads, scale, causal=causal, need_lse=True, dropout=dr, key_mask=km)
    t_f = timeit(lambda: ops.attention(q, k, v, heads, scale, causal=causal, need_lse=True, dropout=dr, key_mask=km))
    t_b = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale, causal=causal, dropout=dr, key_mask=km))
    tc = ops.attention_bwd_uses_tcgen05(q, k, v, o, lse, d_o, heads, scale, causal=causal, dropout=dr, key_mask=km)
    kind = ops.attention_kernel(q, k, v, heads, causal=causal, need_lse=True)
    print(f"{name}: fwd {t_f * 1e3:.1f} us ({kind}), bwd (delta + kernels) {t_b * 1e3:.1f} us, tcgen05 bwd = {tc}", flush=True)


# the first measurements of a process run at ramping clocks: spend them on a throw-away case
case("(clock warm-up, ignore)", 1, 32, 80, 976, 976, True)
case("(clock warm-up, ignore)", 1, 32, 80, 976, 976, True)
print("VB_ATTN_BWD_TC =", os.environ.get("VB_ATTN_BWD_TC", "(default on)"), " VB_ATTN_FWD_TC =", os.environ.get("VB_ATTN_FWD_TC", "(default on)"))
case("opt self-attention 976 x 976 causal", 1, 32, 80, 976, 976, True)
case("q-former cross-attention 32 x 2056", 17, 12, 64, 32, 2056, False)
case("q-former self-attention 32 x 32", 17, 12, 64, 32, 32, False)
case("t5 encoder 976 x 976", 1, 32, 64, 976, 976, False)
case("opt self-attention 976 x 976 causal, key-padding mask", 1, 32, 80, 976, 976, True, mask=True)
case("q-former cross-attention 32 x 2056, dropout 0.1", 17, 12, 64, 32, 2056, False, drop=True)
case("q-former self-attention 32 x 32, dropout 0.1", 17, 12, 64, 32, 32, False, drop=True)
case("t5 encoder 976 x 976, dropout 0.1", 1, 32, 64, 976, 976, False, drop=True)

