"""Repeats the flash attention forward / backward on fixed inputs and reports whether the outputs are bit-identical
from call to call (they must be: no atomics, fixed reduction order)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from eilev_b200 import ops  # noqa: E402

CASES = [(6, 4, 16, 17, 17, False), (2, 3, 16, 40, 40, True), (1, 32, 80, 976, 976, True), (17, 12, 64, 32, 2056, False),
         (2, 4, 64, 130, 515, False), (3, 2, 32, 64, 64, False), (3, 2, 16, 65, 65, False), (6, 4, 16, 17, 17, True)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for b, heads, d, sq, skv, causal in CASES:
    hd = heads * d
    g = torch.Generator(device="cuda").manual_seed(1)
    if sq == skv:
        qkv = torch.randn(b, sq, 3 * hd, device="cuda", generator=g).to(torch.bfloat16)
        q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
    else:
        q = torch.randn(b, sq, hd, device="cuda", generator=g).to(torch.bfloat16)
        k = torch.randn(b, skv, hd, device="cuda", generator=g).to(torch.bfloat16)
        v = torch.randn(b, skv, hd, device="cuda", generator=g).to(torch.bfloat16)
    d_o = torch.randn(b, sq, hd, device="cuda", generator=g).to(torch.bfloat16)
    o0, lse0 = ops.attention(q, k, v, heads, d ** -0.5, causal=causal, need_lse=True)
    g0 = ops.attention_bwd(q, k, v, o0, lse0, d_o, heads, d ** -0.5, causal=causal)
    bad_f = bad_l = bad_b = 0
    junk = []
    for i in range(reps):
        junk.append(torch.randn(1 + 37 * i, 1000, device="cuda"))  # move the allocator around
        o, lse = ops.attention(q, k, v, heads, d ** -0.5, causal=causal, need_lse=True)
        o_nl = ops.attention(q, k, v, heads, d ** -0.5, causal=causal)
        gr = ops.attention_bwd(q, k, v, o0, lse0, d_o, heads, d ** -0.5, causal=causal)
        bad_f += int(not torch.equal(o, o0)) + int(not torch.equal(o_nl, o0))
        bad_l += int(not torch.equal(lse, lse0))
        bad_b += sum(int(not torch.equal(a, c)) for a, c in zip(gr, g0))
    kind = ops.attention_kernel(q, k, v, heads, causal=causal, need_lse=True)
    md = (o.float() - o0.float()).abs().max().item()
    print(f"b={b} h={heads} d={d} sq={sq} skv={skv} causal={causal} [{kind}]: fwd mismatches {bad_f}, lse {bad_l}, "
          f"bwd {bad_b} of {reps} (last fwd max diff {md:.3g}); finite={bool(torch.isfinite(o.float()).all())}", flush=True)
