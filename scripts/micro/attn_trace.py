"""Timeline of CTA 0 of the flash attention kernels on the OPT shape (build with scripts/micro/attn_trace.sh).
Per kernel (forward, dK/dV, dQ): the MMA issuer's and the two warp groups' events, in clocks since the first."""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
os.environ.setdefault("VB_LIB_PATH", str(ROOT / "build" / "libvideoblip_b200_fatrace.so"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from eilev_b200 import _lib, ops  # noqa: E402

lib = _lib.lib()
lib.vb_debug_attn_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
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
NAMES = {0: {1: "wait_u", 2: "got_u", 3: "acc_issued", 4: "wait_c", 5: "got_c", 6: "scores_issued"},
         1: {1: "item_start", 2: "wait_t", 3: "got_t", 4: "arrived"}}


def dump(title):
    buf = (C.c_longlong * (3 * 512))()
    n = (C.c_int * 3)()
    assert lib.vb_debug_attn_trace(buf, n, 1) == 0
    ev = []
    for who in range(3):
        for i in range(n[who]):
            ev.append((buf[who * 512 + 2 * i + 1], who, buf[who * 512 + 2 * i]))
    if not ev:
        return
    ev.sort()
    t0 = ev[0][0]
    print(f"== {title}: {len(ev)} events, {ev[-1][0] - t0} clk from first to last")
    for t, who, tag in ev[:150]:
        kind, item = tag // 100, tag % 100
        nm = NAMES[0 if who == 0 else 1][kind]
        print(f"  {t - t0:7d}  {'mma ' if who == 0 else 'grp' + str(who - 1)}  {nm:14s} item {item}")


lib.vb_debug_attn_trace(None, None, 1)
o, lse = ops.attention(q, k, v, heads, d ** -0.5, causal=causal, need_lse=True)
torch.cuda.synchronize()
dump("forward")
ops.attention_bwd(q, k, v, o, lse, d_o, heads, d ** -0.5, causal=causal)
torch.cuda.synchronize()
dump("backward (dK/dV then dQ share the buffer: the dQ pass appends)")
