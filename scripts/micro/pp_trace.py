"""Timeline of CTA 0 of the ViT attention kernel (attn_tcgen05_pp_kernel) on the ViT-g shape
(build with scripts/micro/pp_trace.sh): the MMA issuer's and the two softmax groups' events in clocks."""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
os.environ.setdefault("VB_LIB_PATH", str(ROOT / "build" / "libvideoblip_b200_pptrace.so"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from eilev_b200 import _lib, ops  # noqa: E402

lib = _lib.lib()
lib.vb_debug_pp_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
frames, heads, d, s = 136, 16, 88, 257
hd = heads * d
qkv = torch.randn(frames, s, 3 * hd, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
for _ in range(3):
    ops.attention(q, k, v, heads, d ** -0.5)
torch.cuda.synchronize()
NAMES = {0: {1: "qk_wait", 2: "qk_go", 3: "pv_wait", 4: "pv_go"},
         1: {1: "tile_start", 2: "got_q", 3: "wait_s", 4: "got_s", 5: "max_done", 6: "p_done", 7: "got_o", 8: "epi_done"}}
lib.vb_debug_pp_trace(None, None, 1)
ops.attention(q, k, v, heads, d ** -0.5)
torch.cuda.synchronize()
buf = (C.c_longlong * (3 * 1024))()
n = (C.c_int * 3)()
assert lib.vb_debug_pp_trace(buf, n, 1) == 0
ev = []
for who in range(3):
    for i in range(n[who]):
        ev.append((buf[who * 1024 + 2 * i + 1], who, buf[who * 1024 + 2 * i]))
ev.sort()
t0 = ev[0][0]
print(f"{len(ev)} events, {ev[-1][0] - t0} clk from first to last")
last = {0: t0, 1: t0, 2: t0}
for t, who, tag in ev[:260]:
    nm = NAMES[0 if who == 0 else 1][tag]
    col = {0: 0, 1: 30, 2: 60}[who]
    print(f"  {t - t0:7d}  " + " " * col + f"{'mma ' if who == 0 else 'grp' + str(who - 1)} {nm:10s} +{t - last[who]}")
    last[who] = t
