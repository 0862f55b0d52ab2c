"""ViT attention micro-benchmark: tcgen05/TMEM kernel vs the mma.sync flash kernel."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eilev_b200 import ops  # noqa: E402


def timeit(fn, iters=5):
    for _ in range(2):
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


frames, heads, d, s = 136, 16, 88, 257
qkv = torch.randn(frames, s, 3 * heads * d, device="cuda").to(torch.bfloat16)
hd = heads * d
q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
fl = 4.0 * frames * heads * s * s * d
t1 = timeit(lambda: ops.attention(q, k, v, heads, d ** -0.5))
t2 = timeit(lambda: ops.attention(q, k, v, heads, d ** -0.5, need_lse=True))
print(f"tcgen05: {t1 * 1e3:.1f} us ({fl / t1 / 1e9:.0f} TFLOP/s)   mma.sync: {t2 * 1e3:.1f} us ({fl / t2 / 1e9:.0f} TFLOP/s)")
import torch.nn.functional as F
qh = q.reshape(frames, s, heads, d).transpose(1, 2)
kh = k.reshape(frames, s, heads, d).transpose(1, 2)
vh = v.reshape(frames, s, heads, d).transpose(1, 2)
t3 = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh))
print(f"torch SDPA (library, strided views): {t3 * 1e3:.1f} us ({fl / t3 / 1e9:.0f} TFLOP/s)")
