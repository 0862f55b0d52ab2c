"""Why is the masked forward faster?  OPT shape, forward only: no mask / all-ones mask / 6 padded keys, twice each."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from eilev_b200 import ops  # noqa: E402

def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts) * 1e3

b, heads, d, L = 1, 32, 80, 976
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(b, L, 3 * heads * d, device="cuda", generator=g).to(torch.bfloat16)
hd = heads * d
q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
ones = torch.ones(b, L, dtype=torch.uint8, device="cuda")
pad = ones.clone(); pad[:, :6] = 0
for rep in range(2):
    for name, km in (("no mask", None), ("all-ones mask", ones), ("6 padded keys", pad)):
        for lse in (True, False):
            t = timeit(lambda: ops.attention(q, k, v, heads, d ** -0.5, causal=True, key_mask=km, need_lse=lse))
            print(f"{name:15s} need_lse={lse}: {t:.1f} us", flush=True)
