"""GEMV (decode projection) micro-benchmark: achieved HBM GB/s per shape at M=1, both as
isolated launches and as a 32-layer chain of distinct weights (no L2 reuse)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eilev_b200 import ops  # noqa: E402

def ev():
    return torch.cuda.Event(enable_timing=True)

m = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1
layers = 32
shapes = [("qkv", 7680, 2560), ("out", 2560, 2560), ("fc1", 10240, 2560), ("fc2", 2560, 10240)]
head = (torch.randn(50272, 2560, device="cuda") * 0.02).to(torch.bfloat16)
W = {n: [(torch.randn(nn, k, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(layers)] for n, nn, k in shapes}
g = torch.ones(2560, device="cuda"); b = torch.zeros(2560, device="cuda")
for name, nn, k in shapes:
    x = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    for ln in (None, (g, b, 1e-5)) if k == 2560 else (None,):
        for _ in range(2):
            for w in W[name]: ops.gemv(x, w, ln=ln)
        s, e = ev(), ev(); s.record()
        for w in W[name]: ops.gemv(x, w, ln=ln)
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) * 1e3 / layers
        print(f"{name:4s} N={nn:5d} K={k:5d} ln={ln is not None!s:5s}: {us:7.1f} us/launch  {nn * k * 2 / us / 1e3:7.1f} GB/s", flush=True)
xh = torch.randn(m, 2560, device="cuda").to(torch.bfloat16)
for _ in range(3): ops.gemv(xh, head, out_dtype=torch.float32)
s, e = ev(), ev(); s.record()
for _ in range(5): ops.gemv(xh, head, out_dtype=torch.float32, ln=(g, b, 1e-5))
e.record(); torch.cuda.synchronize()
us = s.elapsed_time(e) * 1e3 / 5
print(f"head N=50272 K= 2560 (L2-warm x5): {us:7.1f} us/launch  {50272 * 2560 * 2 / us / 1e3:7.1f} GB/s", flush=True)
# graph of the whole chain
x = torch.randn(m, 2560, device="cuda").to(torch.bfloat16)
PREFETCH = "--prefetch" in sys.argv
def chain():
    h = x
    for i in range(layers):
        pf = (lambda t: t) if PREFETCH else (lambda t: None)
        q = ops.gemv(h, W["qkv"][i], ln=(g, b, 1e-5))
        h = ops.gemv(q[:, :2560].contiguous(), W["out"][i], residual=h)
        f = ops.gemv(h, W["fc1"][i], epilogue=ops.EPI_RELU, ln=(g, b, 1e-5))
        h = ops.gemv(f, W["fc2"][i], residual=h)
    return h
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side): chain()
torch.cuda.current_stream().wait_stream(side)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr): out = chain()
for _ in range(3): gr.replay()
s, e = ev(), ev(); s.record()
for _ in range(10): gr.replay()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
tot = sum(nn * k * 2 for _, nn, k in shapes) * layers
print(f"graph chain (128 GEMVs + 32 small copies): {ms:.3f} ms  {tot / ms / 1e6:.1f} GB/s")
