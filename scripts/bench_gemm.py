"""Micro-benchmark of the tcgen05 GEMM on the VideoBLIP shapes (CUDA events, L2 flushed
between iterations); prints TFLOP/s next to torch.matmul (cuBLAS) on the same shapes."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eilev_b200 import ops  # noqa: E402

SHAPES = [
    ("vit.qkv", 34952, 4224, 1408, ops.EPI_NONE),
    ("vit.proj", 34952, 1408, 1408, ops.EPI_NONE),
    ("vit.fc1", 34952, 6144, 1408, ops.EPI_GELU),
    ("vit.fc2", 34952, 1408, 6144, ops.EPI_NONE),
    ("qf.crosskv", 34952, 9216, 1408, ops.EPI_NONE),
    ("opt.qkv", 976, 7680, 2560, ops.EPI_NONE),
    ("opt.fc1", 976, 10240, 2560, ops.EPI_RELU),
    ("opt.fc2", 976, 2560, 10240, ops.EPI_NONE),
    ("opt.head", 976, 50272, 2560, ops.EPI_NONE),
    ("opt.proj", 976, 2560, 2560, ops.EPI_NONE),
    ("opt.dqkv", 976, 2560, 7680, ops.EPI_NONE),
    ("qf.dense", 544, 768, 768, ops.EPI_NONE),
]


def timeit(fn, iters=5):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), sum(ts) / len(ts)


def sweep():
    """--sweep: every tile variant (1-CTA 64..256, CTA-pair 1000+BN) on the big ViT shapes."""
    opt = "--opt" in sys.argv
    which = [s for s in SHAPES if s[0].startswith("opt.")] if opt else SHAPES[:5]
    for name, m, n, k, epi in which:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda")
        out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        res = {}
        widths = (0, 176, 256, 1128, 1144, 1160, 1176, 1192, 1208, 1224, 1256) if opt else (64, 128, 176, 256, 1128, 1256)
        for bn in widths:
            best, _ = timeit(lambda: ops.gemm(a, w, bias, out=out, epilogue=epi, block_n=bn), iters=3)
            res[bn] = round(2.0 * m * n * k / best / 1e9)
        print(name, res, flush=True)


def main():
    if "--sweep" in sys.argv:
        return sweep()
    only = sys.argv[1:]
    rows = []
    for name, m, n, k, epi in SHAPES:
        if only and not any(o in name for o in only):
            continue
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda")
        out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        res = torch.randn(m, n, device="cuda").to(torch.bfloat16) if "fc2" in name or "proj" in name else None
        best, avg = timeit(lambda: ops.gemm(a, w, bias, residual=res, out=out, epilogue=epi))
        cb, ca = timeit(lambda: torch.matmul(a, w.t(), out=out))
        fl = 2.0 * m * n * k
        row = dict(name=name, m=m, n=n, k=k, ours_ms=round(best, 4), ours_avg_ms=round(avg, 4),
                   ours_tflops=round(fl / best / 1e9, 1), cublas_ms=round(cb, 4),
                   cublas_tflops=round(fl / cb / 1e9, 1))
        rows.append(row)
        print(json.dumps(row), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/bench_gemm.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
