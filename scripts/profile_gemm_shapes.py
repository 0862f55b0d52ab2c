"""Launches the five large tcgen05 GEMM shapes of the ViT / cross-K|V path once each inside a
cudaProfilerStart/Stop bracket, in the order of SHAPES, for

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:gemm_tcgen05 -o gpurun_out/r02_gemm_shapes python scripts/profile_gemm_shapes.py

scripts/gemm_traffic.py turns the report into profiles/r02_ncu_gemm_traffic.json (DRAM bytes per launch
per shape), which bench.py reads for roofline.traffic."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eilev_b200 import ops  # noqa: E402
from eilev_b200.engine.packing import ln_fold  # noqa: E402

# (name, M, N, K, epilogue, residual in place, LayerNorm fold, output statistics)
SHAPES = [
    ("vit.qkv", 34952, 4224, 1408, ops.EPI_NONE, False, True, False),
    ("vit.proj", 34952, 1408, 1408, ops.EPI_NONE, True, False, True),
    ("vit.fc1", 34952, 6144, 1408, ops.EPI_GELU, False, True, False),
    ("vit.fc2", 34952, 1408, 6144, ops.EPI_NONE, True, False, True),
    ("qf.crosskv", 34952, 9216, 1408, ops.EPI_NONE, False, False, False),
]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    cases = []
    for name, m, n, k, epi, res, fold, stats in SHAPES:
        a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda", generator=g)
        out = torch.randn(m, n, device="cuda", generator=g).to(torch.bfloat16)
        kw = dict(epilogue=epi, out=out)
        if res:
            kw["residual"] = out
        if fold:
            w, bias, cs = ln_fold(w, bias, torch.ones(k, device="cuda"), torch.zeros(k, device="cuda"))
            kw["ln_fold"] = (ops.row_stats(a), cs, 1e-6)
        if stats:
            kw["stats_out"] = torch.zeros(m, 2, device="cuda", dtype=torch.float64)
            kw["stats_zero"] = torch.zeros(m, 2, device="cuda", dtype=torch.float64)
        cases.append((a, w, bias, kw))
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for a, w, bias, kw in cases:  # warm-up (kernel attributes, allocator)
        ops.gemm(a, w, bias, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for a, w, bias, kw in cases:
        flush.zero_()  # cold L2, as inside the step
        ops.gemm(a, w, bias, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
