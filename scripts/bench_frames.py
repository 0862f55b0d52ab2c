"""uint8 frame path micro-benchmark: one datapoint = 17 clips x 8 frames.
Times, with CUDA events after warm-up, (1) the PIL-exact bicubic resize 448 x 448 -> 224 x 224
(two vb_resize_u8_pass launches), (2) vb_patch_gather_u8 (rescale + normalize + im2col from bytes)
against vb_patch_gather from fp32 frames, and reports achieved HBM GB/s against the algorithmic bytes
of each, plus the host->device copy of the same datapoint as fp32 (the reference's contract) and as
uint8.  Prints one JSON line.  Usage: python scripts/bench_frames.py [in_size]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eilev_b200 import ops  # noqa: E402

CLIPS, T, C, OUT, PATCH, KPAD = 17, 8, 3, 224, 14, 608
in_size = int(sys.argv[1]) if len(sys.argv) > 1 else 448
MEAN, STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


g = torch.Generator().manual_seed(0)
raw_host = torch.randint(0, 256, (CLIPS, C, T, in_size, in_size), dtype=torch.uint8, generator=g).pin_memory()
raw = raw_host.cuda()
small = ops.resize_bicubic_u8(raw, OUT, OUT)
small_host = small.cpu().pin_memory()
f32 = ((small.float() / 255) - torch.tensor(MEAN, device="cuda").view(1, 3, 1, 1, 1)) / torch.tensor(STD, device="cuda").view(1, 3, 1, 1, 1)
f32_host = f32.cpu().pin_memory()
planes = CLIPS * C * T
plan = ops.resize_plan(in_size, in_size, OUT, OUT)
rows = plan[0]["lines"] if plan else 0
resize_bytes = planes * (rows * in_size + 2 * rows * OUT + OUT * OUT)  # read src rows, write + read tmp, write out
patches = CLIPS * T * (OUT // PATCH) ** 2
gather_u8_bytes = planes * OUT * OUT + patches * KPAD * 2
gather_f32_bytes = planes * OUT * OUT * 4 + patches * KPAD * 2
ms_resize = timed(lambda: ops.resize_bicubic_u8(raw, OUT, OUT)) if plan else 0.0
ms_u8 = timed(lambda: ops.patch_gather_u8(small, PATCH, KPAD, 1 / 255, MEAN, STD))
ms_f32 = timed(lambda: ops.patch_gather(f32, PATCH, KPAD))
dst_u8, dst_f32 = torch.empty_like(small), torch.empty_like(f32)
ms_h2d_u8 = timed(lambda: dst_u8.copy_(small_host, non_blocking=True))
ms_h2d_f32 = timed(lambda: dst_f32.copy_(f32_host, non_blocking=True))
print(json.dumps({
    "workload": f"{CLIPS} clips x {T} frames, {in_size}^2 -> {OUT}^2, patch {PATCH}",
    "resize_u8": {"ms": ms_resize, "algorithmic_bytes": resize_bytes, "GBps": resize_bytes / ms_resize / 1e6 if ms_resize else None},
    "patch_gather_u8": {"ms": ms_u8, "algorithmic_bytes": gather_u8_bytes, "GBps": gather_u8_bytes / ms_u8 / 1e6},
    "patch_gather_f32": {"ms": ms_f32, "algorithmic_bytes": gather_f32_bytes, "GBps": gather_f32_bytes / ms_f32 / 1e6},
    "h2d_u8": {"ms": ms_h2d_u8, "bytes": small_host.numel()}, "h2d_f32": {"ms": ms_h2d_f32, "bytes": f32_host.numel() * 4},
}))
