"""Phase timeline of the CTA-pair GEMM (build with scripts/micro/gemm_trace.sh; VB_LIB_PATH points the
binding at the trace build).  Prints, per shape, the median over CTAs of the clock64 deltas between the
kernel's phases: entry -> set-up done -> dependency wait done -> first four k-blocks landed -> last MMA
committed -> accumulator visible to the epilogue -> last slab handed to TMA -> stores drained -> exit."""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
os.environ.setdefault("VB_LIB_PATH", str(ROOT / "build" / "libvideoblip_b200_trace.so"))
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from eilev_b200 import _lib, ops  # noqa: E402

SHAPES = [("opt.proj", 976, 2560, 2560, 0), ("opt.proj.w256", 976, 2560, 2560, 1256), ("opt.fc2", 976, 2560, 10240, 0),
          ("opt.qkv", 976, 7680, 2560, 0), ("opt.fc1", 976, 10240, 2560, 0), ("vit.proj", 34952, 1408, 1408, 0)]
NAMES = ["entry", "setup", "dep_wait", "kb0", "kb1", "kb2", "kb3", "mma_done", "acc_seen", "slab_out", "drained",
         "cluster_sync", "dealloc"]


def main():
    lib = _lib.lib()
    lib.vb_debug_gemm_trace.argtypes = [C.c_void_p]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for name, m, n, k, bn in SHAPES:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        for it in range(3):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.gemm(a, w, out=out, block_n=bn)
            e.record()
            torch.cuda.synchronize()
        buf = (C.c_uint64 * (296 * 16))()
        assert lib.vb_debug_gemm_trace(buf) == 0
        t = torch.tensor(list(buf), dtype=torch.int64).view(296, 16)[:148].double()
        rel = t[:, :13] - t[:, :1]
        med = rel.median(dim=0).values
        mx = rel.max(dim=0).values
        print(f"{name}: {s.elapsed_time(e) * 1e3:.1f} us by events; clk since entry (median / max over CTAs):")
        print("   " + "  ".join(f"{nm}={int(a)}/{int(b)}" for nm, a, b in zip(NAMES, med.tolist(), mx.tolist())), flush=True)


if __name__ == "__main__":
    main()
