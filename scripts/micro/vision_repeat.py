"""Vision tower forward of the small_opt fixture, repeated, with every ops.gemm / ops.attention / ops.row_stats
result (and the statistics buffers the GEMM epilogues accumulate into) recorded in execution order: prints, per
call, the first recorded tensor that differs from call 0 — i.e. the kernel that first produced different bits."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from transformers import Blip2Config  # noqa: E402

from eilev_b200 import ops  # noqa: E402
from eilev_b200.model.v2 import VideoBlipVisionModel  # noqa: E402

fx = torch.load(ROOT / "tests" / "golden" / "small_opt.pt", weights_only=False)
cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
vm = VideoBlipVisionModel(cfg.vision_config)
sd = {k[len("vision_model."):]: v for k, v in fx["state_dict"].items() if k.startswith("vision_model.")}
vm.load_state_dict(sd)
vm = vm.to("cuda").eval()
px = fx["inputs"]["pixel_values"].cuda()

log = []
_gemm, _attn, _rs = ops.gemm, ops.attention, ops.row_stats


def gemm(*a, **kw):
    out = _gemm(*a, **kw)
    log.append(("gemm.out", out.clone()))
    if kw.get("stats_out") is not None:
        log.append(("gemm.stats_out", kw["stats_out"].clone()))
    return out


def attention(*a, **kw):
    out = _attn(*a, **kw)
    log.append(("attention.out", (out[0] if isinstance(out, tuple) else out).clone()))
    return out


def row_stats(*a, **kw):
    out = _rs(*a, **kw)
    log.append(("row_stats", out.clone()))
    return out


ops.gemm, ops.attention, ops.row_stats = gemm, attention, row_stats
runs = []
for i in range(10):
    log.clear()
    last = vm(pixel_values=px, return_dict=False)[0]
    torch.cuda.synchronize()
    runs.append(list(log) + [("last", last.clone())])
for i, r in enumerate(runs):
    first = next((j for j, (a, b) in enumerate(zip(r, runs[0])) if not torch.equal(a[1], b[1])), None)
    if first is None:
        print(f"call {i}: identical to call 0 ({len(r)} recorded tensors)")
    else:
        a, b = r[first][1].float(), runs[0][first][1].float()
        print(f"call {i}: first difference at record {first} = {r[first][0]}: {(a != b).sum().item()} elements, "
              f"max abs diff {(a - b).abs().max().item():.3g}, max rel {((a - b).abs() / b.abs().clamp_min(1e-20)).max().item():.3g}")
