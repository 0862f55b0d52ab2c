"""small_opt: gradients of one micro-step, (a) two fresh eager models against each other, (b) the trainer's graph
against eager; per-parameter relative error for the worst parameters."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from transformers import Blip2Config  # noqa: E402

from eilev_b200.model.v2 import VideoBlipForConditionalGeneration  # noqa: E402
from eilev_b200.train import DataParallelTrainer, freeze_for_recipe  # noqa: E402

fx = torch.load(ROOT / "tests" / "golden" / "small_opt.pt", weights_only=False)
cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
batch = {k: v.cuda() for k, v in fx["inputs"].items()}


def fresh():
    m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(fx["state_dict"])
    m = m.to("cuda").train()
    freeze_for_recipe(m)
    return m


def eager_grads():
    m = fresh()
    o = m(**batch, return_dict=True)
    (o.loss / 2).backward()
    return {n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad}, float(o.loss)


def report(tag, a, b):
    num = sum(float((a[n] - b[n]).pow(2).sum()) for n in b)
    den = sum(float(b[n].pow(2).sum()) for n in b)
    worst = sorted(((float((a[n] - b[n]).norm() / b[n].norm().clamp_min(1e-20)), n) for n in b), reverse=True)[:6]
    print(f"{tag}: global rel L2 {(num / den) ** 0.5:.4g}; worst: " + "; ".join(f"{n} {e:.3g}" for e, n in worst), flush=True)


g1, l1 = eager_grads()
g2, l2 = eager_grads()
print("eager losses", l1, l2)
report("eager vs eager", g2, g1)
tm = fresh()
tr = DataParallelTrainer(tm, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0, grad_accum=2)
tr.capture_graph(batch)
lt = float(tr.micro_step(batch))
gt = {n: p.grad.clone() for n, p in tm.named_parameters() if p.requires_grad}
print("trainer loss", lt)
report("trainer graph vs eager", gt, g1)
g3, l3 = eager_grads()
report("eager (after the trainer) vs eager", g3, g1)
