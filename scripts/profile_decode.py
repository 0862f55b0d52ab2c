"""Runs a short eager greedy decode under the CUDA profiler range (for ncu --profile-from-start off)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
cfg = bench.full_config()
dev = torch.device("cuda", 0)
model = bench.build_gpu_model(cfg, dev).eval()
one = bench.synthetic_batch(7)
n = int(one["attention_mask"].sum()) - bench.TARGET_TOKENS
ids = one["input_ids"][:, :n].to(dev); vm = one["video_input_mask"][:, :n].to(dev)
kw = dict(pixel_values=one["pixel_values"].to(dev), video_input_mask=vm, attention_mask=torch.ones_like(ids),
          do_sample=False, eos_token_id=None, use_cuda_graph=False)
model.generate(ids, max_new_tokens=2, min_new_tokens=2, **kw)
torch.cuda.synchronize()
from eilev_b200.engine import opt as E_opt
feats, _, _ = model._video_features(kw["pixel_values"], False, train=False)
lm = model.language_model
logits, state = E_opt.opt_prefill(lm, lm._pack, ids, kw["attention_mask"], vm.bool(), feats, 8)
tok = logits.argmax(-1)
E_opt.opt_decode_step(lm, lm._pack, tok, state)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    E_opt.opt_decode_step(lm, lm._pack, tok, state)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
