"""Phase timeline of the persistent decode-step kernel (vb_debug_decode_trace): per op type,
mean over CTAs of the time spent staging x, consuming weights, priming, finalising and
waiting at the grid barrier, plus the spread of barrier arrival times."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from eilev_b200 import _lib  # noqa: E402
from eilev_b200.engine import opt as E_opt  # noqa: E402

cfg = bench.full_config()
dev = torch.device("cuda", 0)
model = bench.build_gpu_model(cfg, dev).eval()
one = bench.synthetic_batch(7)
n = int(one["attention_mask"].sum()) - bench.TARGET_TOKENS
ids = one["input_ids"][:, :n].to(dev); vm = one["video_input_mask"][:, :n].to(dev)
with torch.no_grad():
    feats, _, _ = model._video_features(one["pixel_values"].to(dev), False, train=False)
    lm = model.language_model
    logits, state = E_opt.opt_prefill(lm, lm._pack, ids, torch.ones_like(ids), vm.bool(), feats, 16)
    tok = logits.argmax(-1)
    for _ in range(3):
        tok = E_opt.opt_decode_step(lm, lm._pack, tok, state).argmax(-1)
    prog = state["program"]
    n_ops = prog.host.shape[0]
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    trace = torch.zeros(n_ops * sms * 6, dtype=torch.int64, device=dev)
    _lib.lib().vb_debug_decode_trace(trace.data_ptr())
    E_opt.opt_decode_step(lm, lm._pack, tok, state)
    torch.cuda.synchronize()
    _lib.lib().vb_debug_decode_trace(None)
t = trace.cpu().numpy().reshape(n_ops, sms, 6).astype(np.float64) / 1e3  # us
kinds = prog.host["type"]
names = {1: "gemv", 2: "attn", 3: "embed"}
total = (t[-1, :, 4].max() - t[0, :, 0].min())
print(f"step total {total:.1f} us over {n_ops} ops")
labels = ["stage_x", "main", "prime+sync", "finalize", "barrier"]
# per-op-shape breakdown for layer 5 and the head
def describe(i):
    d = np.diff(t[i], axis=1)  # (sms, 5)
    arr = t[i, :, 4]
    kind = names[int(kinds[i])]
    extra = ""
    if kinds[i] == 1:
        extra = f" n={int(prog.host['i64'][i][0])} k={int(prog.host['i64'][i][1])}"
    print(f"op {i:3d} {kind}{extra}: " + "  ".join(f"{l} {d[:, j].mean():6.2f}" for j, l in enumerate(labels))
          + f" | op wall {t[i, :, 5].max() - t[i, :, 0].min():6.2f} | arrival spread {arr.max() - arr.min():5.2f}")
for i in list(range(0, 7)) + list(range(26, 31)) + [n_ops - 2, n_ops - 1]:
    describe(i)
d = np.diff(t, axis=2)  # (ops, sms, 5)
for k in (1, 2, 3):
    sel = kinds == k
    if sel.any():
        print(names[k], "sum over ops of mean-over-CTAs:", "  ".join(f"{l} {d[sel][:, :, j].mean(axis=1).sum():8.1f}" for j, l in enumerate(labels)))
