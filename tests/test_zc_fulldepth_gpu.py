"""Full-depth parity at the benchmarked architecture.

The full 39-layer ViT-g / 12-layer Q-Former / 32-layer OPT-2.7B model (and the 24+24-layer
flan-t5-xl branch) with the sane seeded init (seed 1234, N(0, 0.02), LayerNorm 1 / 0), 2 clips x 8
frames, L = 115-120 — BASELINE.md §2.1's yardstick case — CUDA path vs the fp32 CPU oracle
(oracle/videoblip_ref.py, pinned to the real reference by tests/test_oracle.py) for logits, loss
and the gradients of the 257 trainable tensors.

Stated tolerance: logits max-abs <= 0.15 at logit std ~1 and
rel-L2 <= 2.5 %, loss |d| <= 0.05, global gradient rel-L2 <= 10 %.  Those figures were set from a
2-layer yardstick; tests/golden/bf16_yardstick_fulldepth.py measured the REAL reference class at this
depth on these very inputs, bf16 vs its own fp32 (profiles/r02_bf16_yardstick_fulldepth.json):

    opt-2.7b     bf16-resident: logits 1.57 % / 0.092, grads 12.3 %   (autocast: 0.88 % / 0.047, 9.9 %)
    flan-t5-xl   bf16-resident: logits 2.87 % / 0.123, grads  5.0 %   (autocast: 2.45 % / 0.109, 4.5 %)

i.e. the reference's own bf16-resident run already exceeds two of the stated bounds (OPT gradients, T5
logits rel-L2).  Each metric is therefore held to max(stated bound, 1.1 x the reference's own
bf16-resident gap) — the CUDA path keeps bf16 weights and a bf16 residual stream like that mode — and
both numbers are written to the report.

Reference: /root/reference/eilev/model/v2.py:132-252.  The numbers are written to
gpurun_out/parity_report_fulldepth.json (committed as profiles/r02_parity_fulldepth.json).
"""
import json
import os
from pathlib import Path

import pytest
import torch
from transformers import Blip2Config

pytestmark = pytest.mark.gpu

REPORT = {}

VISION = dict(hidden_size=1408, intermediate_size=6144, num_hidden_layers=39, num_attention_heads=16,
              patch_size=14, image_size=224, hidden_act="gelu", layer_norm_eps=1e-6, qkv_bias=True)
QFORMER = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
               encoder_hidden_size=1408, cross_attention_frequency=2, vocab_size=30522,
               hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
OPT = dict(model_type="opt", hidden_size=2560, num_hidden_layers=32, ffn_dim=10240, num_attention_heads=32,
           vocab_size=50272, max_position_embeddings=2048, word_embed_proj_dim=2560, dropout=0.0,
           attention_dropout=0.0)
T5 = dict(model_type="t5", d_model=2048, d_kv=64, d_ff=5120, num_layers=24, num_decoder_layers=24, num_heads=32,
          vocab_size=32128, feed_forward_proj="gated-gelu", tie_word_embeddings=False, decoder_start_token_id=0,
          pad_token_id=0, eos_token_id=1, dropout_rate=0.0, relative_attention_num_buckets=32,
          relative_attention_max_distance=128)


YARDSTICK = Path(__file__).resolve().parent.parent / "profiles" / "r02_bf16_yardstick_fulldepth.json"
STATED = dict(logits_rel_l2=0.025, logits_max_abs=0.15, grad_rel_l2=0.10)


def _bounds(lm: str) -> dict:
    """max(stated tolerance, 1.1 x the real reference's own bf16-resident-vs-fp32 gap at full depth)."""
    ref = json.loads(YARDSTICK.read_text())[lm]["bf16_params"] if YARDSTICK.exists() else {}
    return {k: max(v, 1.1 * float(ref.get(k, 0.0))) for k, v in STATED.items()}


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _dump(key, **vals):
    REPORT[key] = vals
    out = Path(os.environ.get("GRAFT_REPO_ROOT", ".")) / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_report_fulldepth.json").write_text(json.dumps(REPORT, indent=1))


def _seeded_state_dict(model, seed=1234):
    """The sane seeded init of SURVEY §0.8 / bench.py, generated tensor by tensor (no second copy)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        lk = k.lower()
        if "layernorm" in lk or "layer_norm" in lk or lk.endswith("final_layer_norm.weight"):
            sd[k] = torch.ones(v.shape) if k.endswith("weight") else torch.zeros(v.shape)
        else:
            sd[k] = torch.empty(v.shape).normal_(0.0, 0.02, generator=g)
    return sd


def _opt_inputs(nv=2, t=8, nq=32, text=24, target=12):
    g = torch.Generator().manual_seed(3)
    px = torch.randn(nv, 3, t, 224, 224, generator=g)
    ids, vm = [2], [0]
    for _ in range(nv):
        ids += [1] * nq + [50118] + torch.randint(4, 50000, (text,), generator=g).tolist()
        vm += [1] * nq + [0] * (1 + text)
    lab = [-100] * (len(ids) - target) + ids[-target:]
    pad = (-len(ids)) % 8
    attn = [1] * len(ids) + [0] * pad
    ids += [1] * pad; vm += [0] * pad; lab += [-100] * pad
    return dict(input_ids=torch.tensor([ids]), attention_mask=torch.tensor([attn]), pixel_values=px,
                video_input_mask=torch.tensor([vm]), labels=torch.tensor([lab]))


def _t5_inputs(nv=2, t=8, nq=32):
    g = torch.Generator().manual_seed(4)
    px = torch.randn(nv, 3, t, 224, 224, generator=g)
    ids, vm = [], []
    for _ in range(nv):
        ids += [0] * nq + [3] + torch.randint(4, 32000, (24,), generator=g).tolist()
        vm += [1] * nq + [0] * 25
    ids += [1]; vm += [0]
    pad = (-len(ids)) % 8
    attn = [1] * len(ids) + [0] * pad
    ids += [0] * pad; vm += [0] * pad
    labels = torch.randint(4, 32000, (1, 12), generator=g)
    return dict(input_ids=torch.tensor([ids]), attention_mask=torch.tensor([attn]), pixel_values=px,
                video_input_mask=torch.tensor([vm]), labels=labels)


def _grad_gap(model, sd):
    num = den = 0.0
    worst = ("", 0.0)
    n = 0
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        rg = sd[name].grad
        assert rg is not None, name
        d = float((p.grad.float().cpu() - rg).pow(2).sum())
        e = float(rg.pow(2).sum())
        num += d
        den += e
        n += 1
        if e > 0 and (d / e) ** 0.5 > worst[1] and rg.numel() >= 768:
            worst = (name, (d / e) ** 0.5)
    return (num / den) ** 0.5, n, worst


def test_full_depth_opt_forward_backward_against_oracle():
    from oracle import videoblip_ref as R
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import freeze_for_recipe

    cfg = Blip2Config(vision_config=VISION, qformer_config=QFORMER, text_config=OPT, num_query_tokens=32)
    with torch.device("meta"):
        skeleton = VideoBlipForConditionalGeneration(cfg)
    sd = _seeded_state_dict(skeleton)
    sd["language_model.lm_head.weight"] = sd["language_model.model.decoder.embed_tokens.weight"]  # tied
    inputs = _opt_inputs()
    assert inputs["input_ids"].shape[1] == 120

    with torch.device("cuda"):  # built on the device: no 15 GB host-side random init
        m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(sd)
    m = m.to(torch.bfloat16).train()
    freeze_for_recipe(m)
    for p in m.parameters():
        if p.requires_grad:
            p.data = p.data.float()
    out = m(**{k: v.cuda() for k, v in inputs.items()}, return_dict=True)
    out.loss.backward()
    torch.cuda.synchronize()

    trainable = [k for k in sd if k.startswith(("qformer.", "query_tokens", "language_projection."))]
    for k in trainable:
        sd[k].requires_grad_(True)
    ref = R.videoblip_forward(sd, cfg, **inputs)
    ref["loss"].backward()

    valid = inputs["attention_mask"][0].bool()
    lg, rl = out.logits[0, valid].float().cpu(), ref["logits"][0, valid].detach()
    grad, n_grads, worst = _grad_gap(m, sd)
    r = dict(
        image_embeds=rel_l2(out.vision_outputs.last_hidden_state, ref["image_embeds"].detach()),
        query_output=rel_l2(out.qformer_outputs.last_hidden_state, ref["query_output"].detach()),
        logits_rel_l2=rel_l2(lg, rl), logits_max_abs=float((lg - rl).abs().max()), logits_std=float(rl.std()),
        argmax_agreement=float((lg.argmax(-1) == rl.argmax(-1)).float().mean()),
        loss=float(out.loss.detach()), loss_ref=float(ref["loss"]),
        grad_rel_l2=grad, grads_compared=n_grads, worst_tensor=worst[0], worst_tensor_rel_l2=worst[1],
        layers="39 ViT / 12 Q-Former / 32 OPT", clips=2, frames=8, seq_len=120,
    )
    bounds = _bounds("opt")
    r["bounds"] = bounds
    r["stated"] = STATED
    _dump("full_depth_opt", **r)
    assert n_grads == 257, n_grads
    assert r["logits_max_abs"] <= bounds["logits_max_abs"], r
    assert r["logits_rel_l2"] <= bounds["logits_rel_l2"], r
    assert abs(r["loss"] - r["loss_ref"]) <= 0.05, r
    assert r["grad_rel_l2"] <= bounds["grad_rel_l2"], r


def test_full_depth_t5_forward_backward_against_oracle():
    from oracle import videoblip_ref as R
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import freeze_for_recipe

    cfg = Blip2Config(vision_config=VISION, qformer_config=QFORMER, text_config=T5, num_query_tokens=32)
    with torch.device("meta"):
        skeleton = VideoBlipForConditionalGeneration(cfg)
    sd = _seeded_state_dict(skeleton)
    for k in sd:  # T5 attention is unscaled: keep the logits O(1) as the trained checkpoint does
        if k.startswith("language_model.") and k.endswith((".q.weight", ".k.weight")):
            sd[k] = sd[k] * 0.25
    sd["language_model.encoder.embed_tokens.weight"] = sd["language_model.shared.weight"]
    sd["language_model.decoder.embed_tokens.weight"] = sd["language_model.shared.weight"]
    inputs = _t5_inputs()
    ids = inputs["input_ids"][0]

    with torch.device("cuda"):  # built on the device: no 15 GB host-side random init
        m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(sd)
    m = m.to(torch.bfloat16).train()
    freeze_for_recipe(m)
    for p in m.parameters():
        if p.requires_grad:
            p.data = p.data.float()
    out = m(**{k: v.cuda() for k, v in inputs.items()}, return_dict=True)
    out.loss.backward()
    torch.cuda.synchronize()

    trainable = [k for k in sd if k.startswith(("qformer.", "query_tokens", "language_projection."))]
    for k in trainable:
        sd[k].requires_grad_(True)
    ref = R.videoblip_forward_t5(sd, cfg, **inputs)
    ref["loss"].backward()
    lg, rl = out.logits.float().cpu(), ref["logits"].detach()
    grad, n_grads, worst = _grad_gap(m, sd)
    r = dict(logits_rel_l2=rel_l2(lg, rl), logits_max_abs=float((lg - rl).abs().max()), logits_std=float(rl.std()),
             loss=float(out.loss.detach()), loss_ref=float(ref["loss"]), grad_rel_l2=grad, grads_compared=n_grads,
             worst_tensor=worst[0], worst_tensor_rel_l2=worst[1],
             layers="39 ViT / 12 Q-Former / 24+24 flan-t5-xl", clips=2, frames=8, seq_len=int(len(ids)))
    bounds = _bounds("t5")
    r["bounds"] = bounds
    r["stated"] = STATED
    _dump("full_depth_t5", **r)
    assert n_grads == 257, n_grads
    assert r["logits_max_abs"] <= bounds["logits_max_abs"], r
    assert r["logits_rel_l2"] <= bounds["logits_rel_l2"], r
    assert abs(r["loss"] - r["loss_ref"]) <= 0.05, r
    assert r["grad_rel_l2"] <= bounds["grad_rel_l2"], r
