"""Parity of the CUDA v1 path (eilev_b200.model.v1: HF 4.33.1 Blip2 call signatures over the v2
kernels) with golden outputs of the REAL reference class eilev.model.v1 (tests/golden/v1_*.pt,
tests/golden/make_golden_v1.py).  Same tolerances as tests/test_model_gpu.py: bf16 kernels with
fp32 accumulation vs an fp32 reference — logits rel-L2 <= 3 %, loss |d| <= 0.03; gradients global
rel-L2 <= 9 % on small_opt, 6 % on small_t5 and 25 % on tiny_opt, whose v1 fixture (width 8, six
target tokens) is precision-noisy: the REAL reference's own bf16 run differs from its fp32 run by
14-19 % there, 5.6-6.5 % on small_opt and 1.3-1.8 % on small_t5 (tests/golden/bf16_yardstick_v1.py);
greedy token ids exact up to numerical ties (fp32 top-2 logit margin < 0.1, stored in the fixture;
the bf16 logit error is ~0.05).  First B200 run (profiles/r01_parity_report_v1.json): logits 2.4 % /
0.9 % / 1.0 %, gradients 20 % / 4.9 % / 1.4 % on tiny_opt / small_opt / small_t5."""
import json
import os
from pathlib import Path

import pytest
import torch
from transformers import Blip2Config

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
REPORT = {}


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _dump(key, **vals):
    REPORT[key] = vals
    out = Path(os.environ.get("GRAFT_REPO_ROOT", ".")) / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_report_v1.json").write_text(json.dumps(REPORT, indent=1))


def load(name):
    base = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = Blip2Config(**{k: base["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    return torch.load(GOLDEN / f"v1_{name}.pt", weights_only=False), base["state_dict"], cfg


def build(cfg, sd):
    from eilev_b200.model.v1 import VideoBlipForConditionalGeneration
    m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(sd)
    return m.to("cuda", torch.float32).eval()


def cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def assert_greedy_equal_up_to_ties(got, want, margins, key, tie=0.1):
    """Rows must reproduce the reference's greedy ids; a row may leave them only at a step whose
    fp32 top-2 margin in the reference (margins: (steps, rows)) is a numerical tie — after such a
    flip its continuation legitimately differs, so the row is not compared further."""
    got, want = got.cpu().tolist(), want.tolist()
    assert len(got) == len(want) and all(len(g) == len(w) for g, w in zip(got, want)), (got, want)
    flips = {}
    for row, (g, w) in enumerate(zip(got, want)):
        for t, (a, b) in enumerate(zip(g, w)):
            if a != b:
                m = float(margins[t - (len(w) - margins.shape[0]), row])
                assert m < tie, (key, row, t, g, w, m)
                flips[f"row{row}/step{t}"] = m
                break
    _dump(key, tie_flips=flips, rows=len(want))
    assert len(flips) <= max(1, len(want) // 2), flips


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt", "small_t5"])
def test_v1_forward_matches_reference_golden(name):
    fx, sd, cfg = load(name)
    m = build(cfg, sd)
    i = cuda(fx["inputs"])
    with torch.no_grad():
        out = m(**i, return_dict=True)
        extra = {} if cfg.use_decoder_only_language_model else \
            {"decoder_input_ids": torch.zeros(i["input_ids"].shape[0], 1, dtype=torch.long, device="cuda")}
        nolab = m(pixel_values=i["pixel_values"], input_ids=i["input_ids"], attention_mask=i["attention_mask"],
                  return_dict=True, **extra)
        tup = m(**i, return_dict=False)
    m.check_splice()
    assert out.logits.shape == fx["logits"].shape  # decoder-only: the last labels.size(1) positions
    assert nolab.logits.shape == fx["logits_no_labels"].shape and nolab.loss is None
    assert len(tup) == 5 and tup[1].shape == out.logits.shape
    r = dict(query_output=rel_l2(out.qformer_outputs.last_hidden_state, fx["query_output"]),
             logits=rel_l2(out.logits, fx["logits"]), logits_no_labels=rel_l2(nolab.logits, fx["logits_no_labels"]),
             loss=float(out.loss), loss_ref=float(fx["loss"]))
    _dump(f"v1_forward/{name}", **r)
    assert r["query_output"] < 0.03 and r["logits"] < 0.03 and r["logits_no_labels"] < 0.03, r
    assert abs(r["loss"] - r["loss_ref"]) < 0.03, r


@pytest.mark.parametrize("name,tol", [("tiny_opt", 0.25), ("small_opt", 0.09), ("small_t5", 0.06)])
def test_v1_backward_matches_reference_golden(name, tol):
    """train_v1.py recipe: ViT + LM frozen, Q-Former / query_tokens / language_projection train."""
    from eilev_b200.train import freeze_for_recipe
    fx, sd, cfg = load(name)
    m = build(cfg, sd).train()
    freeze_for_recipe(m)
    out = m(**cuda(fx["inputs"]), return_dict=True)
    out.loss.backward()
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(fx["grads"]), set(got) ^ set(fx["grads"])
    num = den = 0.0
    for n, ref in fx["grads"].items():
        num += float((got[n].float().cpu() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    glob = (num / den) ** 0.5
    _dump(f"v1_backward/{name}", global_rel_l2=glob, loss=float(out.loss.detach()), n=len(got))
    assert abs(float(out.loss) - float(fx["loss"])) < 0.03
    assert glob < tol, glob


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt", "small_t5"])
def test_v1_generate_matches_reference_golden(name):
    """Greedy ids on the left-padded prompt (video slots in front, padding moved to the left end
    for the paged decode kernels) and with no prompt at all ([bos] per row), token-exact on the
    fixtures up to numerical ties; the sampling kwargs of samples/video_blip_generate_action_narration.py:24-32 run."""
    fx, sd, cfg = load(name)
    m = build(cfg, sd)
    g = cuda(fx["gen_inputs"])
    gen = m.generate(**g, max_new_tokens=5, min_new_tokens=5, do_sample=False)
    assert_greedy_equal_up_to_ties(gen, fx["generated"], fx["generated_margins"], f"v1_generate/{name}")
    if cfg.use_decoder_only_language_model:
        gen = m.generate(pixel_values=g["pixel_values"], max_new_tokens=5, min_new_tokens=5, do_sample=False)
        assert_greedy_equal_up_to_ties(gen, fx["generated_no_prompt"], fx["generated_no_prompt_margins"],
                                       f"v1_generate_no_prompt/{name}")
    torch.manual_seed(0)
    s = m.generate(**g, num_beams=4, max_new_tokens=6, temperature=0.7, top_p=0.9, repetition_penalty=1.5,
                   do_sample=True)
    assert s.shape[0] == g["input_ids"].shape[0] and 1 <= s.shape[1] <= 7


def test_v1_vision_model_is_the_v2_class_and_rejects_missing_pixels():
    from eilev_b200.model import v1, v2
    assert v1.VideoBlipVisionModel is v2.VideoBlipVisionModel  # v1.py:14-92 == v2.py:20-103
    fx, sd, cfg = load("tiny_opt")
    m = build(cfg, sd)
    with pytest.raises(ValueError):
        m(pixel_values=None, input_ids=fx["inputs"]["input_ids"].cuda())
