"""Pins the CPU oracle (oracle/videoblip_ref.py) against golden outputs of the real
reference (tests/golden/*.pt, produced by tests/golden/make_golden.py from
/root/reference/eilev/model/v2.py) — and against the live reference when it is present."""
import sys
from pathlib import Path

import pytest
import torch
from transformers import Blip2Config

from oracle import videoblip_ref as R

GOLDEN = Path(__file__).resolve().parent / "golden"
NAMES = ["tiny_opt", "small_opt"]


def load(name):
    fx = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    return fx, cfg


@pytest.mark.parametrize("name", NAMES)
def test_oracle_forward_matches_reference_golden(name):
    fx, cfg = load(name)
    out = R.videoblip_forward(fx["state_dict"], cfg, **fx["inputs"])
    assert torch.allclose(out["image_embeds"], fx["image_embeds"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out["pooler_output"], fx["pooler_output"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out["query_output"], fx["query_output"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out["logits"], fx["logits"], atol=5e-5, rtol=1e-4)
    assert abs(float(out["loss"]) - float(fx["loss"])) < 1e-5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_text_only_matches_reference_golden(name):
    fx, cfg = load(name)
    i = fx["inputs"]
    out = R.videoblip_forward(fx["state_dict"], cfg, i["input_ids"], i["attention_mask"], labels=i["labels"])
    assert torch.allclose(out["logits"], fx["text_only_logits"], atol=5e-5, rtol=1e-4)
    assert abs(float(out["loss"]) - float(fx["text_only_loss"])) < 1e-5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_gradients_match_reference_golden(name):
    fx, cfg = load(name)
    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    trainable = [k for k in sd if k in fx["grads"]]
    assert len(trainable) == len(fx["grads"]) > 0
    for k in trainable:
        sd[k].requires_grad_(True)
    out = R.videoblip_forward(sd, cfg, **fx["inputs"])
    out["loss"].backward()
    for k in trainable:
        g, ref = sd[k].grad, fx["grads"][k]
        assert g is not None, k
        assert torch.allclose(g, ref, atol=1e-6 + 1e-4 * float(ref.abs().max()), rtol=1e-3), k
    # frozen towers receive no gradient in the reference recipe (train_v2.py:124-127)
    assert not any(k.startswith(("vision_model.", "language_model.")) for k in trainable)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_greedy_generate_matches_reference_golden(name):
    fx, cfg = load(name)
    g = fx["gen_inputs"]
    toks = R.greedy_generate(fx["state_dict"], cfg, g["input_ids"], g["attention_mask"], g["pixel_values"],
                             g["video_input_mask"], max_new_tokens=fx["generated"].shape[1])
    assert torch.equal(toks, fx["generated"])


def test_opt_positions_left_and_right_padding():
    am = torch.tensor([[0, 0, 1, 1, 1], [1, 1, 1, 0, 0]])
    assert R.opt_positions(am).tolist() == [[1, 1, 2, 3, 4], [2, 3, 4, 1, 1]]


@pytest.mark.skipif(not Path("/root/reference/eilev/model/v2.py").exists(), reason="reference checkout absent")
def test_oracle_matches_live_reference_random_case():
    import types
    sys.path.insert(0, "/root/reference")
    sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
    from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel
    sys.path.insert(0, str(GOLDEN))
    import make_golden as MG

    spec = dict(MG.CONFIGS["small_opt"])
    spec.update(num_videos=2, time=1, batch=1, text=3)
    cfg = Blip2Config(**spec["config"])
    cfg.text_config.dropout = 0.0
    model = RefModel(cfg).float().eval()
    sd = R.sane_init_(model.state_dict(), seed=99, std=0.1)
    model.load_state_dict(sd)
    model.tie_weights()
    inputs = MG.build_inputs(cfg, spec, seed=3)
    with torch.no_grad():
        ref = model(**inputs, return_dict=True)
    out = R.videoblip_forward(model.state_dict(), cfg, **inputs)
    assert torch.allclose(out["logits"], ref.logits, atol=5e-5, rtol=1e-4)
    assert abs(float(out["loss"]) - float(ref.loss)) < 1e-5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_classify_matches_reference_golden(name):
    """classify (v2.py:326-501): golden scores come from the real reference's classify
    (tests/golden/make_golden_classify.py)."""
    fx, cfg = load(name)
    cx = torch.load(GOLDEN / f"classify_{name}.pt", weights_only=False)
    p = fx["gen_inputs"]
    got = R.classify(fx["state_dict"], cfg, p["input_ids"], cx["class_input_ids"], p["attention_mask"],
                     p["pixel_values"], p["video_input_mask"], cx["class_attention_mask"])
    assert got.shape == cx["scores"].shape
    assert torch.allclose(got, cx["scores"], atol=2e-4, rtol=1e-4), (got, cx["scores"])


T5_NAMES = ["small_t5", "tiny_t5_relu"]  # flan-style gated-gelu / the reference's own T5 test config (ReLU, tied head)


def _load_t5(name="small_t5"):
    fx = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    return fx, cfg


@pytest.mark.parametrize("name", T5_NAMES)
def test_oracle_t5_forward_matches_reference_golden(name):
    """flan-T5 branch (v2.py:228-238 -> T5ForConditionalGeneration): golden from the real reference
    (tests/golden/make_golden_t5.py)."""
    fx, cfg = _load_t5(name)
    out = R.videoblip_forward_t5(fx["state_dict"], cfg, **fx["inputs"])
    assert torch.allclose(out["query_output"], fx["query_output"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out["encoder_last_hidden_state"], fx["encoder_last_hidden_state"], atol=3e-4, rtol=1e-4)
    assert torch.allclose(out["logits"], fx["logits"], atol=1e-3, rtol=1e-4)
    assert abs(float(out["loss"]) - float(fx["loss"])) < 1e-4


@pytest.mark.parametrize("name", T5_NAMES)
def test_oracle_t5_gradients_match_reference_golden(name):
    fx, cfg = _load_t5(name)
    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    trainable = [k for k in sd if k in fx["grads"]]
    assert len(trainable) == len(fx["grads"]) > 0
    for k in trainable:
        sd[k].requires_grad_(True)
    out = R.videoblip_forward_t5(sd, cfg, **fx["inputs"])
    out["loss"].backward()
    for k in trainable:
        g, ref = sd[k].grad, fx["grads"][k]
        assert g is not None, k
        # this fixture's gradients are O(100): compare in relative L2 (fp32 summation order)
        assert float((g - ref).norm()) < 2e-3 * float(ref.norm()) + 1e-4, k  # (key biases have zero gradient)


@pytest.mark.parametrize("name", T5_NAMES)
def test_oracle_t5_greedy_generate_matches_reference_golden(name):
    fx, cfg = _load_t5(name)
    i = fx["inputs"]
    got = R.greedy_generate_t5(fx["state_dict"], cfg, i["input_ids"], i["attention_mask"], i["pixel_values"],
                               i["video_input_mask"], max_new_tokens=6, eos_token_id=cfg.text_config.eos_token_id)
    assert got.tolist() == fx["generated"].tolist()


# ------------------------------------------------------------------------------------- v1
V1_NAMES = ["tiny_opt", "small_opt", "small_t5"]


def load_v1(name):
    base, cfg = load(name)
    return torch.load(GOLDEN / f"v1_{name}.pt", weights_only=False), base["state_dict"], cfg


@pytest.mark.parametrize("name", V1_NAMES)
def test_oracle_v1_forward_matches_reference_golden(name):
    """The concatenation restatement of HF 4.33.1's Blip2 forward vs the real eilev.model.v1 class
    (tests/golden/make_golden_v1.py)."""
    fx, sd, cfg = load_v1(name)
    out = R.videoblip_forward_v1(sd, cfg, **fx["inputs"])
    assert out["logits"].shape == fx["logits"].shape
    assert torch.allclose(out["query_output"], fx["query_output"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out["logits"], fx["logits"], atol=5e-5, rtol=1e-4)
    assert abs(float(out["loss"]) - float(fx["loss"])) < 1e-5
    i = fx["inputs"]
    extra = {} if cfg.use_decoder_only_language_model else {"decoder_input_ids": torch.zeros(i["input_ids"].shape[0], 1, dtype=torch.long)}
    nolab = R.videoblip_forward_v1(sd, cfg, i["pixel_values"], i["input_ids"], i["attention_mask"], **extra)
    assert torch.allclose(nolab["logits"], fx["logits_no_labels"], atol=5e-5, rtol=1e-4)


@pytest.mark.parametrize("name", V1_NAMES)
def test_oracle_v1_gradients_match_reference_golden(name):
    fx, sd, cfg = load_v1(name)
    sd = {k: v.clone() for k, v in sd.items()}
    trainable = [k for k in sd if k in fx["grads"]]
    assert len(trainable) == len(fx["grads"]) > 0
    for k in trainable:
        sd[k].requires_grad_(True)
    R.videoblip_forward_v1(sd, cfg, **fx["inputs"])["loss"].backward()
    for k in trainable:
        g, ref = sd[k].grad, fx["grads"][k]
        assert torch.allclose(g, ref, atol=1e-6 + 1e-4 * float(ref.abs().max()), rtol=1e-3), k


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
def test_oracle_v1_greedy_generate_matches_reference_golden(name):
    fx, sd, cfg = load_v1(name)
    gen = R.greedy_generate_v1(sd, cfg, **fx["gen_inputs"], max_new_tokens=5)
    assert gen.tolist() == fx["generated"].tolist()
    gen = R.greedy_generate_v1(sd, cfg, fx["gen_inputs"]["pixel_values"], max_new_tokens=5)
    assert gen.tolist() == fx["generated_no_prompt"].tolist()


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
def test_v1_is_v2_with_prepended_video_slots(name):
    """The identity the CUDA v1 path rests on (eilev_b200/model/v1.py): the v2 forward on the
    ``[video slots | text]`` layout with labels[:, 0] masked reproduces v1's loss and logits, and the
    left-compacted layout ``[pad | video | text]`` reproduces v1's next-token logits."""
    from eilev_b200.model.v1 import compact_left, prepend_video_slots

    fx, sd, cfg = load_v1(name)
    i = fx["inputs"]
    ids, am, vm, lab = prepend_video_slots(i["input_ids"], i["attention_mask"], i["labels"],
                                           cfg.num_query_tokens, cfg.text_config.pad_token_id, True)
    out = R.videoblip_forward(sd, cfg, ids, am, i["pixel_values"], vm, lab)
    L = i["labels"].shape[1]
    assert torch.allclose(out["logits"][:, -L:], fx["logits"], atol=5e-5, rtol=1e-4)
    assert abs(float(out["loss"]) - float(fx["loss"])) < 1e-5
    g = fx["gen_inputs"]
    ids, am, vm, _ = prepend_video_slots(g["input_ids"], g["attention_mask"], None, cfg.num_query_tokens,
                                         cfg.text_config.pad_token_id, True)
    holes = R.videoblip_forward(sd, cfg, ids, am, g["pixel_values"], vm)["logits"][:, -1]
    cids, cam, cvm = compact_left(ids, am, vm)
    assert bool((cam[:, 1:] >= cam[:, :-1]).all())  # left padded: zeros then ones
    assert int(cvm.sum()) == int(vm.sum()) and bool((cam[cvm.bool()] == 1).all())
    compact = R.videoblip_forward(sd, cfg, cids, cam, g["pixel_values"], cvm)["logits"][:, -1]
    assert torch.allclose(holes, compact, atol=5e-5, rtol=1e-4)
    assert compact.argmax(-1).tolist() == fx["generated"][:, 0].tolist()


# ------------------------------------------------------------------------------------- frames
def test_oracle_normalize_frames_is_the_hf_rescale_and_normalize():
    """Bit-exact against the numpy ``rescale`` / ``normalize`` the reference's pinned BlipImageProcessor
    calls (transformers/image_transforms.py, unchanged in the installed version), and within 2 fp32
    ulp (identical after bf16 rounding) of the installed torchvision-backed BlipImageProcessor."""
    import numpy as np
    from transformers import BlipImageProcessor
    from transformers import image_transforms as IT

    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (2, 3, 4, 32, 48), dtype=torch.uint8, generator=g)  # (N, C, T, H, W)
    got = R.normalize_frames(frames)
    flat = frames.permute(0, 2, 1, 3, 4).reshape(-1, 3, 32, 48)
    want = np.stack([IT.normalize(IT.rescale(f.numpy(), 1 / 255), R.OPENAI_CLIP_MEAN, R.OPENAI_CLIP_STD,
                                  data_format="channels_first", input_data_format="channels_first") for f in flat])
    want = torch.from_numpy(want).view(2, 4, 3, 32, 48).permute(0, 2, 1, 3, 4)
    assert want.dtype == torch.float32 and torch.equal(got, want)
    ip = BlipImageProcessor(size={"height": 32, "width": 48})
    assert tuple(ip.image_mean) == R.OPENAI_CLIP_MEAN and tuple(ip.image_std) == R.OPENAI_CLIP_STD
    live = ip(images=flat, return_tensors="pt").pixel_values.view(2, 4, 3, 32, 48).permute(0, 2, 1, 3, 4)
    assert float((got - live).abs().max()) < 1e-6
    assert float((got.bfloat16() == live.bfloat16()).float().mean()) > 0.9999


def test_oracle_pil_resize_is_bit_exact_with_pillow():
    """oracle/pil_resize_ref.py (restatement of Pillow's Resample.c 8-bit bicubic) against the real
    ``PIL.Image.resize(size, BICUBIC)`` — the resize of the reference's ``process`` — over down-/up-
    scaling, non-square, odd sizes, identity, RGB (bands are independent) and saturating inputs."""
    import numpy as np
    from PIL import Image

    from oracle import pil_resize_ref as P
    rs = np.random.RandomState(0)
    for (h, w), (oh, ow) in [((448, 448), (224, 224)), ((100, 160), (224, 224)), ((224, 224), (224, 224)),
                             ((37, 53), (56, 56)), ((300, 500), (224, 224)), ((224, 300), (224, 224)),
                             ((301, 224), (224, 224)), ((17, 19), (64, 48)), ((720, 1280), (224, 224))]:
        x = rs.randint(0, 256, (h, w)).astype(np.uint8)
        want = np.asarray(Image.fromarray(x).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(P.resize_bicubic_u8(x, oh, ow), want), ((h, w), (oh, ow))
    rgb = rs.randint(0, 256, (120, 90, 3)).astype(np.uint8)
    want = np.asarray(Image.fromarray(rgb).resize((224, 224), resample=Image.BICUBIC))
    assert np.array_equal(np.moveaxis(P.resize_bicubic_u8(np.moveaxis(rgb, -1, 0), 224, 224), 0, -1), want)
    binary = ((rs.rand(64, 64) > 0.5) * 255).astype(np.uint8)  # overshoot of the negative lobes clips at 0 / 255
    assert np.array_equal(P.resize_bicubic_u8(binary, 224, 224),
                          np.asarray(Image.fromarray(binary).resize((224, 224), resample=Image.BICUBIC)))
