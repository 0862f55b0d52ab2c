"""Parity of the CUDA path (eilev_b200.model.v2 through the C ABI) with the CPU oracle and
with the golden outputs of the real reference, on identical seeded weights and inputs.

Tolerances (bf16 kernels with fp32 accumulation vs an fp32 oracle) are stated per test;
the yardstick is the reference's own bf16-vs-fp32 gap (logits rel-L2
1.6 %, max-abs 0.083 at logit std 1)."""
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


def max_abs(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max())


def _dump(key, **vals):
    REPORT[key] = vals
    out = Path(os.environ.get("GRAFT_REPO_ROOT", ".")) / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_report.json").write_text(json.dumps(REPORT, indent=1))


def load(name):
    fx = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    return fx, cfg


def build(cfg, sd, dtype=torch.float32):
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(sd)
    return m.to("cuda", dtype).eval()


def cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
def test_forward_matches_reference_golden(name):
    fx, cfg = load(name)
    m = build(cfg, fx["state_dict"])
    with torch.no_grad():
        out = m(**cuda(fx["inputs"]), return_dict=True)
    m.check_splice()
    r = dict(
        image_embeds=rel_l2(out.vision_outputs.last_hidden_state, fx["image_embeds"]),
        pooler=rel_l2(out.vision_outputs.pooler_output, fx["pooler_output"]),
        query_output=rel_l2(out.qformer_outputs.last_hidden_state, fx["query_output"]),
        logits=rel_l2(out.logits, fx["logits"]),
        logits_max_abs=max_abs(out.logits, fx["logits"]),
        logits_ref_absmax=float(fx["logits"].abs().max()),
        loss=float(out.loss), loss_ref=float(fx["loss"]),
    )
    _dump(f"forward/{name}", **r)
    assert out.logits.shape == fx["logits"].shape
    assert r["image_embeds"] < 0.02 and r["pooler"] < 0.02
    assert r["query_output"] < 0.03
    assert r["logits"] < 0.03, r
    assert r["logits_max_abs"] < 0.03 * r["logits_ref_absmax"] + 0.05, r
    assert abs(r["loss"] - r["loss_ref"]) < 0.03, r


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
def test_text_only_forward(name):
    fx, cfg = load(name)
    m = build(cfg, fx["state_dict"])
    i = cuda(fx["inputs"])
    with torch.no_grad():
        out = m(i["input_ids"], attention_mask=i["attention_mask"], labels=i["labels"], return_dict=True)
    assert out.vision_outputs is None and out.qformer_outputs is None
    assert rel_l2(out.logits, fx["text_only_logits"]) < 0.03
    assert abs(float(out.loss) - float(fx["text_only_loss"])) < 0.03


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
def test_backward_matches_reference_golden(name):
    """train_v2.py:123-130 recipe: ViT + LM frozen, Q-Former / query_tokens / projection train."""
    fx, cfg = load(name)
    m = build(cfg, fx["state_dict"]).train()
    for p in m.vision_model.parameters():
        p.requires_grad = False
    for p in m.language_model.parameters():
        p.requires_grad = False
    m.enable_input_require_grads()
    out = m(**cuda(fx["inputs"]), return_dict=True)
    out.loss.backward()
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(fx["grads"]), (set(got) ^ set(fx["grads"]))
    total_num = total_den = 0.0
    worst = ("", 0.0)
    for n, ref in fx["grads"].items():
        g = got[n].float().cpu()
        assert g.shape == ref.shape, n
        total_num += float((g - ref).pow(2).sum())
        total_den += float(ref.pow(2).sum())
        r = rel_l2(g, ref)
        if float(ref.norm()) > 1e-3 * (total_den ** 0.5 + 1e-12) and r > worst[1]:
            worst = (n, r)
    glob = (total_num / total_den) ** 0.5
    _dump(f"backward/{name}", global_rel_l2=glob, worst=worst, loss=float(out.loss), n=len(got))
    # bf16 gradients through a 4..8-wide toy network are noisy (softmax / LayerNorm backward
    # cancel large terms); the real-dims test below holds the tight bound
    assert glob < 0.09, (glob, worst)
    assert worst[1] < 0.25, worst


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
def test_generate_greedy_matches_reference_golden(name):
    fx, cfg = load(name)
    m = build(cfg, fx["state_dict"])
    g = cuda(fx["gen_inputs"])
    n_new = fx["generated"].shape[1]
    toks = m.generate(**g, max_new_tokens=n_new, min_new_tokens=n_new, do_sample=False, num_beams=1)
    assert toks.shape == fx["generated"].shape
    agree = float((toks.cpu() == fx["generated"]).float().mean())
    _dump(f"generate/{name}", agree=agree, got=toks.tolist(), ref=fx["generated"].tolist())
    # bf16 rounding may flip near-ties of a random-init model; the first token must match
    assert torch.equal(toks[:, 0].cpu(), fx["generated"][:, 0])
    assert agree >= 0.6


def test_vision_model_standalone_shapes_and_bf16():
    from eilev_b200.model.v2 import VideoBlipVisionModel
    fx, cfg = load("small_opt")
    vm = VideoBlipVisionModel(cfg.vision_config)
    vm.load_state_dict({k[len("vision_model."):]: v for k, v in fx["state_dict"].items() if k.startswith("vision_model.")})
    vm = vm.to("cuda", torch.bfloat16).eval()
    px = fx["inputs"]["pixel_values"].cuda()
    out = vm(px, output_hidden_states=True, return_dict=True)
    n, t = px.shape[0], px.shape[2]
    s = (cfg.vision_config.image_size // cfg.vision_config.patch_size) ** 2 + 1
    assert out.last_hidden_state.shape == (n, t * s, cfg.vision_config.hidden_size)
    assert out.pooler_output.shape == (n, t, cfg.vision_config.hidden_size)
    assert len(out.hidden_states) == cfg.vision_config.num_hidden_layers + 1
    assert out.last_hidden_state.dtype == torch.bfloat16
    assert rel_l2(out.last_hidden_state, fx["image_embeds"]) < 0.03
    tup = vm(px, return_dict=False)
    assert len(tup) == 4 and tup[2] is None and tup[3] is None
    with pytest.raises(ValueError):
        vm(None)


def test_output_attentions_shapes_and_values():
    """output_attentions=True (eilev/model/v2.py:78-95; the reference's own test parametrises over it,
    tests/model/test_model_v2.py:8-83): one (num_videos, time, heads, S, S) map per layer, rows sum to 1, equal to
    the oracle's softmax(q k^T / sqrt(d)) of the first layer; the full model accepts the flag."""
    from eilev_b200.model.v2 import VideoBlipVisionModel
    from oracle import videoblip_ref as R
    fx, cfg = load("small_opt")
    vm = VideoBlipVisionModel(cfg.vision_config)
    sd = {k[len("vision_model."):]: v for k, v in fx["state_dict"].items() if k.startswith("vision_model.")}
    vm.load_state_dict(sd)
    vm = vm.to("cuda").eval()
    px = fx["inputs"]["pixel_values"]
    n, _, t, _, _ = px.shape
    last, pooled, hidden, attn = vm(pixel_values=px.cuda(), output_attentions=True, output_hidden_states=True,
                                    return_dict=False)
    vc = cfg.vision_config
    s = (vc.image_size // vc.patch_size) ** 2 + 1
    assert len(attn) == vc.num_hidden_layers and len(hidden) == vc.num_hidden_layers + 1
    for a in attn:
        assert a.shape == (n, t, vc.num_attention_heads, s, s)
        assert torch.allclose(a.float().sum(-1), torch.ones(n, t, vc.num_attention_heads, s, device="cuda"), atol=1e-3)
    # first layer against the oracle's functions
    frames = px.permute(0, 2, 1, 3, 4).flatten(end_dim=1)
    x = R.vision_embeddings(fx["state_dict"], vc, frames)
    p0 = "vision_model.encoder.layers.0."
    y = R._ln(x, fx["state_dict"], p0 + "layer_norm1", vc.layer_norm_eps)
    h, d = vc.num_attention_heads, vc.hidden_size // vc.num_attention_heads
    qkv = R._lin(y, fx["state_dict"], p0 + "self_attn.qkv").reshape(n * t, s, 3, h, d).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * d ** -0.5, dim=-1).view(n, t, h, s, s)
    assert (attn[0].float().cpu() - ref).abs().max().item() < 2e-2  # bf16 q, k vs the fp32 oracle, probabilities up to ~0.5
    # the maps do not disturb the fused path — bit for bit: the folded LayerNorms' row statistics are summed by
    # atomics in no fixed order, but in f64, where adding the f32 partials is exact
    last2 = vm(pixel_values=px.cuda(), return_dict=False)[0]
    assert torch.equal(last, last2)
    m = build(cfg, fx["state_dict"])
    with torch.no_grad():
        out = m(**cuda(fx["inputs"]), output_attentions=True, return_dict=True)
    assert out.logits.shape == fx["logits"].shape and len(out.vision_outputs.attentions) == vc.num_hidden_layers


def test_errors_match_reference_contract():
    fx, cfg = load("tiny_opt")
    m = build(cfg, fx["state_dict"])
    i = cuda(fx["inputs"])
    with pytest.raises(AssertionError):  # v2.py:154-157
        m(i["input_ids"], pixel_values=i["pixel_values"])
    with pytest.raises(Exception):
        m(i["input_ids"].cpu())
    bad = i["video_input_mask"].clone()
    bad[0, 1] = 0
    with torch.no_grad():
        m(i["input_ids"], attention_mask=i["attention_mask"], pixel_values=i["pixel_values"], video_input_mask=bad)
    with pytest.raises(RuntimeError):
        m.check_splice()


def test_train_mode_dropout_matches_oracle_with_injected_masks():
    """Recipe mode (dropout p = 0.1 in the Q-Former and the frozen-but-training OPT): the masks
    are counter hashes, so the oracle can replay them: loss, logits and gradients must agree."""
    from eilev_b200 import ops
    from eilev_b200.engine import opt as E_opt, qformer as E_qf
    from oracle import videoblip_ref as R
    fx, cfg = load("small_opt")
    cfg.qformer_config.hidden_dropout_prob = 0.1
    cfg.qformer_config.attention_probs_dropout_prob = 0.1
    cfg.text_config.dropout = 0.1
    cfg.text_config.attention_dropout = 0.1
    m = build(cfg, fx["state_dict"]).train()
    for p_ in m.vision_model.parameters():
        p_.requires_grad = False
    for p_ in m.language_model.parameters():
        p_.requires_grad = False
    # the masks (and with them the size of the bf16 gradient gap: 6.7 % .. 12.4 % were seen) depend on the
    # process-wide torch seed the model derives its dropout seed from: pin it
    m._dropout_seed = torch.full((1,), 20240607, dtype=torch.int64, device="cuda")
    out = m(**cuda(fx["inputs"]), return_dict=True)
    out.loss.backward()
    seed = m._dropout_seed

    def drop(site, t):
        if site is None:
            return t
        tower, layer, k = site
        salt = (E_qf._SALT_QF if tower == "qf" else E_opt._SALT_OPT) + (layer * 8 + k if layer >= 0 else k)
        rows = t.numel() // t.shape[-1]
        mask = ops.dropout(torch.ones(rows, t.shape[-1], dtype=torch.bfloat16, device="cuda"), 0.1, seed, salt)
        return t * mask.float().cpu().view(t.shape)

    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    for k in fx["grads"]:
        sd[k].requires_grad_(True)
    ref = R.videoblip_forward(sd, cfg, **fx["inputs"], drop=drop)
    ref["loss"].backward()
    assert abs(float(out.loss) - float(ref["loss"])) < 0.05, (float(out.loss), float(ref["loss"]))
    assert abs(float(ref["loss"]) - float(fx["loss"])) > 1e-3  # the masks really changed the computation
    assert rel_l2(out.logits, ref["logits"]) < 0.03
    num = den = 0.0
    for n_, p_ in m.named_parameters():
        if p_.grad is not None:
            rg = sd[n_].grad
            num += float((p_.grad.float().cpu() - rg).pow(2).sum())
            den += float(rg.pow(2).sum())
    glob = (num / den) ** 0.5
    _dump("dropout/small_opt", loss=float(out.loss), loss_ref=float(ref["loss"]), grad_rel_l2=glob,
          logits=rel_l2(out.logits, ref["logits"]))
    # the reference's own bf16-vs-fp32 gradient gap on this fixture WITHOUT dropout is 8.9-9.6 %
    # (tests/golden/bf16_yardstick.py); with 10 % of the activations masked the gap of the same kernels moves
    # between 6.7 % and 12.4 % with the mask draw
    assert glob < 0.14, glob
    # a second step draws different masks; eval mode is deterministic and mask-free
    out2 = m(**cuda(fx["inputs"]), return_dict=True)
    assert abs(float(out2.loss) - float(out.loss)) > 1e-4
    m.eval()
    with torch.no_grad():
        e1 = m(**cuda(fx["inputs"]), return_dict=True)
    assert abs(float(e1.loss) - float(fx["loss"])) < 0.03


REAL_DIMS = dict(
    vision_config=dict(hidden_size=1408, intermediate_size=6144, num_hidden_layers=2, num_attention_heads=16,
                       patch_size=14, image_size=224),
    qformer_config=dict(hidden_size=768, num_hidden_layers=2, num_attention_heads=12, intermediate_size=3072,
                        encoder_hidden_size=1408, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
    text_config=dict(model_type="opt", hidden_size=2560, num_hidden_layers=2, ffn_dim=10240,
                     num_attention_heads=32, vocab_size=50272, max_position_embeddings=2048,
                     word_embed_proj_dim=2560, dropout=0.0, attention_dropout=0.0),
    num_query_tokens=32)


def test_real_dims_shallow_against_oracle():
    """Real BLIP-2 / OPT-2.7B layer shapes (S=257, d=88, d=80, vocab 50272), 2 layers per tower:
    exercises the tcgen05 tiles, the 96/80-wide attention kernels and the fused CE."""
    from oracle import videoblip_ref as R
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    torch.manual_seed(0)
    cfg = Blip2Config(**REAL_DIMS)
    m = VideoBlipForConditionalGeneration(cfg)
    sd = R.sane_init_({k: v.clone() for k, v in m.state_dict().items()}, seed=5, std=0.02)
    sd["language_model.lm_head.weight"] = sd["language_model.model.decoder.embed_tokens.weight"]  # tied
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    nv, t, nq = 2, 2, 32
    px = torch.randn(nv, 3, t, 224, 224, generator=g)
    ids, vm, lab = [2], [0], [-100]
    for _ in range(nv):
        ids += [1] * nq + [50118] + torch.randint(4, 50000, (5,), generator=g).tolist()
        vm += [1] * nq + [0] * 6
        lab += [-100] * (nq + 6)
    tgt = torch.randint(4, 50000, (4,), generator=g).tolist()
    ids += tgt; vm += [0] * 4; lab += tgt
    pad = (-len(ids)) % 8
    attn = [1] * len(ids) + [0] * pad
    ids += [1] * pad; vm += [0] * pad; lab += [-100] * pad
    inputs = dict(input_ids=torch.tensor([ids]), attention_mask=torch.tensor([attn]),
                  pixel_values=px, video_input_mask=torch.tensor([vm]), labels=torch.tensor([lab]))
    trainable = [k for k in sd if k.startswith(("qformer.", "query_tokens", "language_projection."))]
    sdg = {k: v.clone() for k, v in sd.items()}
    for k in trainable:
        sdg[k].requires_grad_(True)
    ref = R.videoblip_forward(sdg, cfg, **inputs)
    ref["loss"].backward()

    m = m.to("cuda", torch.bfloat16).train()
    for p in m.vision_model.parameters():
        p.requires_grad = False
    for p in m.language_model.parameters():
        p.requires_grad = False
    out = m(**cuda(inputs), return_dict=True)
    out.loss.backward()
    valid = torch.tensor(attn).bool()
    r = dict(
        image_embeds=rel_l2(out.vision_outputs.last_hidden_state, ref["image_embeds"]),
        query_output=rel_l2(out.qformer_outputs.last_hidden_state, ref["query_output"]),
        logits=rel_l2(out.logits[0, valid], ref["logits"][0, valid]),
        logits_max_abs=max_abs(out.logits[0, valid], ref["logits"][0, valid]),
        logits_std=float(ref["logits"].std()),
        loss=float(out.loss), loss_ref=float(ref["loss"]),
    )
    num = den = 0.0
    for n_, p in m.named_parameters():
        if p.grad is not None:
            rg = sdg[n_].grad
            num += float((p.grad.float().cpu() - rg).pow(2).sum())
            den += float(rg.pow(2).sum())
    r["grad_rel_l2"] = (num / den) ** 0.5
    _dump("real_dims_shallow", **r)
    assert r["image_embeds"] < 0.02 and r["query_output"] < 0.03
    assert r["logits"] < 0.025, r            # BASELINE.md §2.1 stated tolerance: rel-L2 <= 2.5 %
    assert r["logits_max_abs"] < 0.15, r     # and max-abs <= 0.15 at logit std ~1
    assert abs(r["loss"] - r["loss_ref"]) < 0.05, r
    # yardstick: the reference itself, bf16 vs fp32 on CPU (tests/golden/bf16_yardstick.py),
    # differs by 4.3 % (tiny_opt) / 8.9-9.6 % (small_opt) in global gradient rel-L2
    assert r["grad_rel_l2"] < 0.10, r


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
@pytest.mark.parametrize("class_batch_size", [None, 2])
def test_classify_matches_reference_golden(name, class_batch_size):
    """classify (v2.py:326-501): mean class log-likelihoods vs the real reference's scores.
    Tolerance: |d| <= 0.08 nats on scores of magnitude 3-7 (bf16 LM, fp32 log-softmax)."""
    fx, cfg = load(name)
    cx = torch.load(GOLDEN / f"classify_{name}.pt", weights_only=False)
    m = build(cfg, fx["state_dict"])
    p = cuda(fx["gen_inputs"])
    got = m.classify(p["input_ids"], cx["class_input_ids"].cuda(), p["attention_mask"], p["pixel_values"],
                     p["video_input_mask"], cx["class_attention_mask"].cuda(), class_batch_size=class_batch_size)
    m.check_splice()
    assert got.shape == cx["scores"].shape
    d = max_abs(got, cx["scores"])
    _dump(f"classify/{name}/{class_batch_size}", max_abs=d, ref=cx["scores"].tolist(), got=got.cpu().tolist())
    assert d < 0.08, (got, cx["scores"])
    # text-only prompt and default masks
    got2 = m.classify(p["input_ids"], cx["class_input_ids"].cuda(), p["attention_mask"])
    assert got2.shape == cx["scores"].shape and torch.isfinite(got2).all()
    with pytest.raises(AssertionError):  # v2.py:355
        m.classify(p["input_ids"], cx["class_input_ids"].cuda(), pixel_values=p["pixel_values"])


@pytest.mark.parametrize("batch", [1, 3, 8])
def test_persistent_decode_step_matches_op_by_op(batch, monkeypatch):
    """vb_decode_step (one cooperative launch per token: embed, per-layer projections,
    paged attention, head, grid barriers in between) against the same step issued op by op,
    on a left-padded batch, eagerly and through the CUDA graph.  The two paths share the
    kernels' arithmetic, so the logits agree to bf16 rounding of the split-merge order."""
    from transformers import OPTConfig
    from eilev_b200.engine import opt as E
    from eilev_b200.model.v2 import OPTForCausalLM
    from oracle import videoblip_ref as R

    monkeypatch.setenv("VB_DECODE_PERSISTENT", "1")
    torch.manual_seed(0)
    cfg = OPTConfig(hidden_size=128, num_hidden_layers=3, ffn_dim=256, num_attention_heads=2, vocab_size=300,
                    max_position_embeddings=256, word_embed_proj_dim=128)
    lm = OPTForCausalLM(cfg)
    sd = R.sane_init_({k: v.clone() for k, v in lm.state_dict().items()}, seed=9, std=0.08)
    sd["lm_head.weight"] = sd["model.decoder.embed_tokens.weight"]
    lm.load_state_dict(sd)
    lm = lm.to("cuda", torch.bfloat16).eval()
    g = torch.Generator().manual_seed(batch)
    L = 70
    ids = torch.randint(4, 290, (batch, L), generator=g)
    am = torch.ones(batch, L, dtype=torch.long)
    for b in range(1, batch):
        am[b, : 3 * b] = 0  # left padding
        ids[b, : 3 * b] = 1
    ids, am = ids.cuda(), am.cuda()
    steps = 5
    forced = torch.randint(4, 290, (steps, batch), generator=g).cuda()  # same tokens on every path
    outs = {}
    for mode in ("persistent", "op_by_op", "graph"):
        with torch.no_grad():
            logits, state = E.opt_prefill(lm, lm._pack, ids, am, None, None, steps + 2)
            if mode == "op_by_op":
                state["program"] = None
            else:
                assert E._decode_program(lm, lm._pack, state, batch, ids.device) is not None
            runner = E.DecodeGraph(lm, lm._pack, state, batch, ids.device) if mode == "graph" else None
            seq = [logits.clone()]
            for i in range(steps):
                tok = forced[i]
                logits = runner.step(tok) if runner is not None else E.opt_decode_step(lm, lm._pack, tok, state)
                seq.append(logits.clone())
            outs[mode] = (torch.stack(seq), state["ctx_len"].clone(), state["n_valid"].clone())
    ref = outs["op_by_op"]
    for mode in ("persistent", "graph"):
        got = outs[mode]
        assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2])
        d = max_abs(got[0], ref[0])
        _dump(f"persistent_decode/{mode}/b{batch}", max_abs=d, ref_absmax=float(ref[0].abs().max()))
        assert d < 0.03 * float(ref[0].abs().max()) + 0.02, (mode, d)


def _load_t5():
    fx = torch.load(GOLDEN / "small_t5.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    return fx, cfg


def test_t5_forward_matches_reference_golden():
    """flan-T5 branch (v2.py:228-238): logits / loss / encoder output vs the real reference (fp32).
    Yardstick measured on this fixture (tests/golden/make_golden_t5.py shrinks the random q / k
    projections, T5 attention being unscaled): the reference's OWN bf16 run differs from its fp32 run
    by 1.3 % (logits rel-L2) / 1.3 % (encoder output) / 2.2 % (gradients).  Tolerance: logits and
    encoder output <= 3 %, loss |d| <= 0.03."""
    fx, cfg = _load_t5()
    m = build(cfg, fx["state_dict"])
    with torch.no_grad():
        out = m(**cuda(fx["inputs"]), return_dict=True)
    m.check_splice()
    valid = fx["inputs"]["attention_mask"].bool()
    r = dict(logits=rel_l2(out.logits, fx["logits"]), logits_max_abs=max_abs(out.logits, fx["logits"]),
             logits_ref_absmax=float(fx["logits"].abs().max()),
             enc=rel_l2(out.language_model_outputs.encoder_last_hidden_state.cpu()[valid],
                        fx["encoder_last_hidden_state"][valid]),
             loss=float(out.loss), loss_ref=float(fx["loss"]))
    _dump("t5_forward", **r)
    assert out.logits.shape == fx["logits"].shape
    assert r["enc"] < 0.03 and r["logits"] < 0.03, r
    assert abs(r["loss"] - r["loss_ref"]) < 0.03, r


def test_t5_backward_matches_reference_golden():
    """Gradients of the trainable (Q-Former side) tensors through the frozen T5 (decoder, cross
    K/V projection, encoder, splice): global rel-L2 <= 6 % (the reference's own bf16-vs-fp32 gradient
    gap on this fixture is 2.2 %)."""
    from eilev_b200.train import freeze_for_recipe
    fx, cfg = _load_t5()
    m = build(cfg, fx["state_dict"]).train()
    freeze_for_recipe(m)
    out = m(**cuda(fx["inputs"]), return_dict=True)
    out.loss.backward()
    got = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(fx["grads"])
    num = den = 0.0
    worst = ("", 0.0)
    for n, g in got.items():
        ref = fx["grads"][n]
        num += float((g.float().cpu() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
        e = rel_l2(g, ref)
        if float(ref.norm()) > 1e-3 * den ** 0.5 and e > worst[1]:
            worst = (n, e)
    glob = (num / den) ** 0.5
    _dump("t5_backward", global_rel_l2=glob, worst=worst, loss=float(out.loss.detach()), n=len(got))
    assert glob < 0.06, (glob, worst)


def test_t5_train_mode_dropout_matches_oracle_with_injected_masks():
    """Recipe mode for the seq2seq branch: T5's own dropout_rate (attention probabilities, attention / cross-
    attention / feed-forward outputs, the feed-forward inner activation, embeddings and stack outputs of both
    stacks — every site HF T5 has) plus the Q-Former's dropout.  The counter-hash masks are replayed in the
    oracle: loss, logits and gradients must agree (the frozen T5 runs in train() under HF Trainer:
    eilev/model/v2.py:228-238, scripts/general/train_v2.py:123-130)."""
    from eilev_b200 import ops
    from eilev_b200.engine import qformer as E_qf, t5 as E_t5
    from eilev_b200.train import freeze_for_recipe
    from oracle import videoblip_ref as R
    fx, cfg = _load_t5()
    cfg.qformer_config.hidden_dropout_prob = 0.1
    cfg.qformer_config.attention_probs_dropout_prob = 0.1
    cfg.text_config.dropout_rate = 0.1
    m = build(cfg, fx["state_dict"]).train()
    freeze_for_recipe(m)
    m._dropout_seed = torch.full((1,), 777, dtype=torch.int64, device="cuda")
    out = m(**cuda(fx["inputs"]), return_dict=True)
    out.loss.backward()
    seed = m._dropout_seed
    used = set()

    def drop(site, t):
        if site is None:
            return t
        tower, layer, k = site
        salt = {"qf": E_qf._SALT_QF + (layer * 8 + k if layer >= 0 else k),
                "t5e": E_t5._SALT_T5 + layer * 8 + k,
                "t5d": E_t5._SALT_T5 + E_t5._T5_DEC + layer * 8 + k,
                "t5": E_t5._SALT_T5 + E_t5._T5_STACK + k}[tower]
        used.add(tower)
        rows = t.numel() // t.shape[-1]
        mask = ops.dropout(torch.ones(rows, t.shape[-1], dtype=torch.bfloat16, device="cuda"), 0.1, seed, salt)
        return t * mask.float().cpu().view(t.shape)

    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    for k in fx["grads"]:
        sd[k].requires_grad_(True)
    ref = R.videoblip_forward_t5(sd, cfg, **fx["inputs"], drop=drop)
    ref["loss"].backward()
    assert used == {"qf", "t5e", "t5d", "t5"}, used
    assert abs(float(ref["loss"]) - float(fx["loss"])) > 1e-3  # the masks really changed the computation
    num = den = 0.0
    for n_, p_ in m.named_parameters():
        if p_.grad is not None:
            rg = sd[n_].grad
            num += float((p_.grad.float().cpu() - rg).pow(2).sum())
            den += float(rg.pow(2).sum())
    glob = (num / den) ** 0.5
    r = dict(loss=float(out.loss.detach()), loss_ref=float(ref["loss"]), logits=rel_l2(out.logits, ref["logits"]),
             grad_rel_l2=glob)
    _dump("dropout/small_t5", **r)
    assert abs(r["loss"] - r["loss_ref"]) < 0.05, r
    assert r["logits"] < 0.03, r
    assert glob < 0.08, r   # no-dropout gap of this fixture: 1.3 % (test_t5_backward_matches_reference_golden)
    # eval mode stays mask-free
    m.eval()
    with torch.no_grad():
        e1 = m(**cuda(fx["inputs"]), return_dict=True)
    assert abs(float(e1.loss) - float(fx["loss"])) < 0.03


def test_t5_generate_matches_reference_golden_and_text_only():
    """generate() with the seq2seq LM (v2.py:254-324): [decoder_start] + greedy tokens, token-exact
    on the fixture; beam search and sampling run; text-only forward works."""
    fx, cfg = _load_t5()
    m = build(cfg, fx["state_dict"])
    i = cuda(fx["inputs"])
    gen = m.generate(i["input_ids"], i["pixel_values"], i["video_input_mask"], i["attention_mask"],
                     max_new_tokens=6, do_sample=False)
    assert gen.cpu().tolist() == fx["generated"].tolist(), (gen.cpu().tolist(), fx["generated"].tolist())
    beams = m.generate(i["input_ids"], i["pixel_values"], i["video_input_mask"], i["attention_mask"],
                       max_new_tokens=5, num_beams=3)
    assert beams.shape[0] == 2 and beams.shape[1] <= 6 and int(beams[0, 0]) == cfg.text_config.decoder_start_token_id
    with torch.no_grad():
        out = m(i["input_ids"], attention_mask=i["attention_mask"], labels=i["labels"], return_dict=True)
    assert torch.isfinite(out.loss)


def test_t5_real_dims_shallow_against_oracle():
    """flan-t5-xl layer shapes (d_model 2048, 32 heads x 64, d_ff 5120, vocab 32128) with 2 + 2
    layers behind the real-width ViT / Q-Former: exercises the tcgen05 tiles on the T5 GEMM
    shapes (6144 / 2048 / 10240 wide), the biased flash attention at L = 160 and the
    cross-attention K|V batching.  q / k projections are shrunk as in the golden fixture
    (unscaled attention).  Tolerance: logits rel-L2 <= 2.5 %, loss |d| <= 0.05, grads <= 10 %."""
    from oracle import videoblip_ref as R
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import freeze_for_recipe
    torch.manual_seed(0)
    spec = dict(REAL_DIMS)
    spec["text_config"] = dict(model_type="t5", d_model=2048, d_kv=64, d_ff=5120, num_layers=2, num_decoder_layers=2,
                               num_heads=32, vocab_size=32128, feed_forward_proj="gated-gelu",
                               tie_word_embeddings=False, decoder_start_token_id=0, pad_token_id=0, eos_token_id=1,
                               dropout_rate=0.0)
    cfg = Blip2Config(**spec)
    m = VideoBlipForConditionalGeneration(cfg)
    sd = R.sane_init_({k: v.clone() for k, v in m.state_dict().items()}, seed=6, std=0.02)
    for k in sd:
        if k.startswith("language_model.") and k.endswith((".q.weight", ".k.weight")):
            sd[k] = sd[k] * 0.25
    sd["language_model.encoder.embed_tokens.weight"] = sd["language_model.shared.weight"]
    sd["language_model.decoder.embed_tokens.weight"] = sd["language_model.shared.weight"]
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(4)
    nv, t, nq = 2, 2, 32
    px = torch.randn(nv, 3, t, 224, 224, generator=g)
    ids, vm = [], []
    for _ in range(nv):
        ids += [0] * nq + [3] + torch.randint(4, 32000, (43,), generator=g).tolist()
        vm += [1] * nq + [0] * 44
    ids += [1]; vm += [0]
    pad = (-len(ids)) % 8
    attn = [1] * len(ids) + [0] * pad
    ids += [0] * pad; vm += [0] * pad
    labels = torch.randint(4, 32000, (1, 10), generator=g)
    labels[0, 8:] = -100
    inputs = dict(input_ids=torch.tensor([ids]), attention_mask=torch.tensor([attn]), pixel_values=px,
                  video_input_mask=torch.tensor([vm]), labels=labels)
    trainable = [k for k in sd if k.startswith(("qformer.", "query_tokens", "language_projection."))]
    sdg = {k: v.clone() for k, v in sd.items()}
    for k in trainable:
        sdg[k].requires_grad_(True)
    ref = R.videoblip_forward_t5(sdg, cfg, **inputs)
    ref["loss"].backward()

    m = m.to("cuda", torch.bfloat16).train()
    freeze_for_recipe(m)
    out = m(**cuda(inputs), return_dict=True)
    out.loss.backward()
    r = dict(logits=rel_l2(out.logits, ref["logits"]), logits_max_abs=max_abs(out.logits, ref["logits"]),
             logits_std=float(ref["logits"].std()), loss=float(out.loss.detach()), loss_ref=float(ref["loss"]))
    num = den = 0.0
    for n_, p in m.named_parameters():
        if p.grad is not None:
            rg = sdg[n_].grad
            num += float((p.grad.float().cpu() - rg).pow(2).sum())
            den += float(rg.pow(2).sum())
    r["grad_rel_l2"] = (num / den) ** 0.5
    _dump("t5_real_dims_shallow", **r)
    assert r["logits"] < 0.025, r
    assert abs(r["loss"] - r["loss_ref"]) < 0.05, r
    assert r["grad_rel_l2"] < 0.10, r


def test_trainer_graphs_and_inplace_grad_accumulation_match_plain_autograd():
    """DataParallelTrainer (flat f32 grad views filled in place by the wgrad kernels, one CUDA
    graph with the Q-Former re-pack + one without) against plain autograd on a second copy of
    the model: accumulated gradients after 2 micro-steps and parameters after the optimizer
    step, then one more micro-step through the re-pack graph."""
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import DataParallelTrainer, freeze_for_recipe
    fx, cfg = load("small_opt")

    def fresh():
        m = VideoBlipForConditionalGeneration(cfg)
        m.load_state_dict(fx["state_dict"])
        m = m.to("cuda").train()
        freeze_for_recipe(m)
        return m

    batch = cuda(fx["inputs"])
    ref = fresh()
    tr_model = fresh()
    tr = DataParallelTrainer(tr_model, lr=1e-3, weight_decay=0.05, max_grad_norm=1.0, grad_accum=2)
    tr.capture_graph(batch)
    names = [n for n, p in ref.named_parameters() if p.requires_grad]
    # the q / k / v weight (and bias) gradients of a layer sit back to back in the flat buffer, so the fused
    # projection's wgrad GEMM accumulates into them directly (engine/qformer.py::_GradOut.weights_fused)
    from eilev_b200.engine import qformer as E_qf
    sink = {n: p.grad for n, p in E_qf.qformer_param_list(tr_model) if p.requires_grad and p.grad is not None}
    pre = "qformer.encoder.layer.0.attention.attention."
    gout = E_qf._GradOut(sink, torch.device("cuda"))
    wv = gout._fused_view([pre + f"{nm}.weight" for nm in ("query", "key", "value")])
    bv = gout._fused_view([pre + f"{nm}.bias" for nm in ("query", "key", "value")])
    assert wv is not None and wv.shape[0] == 3 * sink[pre + "query.weight"].shape[0]
    assert bv is not None and bv.data_ptr() == sink[pre + "query.bias"].data_ptr()

    def ref_micro():
        out = ref(**batch, return_dict=True)
        (out.loss / 2).backward()
        return float(out.loss.detach())

    l_ref = [ref_micro(), ref_micro()]
    l_tr = [float(tr.micro_step(batch)) for _ in range(1)]
    # gradients after the first micro-step (re-pack graph) + second (warm graph) before the update
    g_mid = {n: p.grad.clone() for n, p in tr_model.named_parameters() if p.requires_grad}
    ref1 = fresh()
    o1 = ref1(**batch, return_dict=True)
    (o1.loss / 2).backward()
    num = sum(float((g_mid[n] - p.grad).pow(2).sum()) for n, p in ref1.named_parameters() if p.requires_grad)
    den = sum(float(p.grad.pow(2).sum()) for n, p in ref1.named_parameters() if p.requires_grad)
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5  # same kernels, bf16 wgrad operands
    l_tr.append(float(tr.micro_step(batch)))  # second micro-step -> optimizer step
    assert abs(l_tr[0] - l_ref[0]) < 1e-3 and abs(l_tr[1] - l_ref[1]) < 1e-3, (l_tr, l_ref)
    from eilev_b200.train import no_weight_decay
    tp = [(n, p) for n, p in ref.named_parameters() if p.requires_grad]
    opt = torch.optim.AdamW([dict(params=[p for n, p in tp if not no_weight_decay(n)], weight_decay=0.05),
                             dict(params=[p for n, p in tp if no_weight_decay(n)], weight_decay=0.0)],
                            lr=1e-3, betas=(0.9, 0.999), eps=1e-8)  # HF Trainer's groups
    torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.requires_grad], 1.0)
    opt.step()
    pr = dict(ref.named_parameters())
    pt = dict(tr_model.named_parameters())
    worst = max(float((pt[n].detach() - pr[n].detach()).abs().max()) for n in names)
    assert worst < 2e-3, worst  # lr 1e-3: one Adam step moves every weight by ~1e-3
    # next micro-step runs the re-pack graph on the UPDATED parameters
    ref.zero_grad(set_to_none=True)
    l3_ref = ref_micro()
    l3_tr = float(tr.micro_step(batch))
    assert abs(l3_tr - l3_ref) < 5e-3, (l3_tr, l3_ref)


def _oracle_sequence_logprob(fx, cfg, row, seq):
    """fp32 oracle: sum of log p(token | prompt, previous tokens) of a generated continuation."""
    import torch.nn.functional as F
    from oracle import videoblip_ref as R
    sd, gi = fx["state_dict"], fx["gen_inputs"]
    feats = R.video_features(sd, cfg, gi["pixel_values"])[0]
    emb = R.splice(sd, gi["input_ids"], gi["video_input_mask"], feats)
    table = sd["language_model.model.decoder.embed_tokens.weight"].float()
    ids = torch.tensor(seq)
    e = torch.cat([emb[row:row + 1], table[ids][None]], 1)
    am = torch.cat([gi["attention_mask"][row:row + 1], torch.ones(1, len(seq), dtype=torch.long)], 1)
    h = R.opt_decoder(sd, cfg.text_config, e, am)
    n = emb.shape[1]
    lp = F.log_softmax(F.linear(h[0, n - 1:n - 1 + len(seq)], table), -1)
    return float(lp.gather(1, ids[:, None]).sum())


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt"])
def test_beam_search_and_repetition_penalty_match_reference_golden(name):
    """generate() with the kwargs the reference's samples use (num_beams, length_penalty,
    min_new_tokens, repetition_penalty; samples/*.py) — token ids vs the real reference's HF beam
    search on the fixtures (tests/golden/make_golden_beams.py).  Ids must be identical, except
    that a row may differ when the two hypotheses are a numerical tie: their fp32-oracle sequence
    log-probabilities within 0.1 nat (the bf16 logit error is ~0.05; the one such case on these
    fixtures, small_opt / beams3 / row 1, is 0.031 nat apart)."""
    import sys
    sys.path.insert(0, str(GOLDEN))
    import make_golden_beams as MB
    fx, cfg = load(name)
    gold = torch.load(GOLDEN / f"beams_{name}.pt", weights_only=False)
    m = build(cfg, fx["state_dict"])
    gi = cuda(fx["gen_inputs"])
    bad, ties = {}, {}
    for case, kw in MB.CASES.items():
        got = m.generate(**gi, **kw).cpu()
        want = gold[case]
        if got.shape != want.shape:
            bad[case] = (got.tolist(), want.tolist())
            continue
        for row in range(want.shape[0]):
            if torch.equal(got[row], want[row]):
                continue
            d = abs(_oracle_sequence_logprob(fx, cfg, row, got[row].tolist())
                    - _oracle_sequence_logprob(fx, cfg, row, want[row].tolist()))
            (ties if d < 0.1 else bad)[f"{case}/row{row}"] = (got[row].tolist(), want[row].tolist(), d)
    _dump(f"beams/{name}", mismatches=bad, ties=ties)
    assert not bad, bad
    assert len(ties) <= 1, ties


@pytest.mark.parametrize("kw", [dict(max_new_tokens=12), dict(max_new_tokens=16, min_new_tokens=16),
                                dict(max_new_tokens=24, min_new_tokens=3), dict(max_new_tokens=9, min_new_tokens=0)])
def test_greedy_generate_with_device_side_bookkeeping_equals_the_python_loop(kw):
    """generate()'s plain greedy search runs its token bookkeeping inside the decode graph
    (DecodeGraph.greedy_*); the ids must equal the Python loop's (device_bookkeeping=False) — including the EOS
    stop, the pad fill of finished rows and min_new_tokens — and a second call must reuse the captured graph."""
    fx, cfg = load("small_opt")
    m = build(cfg, fx["state_dict"])
    gi = cuda(fx["gen_inputs"])
    want = m.generate(**gi, device_bookkeeping=False, **kw)
    got = m.generate(**gi, **kw)
    assert torch.equal(got, want), (got.tolist(), want.tolist())
    again = m.generate(**gi, **kw)
    assert torch.equal(again, want)
    # an EOS that is certain to appear: the most frequent generated token
    eos = int(want.flatten().mode().values)
    want_e = m.generate(**gi, device_bookkeeping=False, eos_token_id=eos, **kw)
    got_e = m.generate(**gi, eos_token_id=eos, **kw)
    assert torch.equal(got_e, want_e), (got_e.tolist(), want_e.tolist())



def test_torch_ddp_wrap_returns_the_same_gradients(tmp_path):
    """HF Trainer (scripts/general/train_v2.py:169-190) wraps the model in torch DistributedDataParallel when
    launched on several GPUs: the Q-Former gradients leave our hand-written backward through ONE autograd
    Function, so DDP's per-parameter hooks must still fire (bucketed all-reduce, here over a 1-rank NCCL
    group) and hand back exactly the gradients of the unwrapped model."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    fx, cfg = load("small_opt")

    def frozen(m):
        for p in m.vision_model.parameters():
            p.requires_grad = False
        for p in m.language_model.parameters():
            p.requires_grad = False
        return m

    plain = frozen(build(cfg, fx["state_dict"])).eval()   # eval(): no dropout, so both runs see the same function
    plain(**cuda(fx["inputs"]), return_dict=True).loss.backward()
    want = {n: p.grad.clone() for n, p in plain.named_parameters() if p.grad is not None}
    assert want
    dist.init_process_group("nccl", init_method=f"file://{tmp_path}/rendezvous", rank=0, world_size=1)
    try:
        wrapped = DDP(frozen(build(cfg, fx["state_dict"])).eval(), device_ids=[0])
        out = wrapped(**cuda(fx["inputs"]), return_dict=True)
        out.loss.backward()
        got = {n: p.grad for n, p in wrapped.module.named_parameters() if p.grad is not None}
        assert set(got) == set(want)
        worst = max(max_abs(got[n], want[n]) for n in want)
        _dump("ddp_wrap/small_opt", n=len(got), worst_abs_diff=worst)
        for n in want:   # same kernels; only the mma.sync attention backward's fp32 atomics may reorder sums
            assert rel_l2(got[n], want[n]) < 1e-3 or max_abs(got[n], want[n]) < 1e-6, (n, rel_l2(got[n], want[n]))
    finally:
        dist.destroy_process_group()


def test_unfrozen_towers_are_reported_once():
    """Only the recipe's parameters are differentiated (train_v2.py:123-130): a model whose ViT / LM still have
    requires_grad=True gets a warning on its first grad-enabled forward instead of silently missing gradients."""
    import warnings
    fx, cfg = load("tiny_opt")
    m = build(cfg, fx["state_dict"]).train()
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        m(**cuda(fx["inputs"]), return_dict=True)
        m(**cuda(fx["inputs"]), return_dict=True)
    msgs = [str(w.message) for w in rec if "eilev_b200 computes gradients" in str(w.message)]
    assert len(msgs) == 1 and "vision_model / language_model parameters" in msgs[0], msgs
    from eilev_b200.train import freeze_for_recipe
    m2 = build(cfg, fx["state_dict"]).train()
    freeze_for_recipe(m2)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        m2(**cuda(fx["inputs"]), return_dict=True).loss.backward()
    assert not [w for w in rec if "eilev_b200 computes gradients" in str(w.message)]
