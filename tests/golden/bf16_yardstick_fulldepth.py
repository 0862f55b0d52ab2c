"""Yardstick at FULL DEPTH: the REAL reference class (baseline/_ref, eilev/model/v2.py on HF transformers)
in bf16 vs itself in fp32 on CPU, on exactly the inputs and seeded weights of tests/test_zc_fulldepth_gpu.py
(39 ViT / 12 Q-Former / 32 OPT or 24+24 flan-t5-xl layers, 2 clips x 8 frames, L = 120).  How far apart are
logits and the 257 gradients from precision alone?  That gap is what the CUDA path's tolerance is stated
against.  Run in the authoring container:

    python tests/golden/bf16_yardstick_fulldepth.py [opt|t5]      # ~10 min on 8 vCPUs

Writes profiles/r02_bf16_yardstick_fulldepth.json.
"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from baseline import install_ref  # noqa: E402

install_ref.add_to_path()
from transformers import Blip2Config  # noqa: E402
from transformers.initialization import no_init_weights  # noqa: E402

import test_zc_fulldepth_gpu as T  # noqa: E402
from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel  # noqa: E402


def build(cfg, sd):
    with no_init_weights():
        m = RefModel(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("lm_head" in k or "embed_tokens" in k for k in missing), (missing, unexpected)
    m.tie_weights()
    for p in m.vision_model.parameters():
        p.requires_grad = False
    for p in m.language_model.parameters():
        p.requires_grad = False
    m.language_model.get_input_embeddings().register_forward_hook(lambda mod, i, o: o.requires_grad_(True))
    return m.eval()  # dropout off: precision only


def run(m, inputs, mode):
    for p in m.parameters():
        p.grad = None
    t0 = time.time()
    if mode == "autocast":  # what HF Trainer --bf16 does (fp32 master weights)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out = m(**inputs, return_dict=True)
    else:
        out = m(**inputs, return_dict=True)
    out.loss.backward()
    grads = {n: p.grad.float().clone() for n, p in m.named_parameters() if p.grad is not None}
    print(mode, "loss", float(out.loss), f"{time.time() - t0:.0f}s", flush=True)
    return out.logits.detach().float(), float(out.loss), grads


def gap(a, b):
    return float((a - b).norm() / b.norm())


def main(which):
    res = {}
    for lm in which:
        text = T.OPT if lm == "opt" else T.T5
        cfg = Blip2Config(vision_config=T.VISION, qformer_config=T.QFORMER, text_config=text, num_query_tokens=32)
        with torch.device("meta"):
            from eilev_b200.model.v2 import VideoBlipForConditionalGeneration as Ours
            skeleton = Ours(cfg)
        sd = T._seeded_state_dict(skeleton)
        if lm == "opt":
            sd["language_model.lm_head.weight"] = sd["language_model.model.decoder.embed_tokens.weight"]
            inputs = T._opt_inputs()
        else:
            for k in sd:
                if k.startswith("language_model.") and k.endswith((".q.weight", ".k.weight")):
                    sd[k] = sd[k] * 0.25
            sd["language_model.encoder.embed_tokens.weight"] = sd["language_model.shared.weight"]
            sd["language_model.decoder.embed_tokens.weight"] = sd["language_model.shared.weight"]
            inputs = T._t5_inputs()
        m = build(cfg, sd)
        valid = inputs["attention_mask"][0].bool() if lm == "opt" else slice(None)
        l32, loss32, g32 = run(m, inputs, "fp32")
        out = {}
        for mode in ("autocast", "bf16_params"):
            if mode == "bf16_params":
                m = m.to(torch.bfloat16)
                inputs = dict(inputs, pixel_values=inputs["pixel_values"].to(torch.bfloat16))
            lg, loss, g = run(m, inputs, mode)
            num = sum(float((g[n] - g32[n]).pow(2).sum()) for n in g32)
            den = sum(float(g32[n].pow(2).sum()) for n in g32)
            a, b = lg[0, valid], l32[0, valid]
            out[mode] = dict(logits_rel_l2=gap(a, b), logits_max_abs=float((a - b).abs().max()),
                             loss=loss, loss_fp32=loss32, grad_rel_l2=(num / den) ** 0.5, grads_compared=len(g32))
            print(lm, mode, out[mode], flush=True)
        res[lm] = out
        del m
    dst = ROOT / "profiles" / "r02_bf16_yardstick_fulldepth.json"
    prev = json.loads(dst.read_text()) if dst.exists() else {}
    prev.update(res)
    dst.write_text(json.dumps(prev, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:] or ["opt", "t5"])
