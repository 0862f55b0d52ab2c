"""Generates the golden fixtures under tests/golden/ by running the REAL reference
(/root/reference/eilev/model/v2.py on the installed HuggingFace transformers) on CPU fp32.

Run in the authoring container only (the reference checkout does not travel to the GPU
box):   python tests/golden/make_golden.py

Each fixture ``<name>.pt`` holds: the Blip2Config dict, the seeded state_dict, the inputs,
and the reference outputs (loss, logits, intermediate activations, gradients of the
trainable tensors, greedy tokens).  tests/test_oracle.py pins oracle/videoblip_ref.py
against these; tests/test_model_gpu.py compares the CUDA path against the oracle and
against these fixtures.
"""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from transformers import Blip2Config  # noqa: E402

from oracle.videoblip_ref import sane_init_  # noqa: E402

CONFIGS = {
    # the reference's own tiny test config (tests/model/test_model_v2.py:95-121): head_dim 2/4
    "tiny_opt": dict(
        config=dict(
            vision_config=dict(hidden_size=8, intermediate_size=16, projection_dim=4, num_hidden_layers=2,
                               num_attention_heads=4, patch_size=8, image_size=32),
            qformer_config=dict(hidden_size=8, num_hidden_layers=2, num_attention_heads=2,
                                intermediate_size=16, encoder_hidden_size=8),
            text_config=dict(model_type="opt", hidden_size=8, num_hidden_layers=2, ffn_dim=16,
                             num_attention_heads=2, vocab_size=99, max_position_embeddings=64,
                             word_embed_proj_dim=8),
            num_query_tokens=4),
        std=0.5, num_videos=2, time=2, batch=1, text=5, pad_to=0),
    # aligned small config: exercises the vectorised / tcgen05-eligible code paths
    "small_opt": dict(
        config=dict(
            vision_config=dict(hidden_size=64, intermediate_size=128, projection_dim=32, num_hidden_layers=3,
                               num_attention_heads=4, patch_size=14, image_size=56),
            qformer_config=dict(hidden_size=48, num_hidden_layers=4, num_attention_heads=3,
                                intermediate_size=96, encoder_hidden_size=64),
            text_config=dict(model_type="opt", hidden_size=80, num_hidden_layers=3, ffn_dim=160,
                             num_attention_heads=5, vocab_size=264, max_position_embeddings=128,
                             word_embed_proj_dim=80),
            num_query_tokens=8),
        std=0.15, num_videos=3, time=2, batch=2, text=6, pad_to=8),
}


def build_inputs(cfg, spec, seed=7):
    g = torch.Generator().manual_seed(seed)
    nq = cfg.num_query_tokens
    vocab = cfg.text_config.vocab_size
    img = cfg.vision_config.image_size
    nv, t, batch = spec["num_videos"], spec["time"], spec["batch"]
    pixel_values = torch.randn(nv, 3, t, img, img, generator=g)
    # distribute clips over the batch rows: row b gets clips_per[b] clips
    clips_per = [nv // batch + (1 if b < nv % batch else 0) for b in range(batch)]
    rows = []
    for b in range(batch):
        ids, vm, lab = [2], [0], [-100]
        for _ in range(clips_per[b]):
            ids += [1] * nq + [vocab - 1]
            vm += [1] * nq + [0]
            lab += [-100] * (nq + 1)
            txt = torch.randint(4, vocab - 2, (spec["text"],), generator=g).tolist()
            ids += txt
            vm += [0] * len(txt)
            lab += [-100] * len(txt)
        tgt = torch.randint(4, vocab - 2, (3 + b,), generator=g).tolist()
        ids += tgt
        vm += [0] * len(tgt)
        lab += tgt
        rows.append((ids, vm, lab))
    L = max(len(r[0]) for r in rows)
    if spec["pad_to"]:
        L = (L + spec["pad_to"] - 1) // spec["pad_to"] * spec["pad_to"]
    input_ids = torch.full((batch, L), 1, dtype=torch.long)
    attn = torch.zeros((batch, L), dtype=torch.long)
    vmask = torch.zeros((batch, L), dtype=torch.long)
    labels = torch.full((batch, L), -100, dtype=torch.long)
    for b, (ids, vm, lab) in enumerate(rows):  # right padding, as in training
        n = len(ids)
        input_ids[b, :n] = torch.tensor(ids)
        attn[b, :n] = 1
        vmask[b, :n] = torch.tensor(vm)
        labels[b, :n] = torch.tensor(lab)
    return dict(pixel_values=pixel_values, input_ids=input_ids, attention_mask=attn,
                video_input_mask=vmask, labels=labels)


def left_pad(inputs):
    """Left-padded variant of the prompt (generation layout, generate_narration_texts.py:229-230)."""
    out = {k: v.clone() for k, v in inputs.items()}
    attn = inputs["attention_mask"]
    L = attn.shape[1]
    for b in range(attn.shape[0]):
        n = int(attn[b].sum())
        for key, fill in (("input_ids", 1), ("attention_mask", 0), ("video_input_mask", 0)):
            row = torch.full((L,), fill, dtype=torch.long)
            row[L - n:] = inputs[key][b, :n]
            out[key][b] = row
    out.pop("labels")
    return out


def main():
    # pytorchvideo is absent here; eilev/__init__ does not need it for eilev.model.v2
    sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
    from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel

    for name, spec in CONFIGS.items():
        torch.manual_seed(0)
        cfg = Blip2Config(**spec["config"])
        cfg.text_config.dropout = 0.0
        cfg.text_config.attention_dropout = 0.0
        cfg.qformer_config.hidden_dropout_prob = 0.0
        cfg.qformer_config.attention_probs_dropout_prob = 0.0
        model = RefModel(cfg).float().eval()
        sd = model.state_dict()
        sane_init_(sd, seed=1234, std=spec["std"])
        model.load_state_dict(sd)
        model.tie_weights()
        for p in model.vision_model.parameters():
            p.requires_grad = False
        for p in model.language_model.parameters():
            p.requires_grad = False
        inputs = build_inputs(cfg, spec)
        # pinned-4.33.1 semantics: only the LM input embeddings get the require-grad hook
        emb = model.language_model.get_input_embeddings()
        hook = emb.register_forward_hook(lambda m, i, o: o.requires_grad_(True))
        out = model(**inputs, return_dict=True)
        out.loss.backward()
        hook.remove()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        with torch.no_grad():
            vis = model.vision_model(inputs["pixel_values"], return_dict=True)
            lp = left_pad(inputs)
            gen = model.generate(**lp, max_new_tokens=6, min_new_tokens=6, do_sample=False, num_beams=1)
            text_only = model(inputs["input_ids"], attention_mask=inputs["attention_mask"],
                              labels=inputs["labels"], return_dict=True)
        fixture = dict(
            config=cfg.to_dict(), std=spec["std"],
            state_dict={k: v.clone() for k, v in model.state_dict().items()},
            inputs=inputs, gen_inputs=lp,
            loss=out.loss.detach(), logits=out.logits.detach(),
            image_embeds=vis.last_hidden_state.detach(), pooler_output=vis.pooler_output.detach(),
            query_output=out.qformer_outputs.last_hidden_state.detach(),
            grads=grads, generated=gen, text_only_loss=text_only.loss.detach(),
            text_only_logits=text_only.logits.detach(),
        )
        path = Path(__file__).resolve().parent / f"{name}.pt"
        torch.save(fixture, path)
        print(name, "loss", float(out.loss), "L", inputs["input_ids"].shape, "grads", len(grads),
              "gen", gen.tolist(), "bytes", path.stat().st_size)


if __name__ == "__main__":
    main()
