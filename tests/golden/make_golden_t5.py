"""Golden fixture for the flan-T5 (seq2seq) branch of VideoBLIP (eilev/model/v2.py:228-238),
produced by the REAL reference on CPU fp32.  Run in the authoring container only:
    python tests/golden/make_golden_t5.py
Writes tests/golden/small_t5.pt: config, seeded state_dict, inputs, loss, logits, encoder
output and the gradients of the trainable tensors (Q-Former side), as make_golden.py does for OPT.
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

CONFIG = dict(
    vision_config=dict(hidden_size=64, intermediate_size=128, projection_dim=32, num_hidden_layers=2,
                       num_attention_heads=4, patch_size=14, image_size=56),
    qformer_config=dict(hidden_size=48, num_hidden_layers=2, num_attention_heads=3, intermediate_size=96,
                        encoder_hidden_size=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
    text_config=dict(model_type="t5", d_model=128, d_kv=64, d_ff=256, num_layers=3, num_decoder_layers=2,
                     num_heads=3, vocab_size=264, feed_forward_proj="gated-gelu", tie_word_embeddings=False,
                     decoder_start_token_id=0, pad_token_id=0, eos_token_id=1, dropout_rate=0.0,
                     relative_attention_num_buckets=32, relative_attention_max_distance=128),
    num_query_tokens=8)


# The reference's OWN T5 test configuration (tests/model/test_model_v2.py:122-146): T5Config defaults,
# i.e. the original T5 — non-gated ReLU feed-forward (T5DenseActDense), head tied to the embedding,
# decoder output scaled by d_model**-0.5 — with a small image and vocabulary to keep the fixture small.
CONFIG_RELU = dict(
    vision_config=dict(hidden_size=8, intermediate_size=16, projection_dim=4, num_hidden_layers=2,
                       num_attention_heads=4, patch_size=8, image_size=32),
    qformer_config=dict(hidden_size=8, num_hidden_layers=2, num_attention_heads=2, intermediate_size=16,
                        encoder_hidden_size=8, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
    text_config=dict(model_type="t5", d_model=8, d_kv=4, d_ff=16, num_layers=2, num_heads=2,
                     decoder_start_token_id=0, vocab_size=264, dropout_rate=0.0),
    num_query_tokens=4)


def main(name="small_t5", config=CONFIG, init_seed=4321, init_std=0.06, img=56, text_len=40):
    sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
    from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel

    torch.manual_seed(0)
    cfg = Blip2Config(**config)
    model = RefModel(cfg).float().eval()
    sd = model.state_dict()
    sane_init_(sd, seed=init_seed, std=init_std)
    # T5 attention has no 1/sqrt(d) factor: trained checkpoints carry it in small q / k weights.
    # Shrink the random q / k projections accordingly so the scores are O(0.3) and a bf16 run is a
    # meaningful parity target (with O(4) scores the reference's own bf16 logits are 13 % off).
    for k in sd:
        if k.startswith("language_model.") and k.endswith((".q.weight", ".k.weight")):
            sd[k] = sd[k] * 0.25
    model.load_state_dict(sd)
    model.tie_weights()
    for p in model.vision_model.parameters():
        p.requires_grad = False
    for p in model.language_model.parameters():
        p.requires_grad = False
    g = torch.Generator().manual_seed(5)
    nq, vocab = cfg.num_query_tokens, cfg.text_config.vocab_size
    nv, t, batch = 3, 2, 2
    pixel_values = torch.randn(nv, 3, t, img, img, generator=g)
    clips_per = [2, 1]
    rows = []
    for b in range(batch):  # seq2seq layout (data/utils.py:200-217): no bos, eos after the prompt
        ids, vm = [], []
        for _ in range(clips_per[b]):
            ids += [0] * nq + [3]
            vm += [1] * nq + [0]
            txt = torch.randint(4, vocab - 2, (text_len + 9 * b,), generator=g).tolist()
            ids += txt
            vm += [0] * len(txt)
        ids += [1]
        vm += [0]
        rows.append((ids, vm))
    L = (max(len(r[0]) for r in rows) + 7) // 8 * 8
    input_ids = torch.zeros((batch, L), dtype=torch.long)
    attn = torch.zeros((batch, L), dtype=torch.long)
    vmask = torch.zeros((batch, L), dtype=torch.long)
    for b, (ids, vm) in enumerate(rows):
        n = len(ids)
        input_ids[b, :n] = torch.tensor(ids)
        attn[b, :n] = 1
        vmask[b, :n] = torch.tensor(vm)
    labels = torch.full((batch, 9), -100, dtype=torch.long)
    labels[0, :9] = torch.randint(4, vocab - 2, (9,), generator=g)
    labels[1, :6] = torch.randint(4, vocab - 2, (6,), generator=g)
    inputs = dict(pixel_values=pixel_values, input_ids=input_ids, attention_mask=attn,
                  video_input_mask=vmask, labels=labels)
    emb = model.language_model.get_input_embeddings()
    hook = emb.register_forward_hook(lambda m, i, o: o.requires_grad_(True))
    out = model(**inputs, return_dict=True)
    out.loss.backward()
    hook.remove()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    gen_inputs = {k: v for k, v in inputs.items() if k != "labels"}
    with torch.no_grad():
        gen = model.generate(**gen_inputs, max_new_tokens=6, do_sample=False, num_beams=1)
    fixture = dict(generated=gen, config=cfg.to_dict(), state_dict={k: v.clone() for k, v in model.state_dict().items()},
                   inputs=inputs, loss=out.loss.detach(), logits=out.logits.detach(),
                   encoder_last_hidden_state=out.language_model_outputs.encoder_last_hidden_state.detach(),
                   query_output=out.qformer_outputs.last_hidden_state.detach(), grads=grads)
    path = Path(__file__).resolve().parent / f"{name}.pt"
    torch.save(fixture, path)
    print("generated", gen.tolist())
    print("loss", float(out.loss), "L", tuple(input_ids.shape), "grads", len(grads), "bytes", path.stat().st_size)


if __name__ == "__main__":
    which = sys.argv[1:] or ["small_t5", "tiny_t5_relu"]
    if "small_t5" in which:
        main()
    if "tiny_t5_relu" in which:
        main("tiny_t5_relu", CONFIG_RELU, init_seed=99, init_std=0.5, img=32, text_len=10)
