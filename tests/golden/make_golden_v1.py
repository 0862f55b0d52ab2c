"""Golden fixtures for v1 (eilev/model/v1.py) from the REAL reference class on CPU fp32.
Run in the authoring container only:   python tests/golden/make_golden_v1.py

The reference's v1 inherits ``Blip2ForConditionalGeneration.forward / generate`` of its pinned
transformers 4.33.1, which PREPENDS the projected query rows to the embedded prompt.  The
installed transformers 5.5.0 scatters them into ``image_token_index`` placeholders instead, so
the real ``eilev.model.v1.VideoBlipForConditionalGeneration`` is run here behind a shim that
restores the 4.33.1 contract:
  * ``config.image_token_index`` is set and ``num_query_tokens`` placeholders are prepended to
    every prompt (with ones in the attention mask) — identical arithmetic to the 4.33.1 ``cat``;
  * ``VideoBlipVisionModel.forward`` (v1.py:17-92) is wrapped to swallow the
    ``interpolate_pos_encoding`` keyword 5.5.0 passes;
  * 5.5.0's ``generate`` returns prompt + new tokens; the new tokens are kept (4.33.1 returns only
    them for a decoder-only LM fed with embeddings).
The loss block (last ``labels.size(1)`` logits, shift, mean CE) is the same code in both versions.

Writes tests/golden/v1_{tiny_opt,small_opt,small_t5}.pt: inputs + reference outputs (incl. the fp32
top-2 logit margin of every greedy step, so a test can tell a numerical tie from an error) only — the
weights are the ``state_dict`` of the v2 fixture of the same name (v1 and v2 share all keys).
"""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from transformers import Blip2Config  # noqa: E402

HERE = Path(__file__).resolve().parent
KEYS = ("vision_config", "qformer_config", "text_config", "num_query_tokens")


def build_inputs(cfg, batch, time, text_lens, seed):
    g = torch.Generator().manual_seed(seed)
    img = cfg.vision_config.image_size
    vocab = cfg.text_config.vocab_size
    pad = cfg.text_config.pad_token_id
    pixel_values = torch.randn(batch, 3, time, img, img, generator=g)
    L = max(text_lens)
    ids = torch.full((batch, L), pad, dtype=torch.long)
    am = torch.zeros((batch, L), dtype=torch.long)
    for b, n in enumerate(text_lens):  # right padding, as the v1 collator does for training
        ids[b, :n] = torch.randint(4, vocab - 3, (n,), generator=g)
        am[b, :n] = 1
    return pixel_values, ids, am


def left_pad(ids, am, pad):
    out_ids, out_am = torch.full_like(ids, pad), torch.zeros_like(am)
    L = ids.shape[1]
    for b in range(ids.shape[0]):
        n = int(am[b].sum())
        out_ids[b, L - n:] = ids[b, :n]
        out_am[b, L - n:] = 1
    return out_ids, out_am


def main():
    sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
    from eilev.model import v1 as ref_v1

    orig = ref_v1.VideoBlipVisionModel.forward

    def vision_forward(self, pixel_values=None, output_attentions=None, output_hidden_states=None,
                       return_dict=None, interpolate_pos_encoding=False, **_):
        return orig(self, pixel_values, output_attentions, output_hidden_states, return_dict)

    ref_v1.VideoBlipVisionModel.forward = vision_forward

    for name, time, text_lens in (("tiny_opt", 2, [7, 4]), ("small_opt", 2, [11, 6, 9]), ("small_t5", 2, [37, 22])):
        base = torch.load(HERE / f"{name}.pt", weights_only=False)
        cfg = Blip2Config(**{k: base["config"][k] for k in KEYS})
        tcfg = cfg.text_config
        decoder_only = cfg.use_decoder_only_language_model
        if decoder_only:
            tcfg.dropout = 0.0
            tcfg.attention_dropout = 0.0
        else:
            tcfg.dropout_rate = 0.0
        cfg.qformer_config.hidden_dropout_prob = 0.0
        cfg.qformer_config.attention_probs_dropout_prob = 0.0
        img_tok = tcfg.vocab_size - 2
        cfg.image_token_index = img_tok
        model = ref_v1.VideoBlipForConditionalGeneration(cfg).float().eval()
        missing = model.load_state_dict(base["state_dict"], strict=False)
        assert not missing.unexpected_keys and all("lm_head" in k for k in missing.missing_keys), missing
        model.tie_weights()
        for p in model.vision_model.parameters():
            p.requires_grad = False
        for p in model.language_model.parameters():
            p.requires_grad = False

        nq = cfg.num_query_tokens
        batch = len(text_lens)
        pixel_values, ids, am = build_inputs(cfg, batch, time, text_lens, seed=11)

        def shim(i, a):
            return (torch.cat([torch.full((i.shape[0], nq), img_tok, dtype=torch.long), i], 1),
                    torch.cat([torch.ones(i.shape[0], nq, dtype=torch.long), a], 1))

        if decoder_only:  # labels aligned with the prompt, -100 on its first half and on padding
            labels = ids.clone()
            labels[am == 0] = -100
            for b, n in enumerate(text_lens):
                labels[b, : n // 2] = -100
        else:
            g = torch.Generator().manual_seed(3)
            labels = torch.full((batch, 7), -100, dtype=torch.long)
            for b in range(batch):
                n = 7 - 2 * b
                labels[b, :n] = torch.randint(4, tcfg.vocab_size - 3, (n,), generator=g)

        emb = model.language_model.get_input_embeddings()
        hook = emb.register_forward_hook(lambda m, i, o: o.requires_grad_(True))
        sids, sam = shim(ids, am)
        out = model(pixel_values=pixel_values, input_ids=sids, attention_mask=sam, labels=labels, return_dict=True)
        out.loss.backward()
        hook.remove()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        with torch.no_grad():
            nolab = model(pixel_values=pixel_values, input_ids=sids, attention_mask=sam,
                          **({} if decoder_only else {"decoder_input_ids": torch.zeros(batch, 1, dtype=torch.long)}),
                          return_dict=True)
            gids, gam = left_pad(ids, am, tcfg.pad_token_id) if decoder_only else (ids, am)
            s_gids, s_gam = shim(gids, gam)
            kw = dict(max_new_tokens=5, min_new_tokens=5, do_sample=False, num_beams=1, output_scores=True,
                      return_dict_in_generate=True)

            def margins(scores):  # (steps, rows) fp32 top-1 minus top-2 logit of every greedy step
                top = torch.stack([s.float().topk(2, dim=-1).values for s in scores])
                return top[..., 0] - top[..., 1]

            res = model.generate(pixel_values=pixel_values, input_ids=s_gids, attention_mask=s_gam, **kw)
            gen, gen_margins = res.sequences, margins(res.scores)
            gen_noprompt = noprompt_margins = None
            if decoder_only:
                gen = gen[:, s_gids.shape[1]:]
                res = model.generate(pixel_values=pixel_values, **kw)
                full, noprompt_margins = res.sequences, margins(res.scores)
                gen_noprompt = full[:, nq + 1:]  # 5.5.0 returns [placeholders, bos] + new
                assert gen.shape == (batch, 5) and gen_noprompt.shape == (batch, 5), (gen.shape, full.shape)
        fixture = dict(
            base=name, image_token_index=img_tok,
            inputs=dict(pixel_values=pixel_values, input_ids=ids, attention_mask=am, labels=labels),
            gen_inputs=dict(pixel_values=pixel_values, input_ids=gids, attention_mask=gam),
            loss=out.loss.detach(), logits=out.logits.detach(), logits_no_labels=nolab.logits.detach(),
            query_output=out.qformer_outputs.last_hidden_state.detach(),
            grads=grads, generated=gen, generated_no_prompt=gen_noprompt,
            generated_margins=gen_margins, generated_no_prompt_margins=noprompt_margins,
        )
        path = HERE / f"v1_{name}.pt"
        torch.save(fixture, path)
        print(name, "loss", float(out.loss), "logits", tuple(out.logits.shape), "no-labels", tuple(nolab.logits.shape),
              "grads", len(grads), "gen", gen.tolist(), "no prompt", None if gen_noprompt is None else gen_noprompt.tolist(),
              "min margin", float(gen_margins.min()), None if noprompt_margins is None else float(noprompt_margins.min()),
              "bytes", path.stat().st_size)


if __name__ == "__main__":
    main()
