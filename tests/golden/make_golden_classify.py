"""Golden vectors for ``classify`` (eilev/model/v2.py:326-501), produced by the REAL reference.

Run in the authoring container only:   python tests/golden/make_golden_classify.py

The reference's ``_calc_class_log_likelihood`` iterates the LM's ``past_key_values`` as
legacy ``((k, v), ...)`` tuples (v2.py:457-460, transformers 4.33.1).  The installed
transformers 5.5.0 returns / expects ``Cache`` objects, so this script wraps the LM's
``forward`` with a shim that converts tuple <-> ``DynamicCache`` at the call boundary; the
reference's own code (v2.py) runs unmodified.  As a cross-check the same scores are also
recomputed without any cache from the reference's ``forward`` on [prompt ; class].

Writes tests/golden/classify_<name>.pt = inputs + scores (the weights are those of
tests/golden/<name>.pt).
"""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from transformers import Blip2Config  # noqa: E402
from transformers.cache_utils import DynamicCache  # noqa: E402


def install_legacy_cache_shim(lm):
    inner = lm.forward

    def forward(*args, past_key_values=None, **kw):
        if isinstance(past_key_values, tuple):
            past_key_values = DynamicCache(ddp_cache_data=[(k, v) for k, v in past_key_values])
        out = inner(*args, past_key_values=past_key_values, **kw)
        pkv = getattr(out, "past_key_values", None)
        if pkv is not None and not isinstance(pkv, tuple):
            out.past_key_values = tuple((layer.keys, layer.values) for layer in pkv.layers)
        return out

    lm.forward = forward


def main():
    sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
    from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel

    here = Path(__file__).resolve().parent
    for name in ("tiny_opt", "small_opt"):
        fx = torch.load(here / f"{name}.pt", weights_only=False)
        cfg = Blip2Config(**{k: fx["config"][k] for k in
                             ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
        model = RefModel(cfg).float().eval()
        model.load_state_dict(fx["state_dict"])
        model.tie_weights()
        install_legacy_cache_shim(model.language_model)
        g = torch.Generator().manual_seed(11)
        vocab = cfg.text_config.vocab_size
        prompt = fx["gen_inputs"]  # left-padded
        n_cls, lc = 5, 4
        class_ids = torch.randint(4, vocab - 2, (n_cls, lc), generator=g)
        class_mask = torch.ones(n_cls, lc, dtype=torch.long)
        class_mask[1, 2:] = 0
        class_mask[3, 1:] = 0
        class_mask[4, 3:] = 0
        class_ids[class_mask == 0] = 1  # pad id
        with torch.no_grad():
            scores = model.classify(prompt["input_ids"], class_ids, prompt["attention_mask"],
                                    prompt["pixel_values"], prompt["video_input_mask"], class_mask)
            chunked = model.classify(prompt["input_ids"], class_ids, prompt["attention_mask"],
                                     prompt["pixel_values"], prompt["video_input_mask"], class_mask,
                                     class_batch_size=2)
            # cache-free cross-check through the reference's forward
            b, lp = prompt["input_ids"].shape
            nv_per = prompt["video_input_mask"].sum(1) // cfg.num_query_tokens
            check = torch.zeros(b, n_cls)
            for ci in range(n_cls):
                ids = torch.cat([prompt["input_ids"], class_ids[ci][None].expand(b, -1)], 1)
                am = torch.cat([prompt["attention_mask"], class_mask[ci][None].expand(b, -1)], 1)
                vm = torch.cat([prompt["video_input_mask"], torch.zeros(b, lc, dtype=torch.long)], 1)
                out = model(ids, attention_mask=am, pixel_values=prompt["pixel_values"], video_input_mask=vm,
                            return_dict=True)
                logp = out.logits[:, lp - 1:lp + lc - 1].log_softmax(-1)
                tok = logp.gather(2, class_ids[ci][None, :, None].expand(b, -1, 1))[..., 0]
                check[:, ci] = (tok * class_mask[ci]).sum(1) / class_mask[ci].sum()
            del nv_per
        assert torch.allclose(scores, chunked, atol=1e-5), (scores, chunked)
        assert torch.allclose(scores, check, atol=1e-4), (scores, check)
        torch.save(dict(class_input_ids=class_ids, class_attention_mask=class_mask, scores=scores),
                   here / f"classify_{name}.pt")
        print(name, scores)


if __name__ == "__main__":
    main()
