"""Yardstick for the v1 fixtures: the REAL reference class (eilev/model/v1.py behind the shim of
make_golden_v1.py) in bf16 vs its own fp32 golden outputs, on CPU — how far apart are logits and
gradients from precision alone?  Also prints the fp32 top-2 logit margin of every greedy step of the
no-prompt generate golden (a margin below the bf16 logit error is a numerical tie).
Run in the authoring container:  python tests/golden/bf16_yardstick_v1.py"""
import sys, types, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/reference'); sys.path.insert(0, '/root/repo/tests/golden')
sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
from transformers import Blip2Config
from eilev.model import v1 as ref_v1
from oracle import videoblip_ref as R

orig = ref_v1.VideoBlipVisionModel.forward
ref_v1.VideoBlipVisionModel.forward = lambda self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None, interpolate_pos_encoding=False, **_: orig(self, pixel_values, output_attentions, output_hidden_states, return_dict)

for name in ['tiny_opt', 'small_opt', 'small_t5']:
    base = torch.load(f'/root/repo/tests/golden/{name}.pt', weights_only=False)
    fx = torch.load(f'/root/repo/tests/golden/v1_{name}.pt', weights_only=False)
    cfg = Blip2Config(**{k: base["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    cfg.image_token_index = fx["image_token_index"]
    nq = cfg.num_query_tokens
    for mode in ['bf16_params', 'autocast']:
        model = ref_v1.VideoBlipForConditionalGeneration(cfg).float().eval()
        model.load_state_dict(base['state_dict'], strict=False); model.tie_weights()
        for p in model.vision_model.parameters(): p.requires_grad = False
        for p in model.language_model.parameters(): p.requires_grad = False
        if mode == 'bf16_params':
            model = model.to(torch.bfloat16)
        emb = model.language_model.get_input_embeddings()
        emb.register_forward_hook(lambda m, i, o: o.requires_grad_(True))
        i = fx['inputs']
        b = i['input_ids'].shape[0]
        inp = dict(pixel_values=i['pixel_values'], labels=i['labels'],
                   input_ids=torch.cat([torch.full((b, nq), cfg.image_token_index), i['input_ids']], 1),
                   attention_mask=torch.cat([torch.ones(b, nq, dtype=torch.long), i['attention_mask']], 1))
        if mode == 'bf16_params': inp['pixel_values'] = inp['pixel_values'].to(torch.bfloat16)
        try:
            if mode == 'autocast':
                with torch.autocast('cpu', dtype=torch.bfloat16):
                    out = model(**inp, return_dict=True)
            else:
                out = model(**inp, return_dict=True)
            out.loss.backward()
        except Exception as e:
            print(name, mode, 'failed', repr(e)[:200]); continue
        num = den = 0
        for n, p in model.named_parameters():
            if p.grad is not None:
                r = fx['grads'][n]; num += float((p.grad.float() - r).pow(2).sum()); den += float(r.pow(2).sum())
        lg = float((out.logits.float() - fx['logits']).norm() / fx['logits'].norm())
        print(name, mode, 'loss', float(out.loss), 'ref', float(fx['loss']), 'logits rel', round(lg, 4),
              'grad global rel-L2', round((num / den) ** 0.5, 4))
    if cfg.use_decoder_only_language_model:
        sd = base['state_dict']
        pv = fx['gen_inputs']['pixel_values']
        tcfg = cfg.text_config
        feats = R.video_features(sd, cfg, pv)[0]
        table = sd["language_model.model.decoder.embed_tokens.weight"].float()
        bsz = pv.shape[0]
        for label, ids, am, want in (("prompt", fx['gen_inputs']['input_ids'], fx['gen_inputs']['attention_mask'], fx['generated']),
                                     ("no prompt", torch.full((bsz, 1), tcfg.bos_token_id), torch.ones(bsz, 1, dtype=torch.long), fx['generated_no_prompt'])):
            emb = torch.cat([feats.view(bsz, nq, -1), table[ids]], 1)
            mask = torch.cat([torch.ones(bsz, nq, dtype=torch.long), am], 1)
            margins = []
            for t in range(want.shape[1]):
                lg = torch.nn.functional.linear(R.opt_decoder(sd, tcfg, emb, mask)[:, -1], table)
                top = lg.topk(2, -1).values
                margins.append([round(float(x), 4) for x in (top[:, 0] - top[:, 1])])
                emb = torch.cat([emb, table[want[:, t]][:, None]], 1)
                mask = torch.cat([mask, torch.ones(bsz, 1, dtype=torch.long)], 1)
            print(name, label, 'fp32 top-2 margins per step (rows):', margins)
