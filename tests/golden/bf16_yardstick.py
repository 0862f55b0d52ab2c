"""Yardstick: the REAL reference (eilev/model/v2.py on HF) in bf16 vs its own fp32 golden
outputs, on CPU: how far apart are logits and gradients from precision alone?  Run in the
authoring container: python tests/golden/bf16_yardstick.py
Measured: tiny_opt grads 3.8-4.3 %, small_opt grads 8.9-9.6 % global rel-L2; logits 0.8-1.5 %."""
import sys, types, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/reference'); sys.path.insert(0,'/root/repo/tests/golden')
sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
from transformers import Blip2Config
from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel
for name in ['tiny_opt','small_opt']:
    fx = torch.load(f'/root/repo/tests/golden/{name}.pt', weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config","qformer_config","text_config","num_query_tokens")})
    for mode in ['bf16_params','autocast']:
        model = RefModel(cfg).float().eval()
        model.load_state_dict(fx['state_dict']); model.tie_weights()
        for p in model.vision_model.parameters(): p.requires_grad=False
        for p in model.language_model.parameters(): p.requires_grad=False
        if mode=='bf16_params':
            model = model.to(torch.bfloat16)
        emb = model.language_model.get_input_embeddings()
        h = emb.register_forward_hook(lambda m,i,o: o.requires_grad_(True))
        inp = dict(fx['inputs'])
        if mode=='bf16_params': inp['pixel_values']=inp['pixel_values'].to(torch.bfloat16)
        try:
            if mode=='autocast':
                with torch.autocast('cpu', dtype=torch.bfloat16):
                    out = model(**inp, return_dict=True)
            else:
                out = model(**inp, return_dict=True)
            out.loss.backward()
        except Exception as e:
            print(name, mode, 'failed', repr(e)[:200]); continue
        num=den=0
        for n,p in model.named_parameters():
            if p.grad is not None:
                r=fx['grads'][n]; num+=float((p.grad.float()-r).pow(2).sum()); den+=float(r.pow(2).sum())
        lg = float((out.logits.float()-fx['logits']).norm()/fx['logits'].norm())
        print(name, mode, 'loss', float(out.loss), 'ref', float(fx['loss']), 'logits rel', round(lg,4), 'grad global rel-L2', round((num/den)**0.5,4))
