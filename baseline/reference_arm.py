"""The reference arm: yukw777/EILEV's own ``VideoBlipForConditionalGeneration`` (installed unmodified into
``baseline/_ref`` by ``baseline/install_ref.py``) driven through its public API — ``model(**batch)`` +
``loss.backward()`` — on the benchmark's synthetic datapoint.  None of this repo's kernels, engine or
oracle is on this path.

* ``device="cpu"``  — fp32 on the host cores: ``bench.py --impl reference`` (the CPU baseline arm).
* ``device="cuda"`` — bf16 weights, HuggingFace SDPA attention, on the B200: the ``library_bar`` the
  hand-written kernels have to beat.

Training-recipe state as in scripts/general/train_v2.py:116-130: vision tower and LM frozen, train mode,
input-require-grads.  The recipe's ``enable_input_require_grads()`` hooked only the LM input embeddings
on the pinned transformers 4.33.1; transformers 5.5 also hooks the vision embeddings, which makes the frozen
ViT build an autograd graph (> 60 GB at 17 clips), so the 4.33.1 behaviour is restored
with a forward hook on ``language_model.get_input_embeddings()``.
"""
from __future__ import annotations

import time

import torch

from . import install_ref


def build_reference_model(cfg, device: str, dtype: torch.dtype, seed: int = 1234):
    install_ref.add_to_path()
    from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel  # the real class
    from transformers.initialization import no_init_weights

    with no_init_weights(), torch.device(device):  # skip HF's ~1 min random init; seeded init below
        model = RefModel(cfg)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            ln = name.lower()
            if "layernorm" in ln or "layer_norm" in ln:
                p.fill_(1.0) if name.endswith("weight") else p.zero_()
            else:
                p.normal_(0.0, 0.02, generator=g)
        for name, b in model.named_buffers():
            if b.dtype.is_floating_point and not torch.isfinite(b).all():
                b.zero_()
    model.tie_weights()
    model = model.to(dtype)
    for p in model.vision_model.parameters():
        p.requires_grad = False
    for p in model.language_model.parameters():
        p.requires_grad = False
    emb = model.language_model.get_input_embeddings()
    emb.register_forward_hook(lambda m, i, o: o.requires_grad_(True))
    return model.train()


def step(model, batch) -> float:
    for p in model.parameters():
        p.grad = None
    out = model(**batch, return_dict=True)
    out.loss.backward()
    return float(out.loss.detach())


def run(cfg, batch, device: str, steps: int, warmup: int, threads: int | None = None):
    """Returns (seconds per fwd+bwd step, last loss, model-construction seconds)."""
    if device == "cpu":
        import os
        torch.set_num_threads(threads or os.cpu_count() or 1)
        dtype = torch.float32
    else:
        dtype = torch.bfloat16
    t0 = time.perf_counter()
    model = build_reference_model(cfg, device, dtype)
    build_s = time.perf_counter() - t0
    batch = {k: (v.to(device, dtype) if v.is_floating_point() else v.to(device)) for k, v in batch.items()}
    loss = float("nan")
    for _ in range(warmup):
        loss = step(model, batch)
    if device != "cpu":
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            loss = step(model, batch)
        e.record()
        torch.cuda.synchronize()
        dt = s.elapsed_time(e) * 1e-3 / max(steps, 1)
    else:
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = step(model, batch)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    del model
    return dt, loss, build_s
