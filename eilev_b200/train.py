"""Data-parallel training step for the VideoBLIP recipe (one process per GPU).

Mirrors what ``scripts/general/train_v2.py:104-219`` obtains from HuggingFace ``Trainer`` +
DDP for this model (README.md:139-164: bs 1 x grad-accum 16 x 8 GPUs, bf16, AdamW lr 1e-5,
weight decay 0.05, max_grad_norm 1.0):

    micro_step  = model(**batch).loss / accum ; backward  (accumulate, no communication)
    every `accum` micro-steps:
        ONE all-reduce (NCCL over NVLink / NVSwitch) of the flat gradient buffer of the 257
        trainable tensors (107 M f32 = 428 MB) -> global-norm clip -> fused AdamW kernel,
        all on the compute stream, no host synchronisation.

The datapoints shard across ranks with no activation traffic; the frozen ViT
and LM never communicate.  The trainable parameters and their gradients are re-pointed into
two flat f32 buffers so the collective and the optimizer are single launches.
"""
from __future__ import annotations

import math
from typing import Callable, Iterable

import torch
import torch.distributed as dist

from . import ops
from .engine.qformer import qformer_param_list


def freeze_for_recipe(model) -> None:
    """train_v2.py:123-130 — only Q-Former, query_tokens and language_projection train."""
    for p in model.vision_model.parameters():
        p.requires_grad = False
    for p in model.language_model.parameters():
        p.requires_grad = False
    model.enable_input_require_grads()


def no_weight_decay(name: str) -> bool:
    """Trainer.get_decay_parameter_names: parameters of LayerNorm modules and every ``bias`` are
    excluded from weight decay (transformers/trainer.py ``create_optimizer``)."""
    low = name.lower()
    return "bias" in low or "layernorm" in low or "layer_norm" in low


class FlatBuffers:
    """f32 master parameters / gradients of the trainable tensors in two contiguous buffers;
    ``param.data`` and ``param.grad`` become views, so autograd accumulates straight into the
    buffer the collective and the optimizer consume."""

    def __init__(self, named_params: Iterable[tuple[str, torch.nn.Parameter]]) -> None:
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("no trainable parameters")
        # HF Trainer (the reference's optimizer factory, train_v2.py:207-218 -> Trainer.create_optimizer)
        # applies weight decay to everything EXCEPT LayerNorm parameters and biases.  The decayed
        # tensors come first, so the fused AdamW runs as two launches over two contiguous segments.
        self.named = [x for x in named if not no_weight_decay(x[0])] + [x for x in named if no_weight_decay(x[0])]
        dev = self.named[0][1].device
        self.offsets = []
        total = 0
        for _, p in self.named:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4  # keep every view 16-byte aligned
        self.numel = total
        # first element of the no-decay segment (== numel when every tensor is decayed)
        self.decay_numel = next((off for (n, _), off in zip(self.named, self.offsets) if no_weight_decay(n)), total)
        self.params = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for (_, p), off in zip(self.named, self.offsets):
                view = self.params[off:off + p.numel()].view(p.shape)
                view.copy_(p.data.float())
                p.data = view
                p.grad = self.grads[off:off + p.numel()].view(p.shape)

    def zero_grad(self) -> None:
        self.grads.zero_()
        for (_, p), off in zip(self.named, self.offsets):  # re-attach (a caller may have set None)
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * off:
                p.grad = self.grads[off:off + p.numel()].view(p.shape)


class DataParallelTrainer:
    def __init__(self, model, *, lr: float = 1e-5, weight_decay: float = 0.05,
                 betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 1.0, grad_accum: int = 16, process_group=None,
                 lr_schedule: Callable[[int], float] | None = None,
                 update_fn: Callable | None = None) -> None:
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.max_grad_norm, self.grad_accum = max_grad_norm, grad_accum
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.lr_schedule = lr_schedule
        self.flat = FlatBuffers(qformer_param_list(model))
        model._grad_sink = True  # wgrad kernels accumulate straight into the flat f32 grad views
        self.exp_avg = torch.zeros_like(self.flat.params)
        self.exp_avg_sq = torch.zeros_like(self.flat.params)
        self.opt_step = 0
        self.micro = 0
        self._update = update_fn or self._fused_adamw
        self._pool_fresh = False
        dev = self.flat.params.device
        self._sumsq = torch.zeros((), dtype=torch.float32, device=dev)
        self._scale = torch.ones((), dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ steps
    def _fwd_bwd(self, batch: dict) -> torch.Tensor:
        out = self.model(**batch)
        loss = out.loss if hasattr(out, "loss") else out[0]
        (loss / self.grad_accum).backward()
        return loss.detach()

    def capture_graph(self, example_batch: dict) -> None:
        """Capture one micro-step (forward + backward, ~1 400 kernel launches) into a CUDA graph
        so a step costs one graph launch instead of ~1 400 Python/ctypes launches.  Batches fed
        to micro_step() afterwards must have the shapes of `example_batch` (the recipe's
        fixed-shape datapoints do; any other shape falls back to eager launches).  Call at an
        accumulation boundary: the warm-up steps' gradients are discarded."""
        assert self.micro % self.grad_accum == 0, "capture at an accumulation boundary"
        self._static = {k: v.clone() for k, v in example_batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: packs the frozen towers, sets kernel attributes
                self._fwd_bwd(self._static)
        torch.cuda.current_stream().wait_stream(side)
        self.flat.zero_grad()
        # Two graphs over one memory pool: the first micro-step after an optimizer step re-packs
        # the trainable Q-Former (f32 master -> bf16 operands + transposes, ~220 launches); the
        # other 15 of 16 replay a graph captured with the pack cache warm, which reads the very
        # buffers the first graph fills.
        self.model._pack.clear()
        from . import _lib
        before = _lib.launch_count()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._fwd_bwd(self._static)
        self.launches_per_graph_repack = _lib.launch_count() - before
        before = _lib.launch_count()
        self._graph_warm = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_warm, pool=self._graph.pool()):
            self._static_loss_warm = self._fwd_bwd(self._static)
        self.launches_per_graph = _lib.launch_count() - before
        self.flat.zero_grad()  # capture itself does not execute, but keep the contract explicit
        self._pool_fresh = False  # nothing has been replayed yet

    def micro_step(self, batch: dict) -> torch.Tensor:
        """One datapoint: forward + backward, gradients accumulate locally (DDP no_sync)."""
        graph = getattr(self, "_graph", None)
        if graph is not None and all(
                k in self._static and v.shape == self._static[k].shape and v.dtype == self._static[k].dtype
                for k, v in batch.items()) and len(batch) == len(self._static):
            for k, v in batch.items():
                if v.data_ptr() != self._static[k].data_ptr():
                    self._static[k].copy_(v, non_blocking=True)
            if not self._pool_fresh:
                # the packed bf16 Q-Former operands inside the graphs' pool are older than the
                # parameters (an optimizer step happened, possibly followed by eager micro-steps
                # of other shapes): replay the graph that re-packs them
                graph.replay()
                loss = self._static_loss
                self._pool_fresh = True
            else:
                self._graph_warm.replay()
                loss = self._static_loss_warm
        else:
            loss = self._fwd_bwd(batch)
        self.micro += 1
        if self.micro % self.grad_accum == 0:
            self.optimizer_step()
        return loss

    def grad_norm_and_scale(self) -> None:
        """scale = clip / world: gradients were summed over ranks (already divided by accum in
        micro_step); the clip coefficient uses the norm of the averaged gradient, as
        torch.nn.utils.clip_grad_norm_ does inside Trainer."""
        g = self.flat.grads
        self._sumsq.zero_()
        if g.is_cuda:
            ops.sumsq(g, self._sumsq)
        else:
            self._sumsq += (g.double() ** 2).sum().float()
        norm = torch.sqrt(self._sumsq) / self.world
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            clip = torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0)
        else:
            clip = torch.ones_like(norm)
        self._scale.copy_(clip / self.world)
        self.last_grad_norm = norm

    def optimizer_step(self) -> None:
        if self.world > 1:
            dist.all_reduce(self.flat.grads, op=dist.ReduceOp.SUM, group=self.group)
        self.grad_norm_and_scale()
        self.opt_step += 1
        lr = self.lr_schedule(self.opt_step) if self.lr_schedule is not None else self.lr
        cut = self.flat.decay_numel
        for lo, hi, wd in ((0, cut, self.weight_decay), (cut, self.flat.numel, 0.0)):
            if hi > lo:
                self._update(self.flat.params[lo:hi], self.flat.grads[lo:hi], self.exp_avg[lo:hi],
                             self.exp_avg_sq[lo:hi], lr=lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                             weight_decay=wd, step=self.opt_step, grad_scale=self._scale)
        # parameters changed in place through the flat buffer: drop the packed bf16 copies
        self.model._pack.clear()
        self._pool_fresh = False
        self.flat.zero_grad()

    @staticmethod
    def _fused_adamw(param, grad, exp_avg, exp_avg_sq, **kw) -> None:
        ops.adamw_(param, grad, exp_avg, exp_avg_sq, **kw)


def linear_schedule(base_lr: float, total_steps: int, warmup_steps: int = 0) -> Callable[[int], float]:
    """HF ``get_linear_schedule_with_warmup`` (Trainer default lr_scheduler_type='linear')."""

    def f(step: int) -> float:
        s = step - 1
        if s < warmup_steps:
            return base_lr * s / max(1, warmup_steps)
        return base_lr * max(0.0, (total_steps - s) / max(1, total_steps - warmup_steps))

    return f


def torch_adamw_reference(param, grad, exp_avg, exp_avg_sq, *, lr, beta1, beta2, eps, weight_decay,
                          step, grad_scale=None) -> None:
    """Plain-torch AdamW with the kernel's contract — used by the CPU (gloo) tests of the
    data-parallel host logic only."""
    g = grad * (grad_scale if grad_scale is not None else 1.0)
    param.mul_(1 - lr * weight_decay)
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = exp_avg_sq.sqrt() / math.sqrt(bc2) + eps
    param.addcdiv_(exp_avg, denom, value=-lr / bc1)


def gather_generated(generated_ids: torch.Tensor, pad_token_id: int, group=None) -> torch.Tensor:
    """Evaluation-time exchange of ``generate`` results (replicas only + an all-gather of
    the generated ids; scripts/general/generate_narration_texts.py:120-124 does it with accelerate's
    ``pad_across_processes(dim=1)`` + ``gather``).  Every rank holds (rows_r, len_r) token ids; they are
    right-padded with ``pad_token_id`` to the longest length over ranks, rows to the largest row
    count (the last batch of a sharded eval set may be short), gathered, and the filler rows dropped:
    the result is the same (sum rows_r, max len_r) tensor on every rank, rank order preserved.
    Integer bookkeeping over at most a few KB — NCCL's all_gather as is."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return generated_ids
    world = dist.get_world_size(group)
    dev = generated_ids.device
    shape = torch.tensor([generated_ids.shape[0], generated_ids.shape[1]], dtype=torch.long, device=dev)
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    rows = [int(s[0]) for s in shapes]
    max_rows, max_len = max(rows), max(int(s[1]) for s in shapes)
    padded = torch.full((max_rows, max_len), int(pad_token_id), dtype=generated_ids.dtype, device=dev)
    padded[:generated_ids.shape[0], :generated_ids.shape[1]] = generated_ids
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:r] for o, r in zip(out, rows)], dim=0)
