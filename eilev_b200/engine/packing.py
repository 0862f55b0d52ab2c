"""Device-resident packed weights.

The module tree keeps the reference's (HuggingFace BLIP-2) parameter names and dtypes so
checkpoints load unchanged; the kernels want bf16 GEMM operands (fused / concatenated /
pre-transposed where that saves a launch) and fp32 biases + LayerNorm parameters.  Packed
copies are cached per (storage, version) so frozen towers are packed exactly once and the
trainable Q-Former is re-packed only after an optimizer step touched it.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from .. import ops


def _sig(params: Sequence[torch.Tensor]):
    return tuple((p.data_ptr(), p._version, p.dtype, p.device, tuple(p.shape)) for p in params)


class PackCache:
    def __init__(self) -> None:
        self._store: dict[str, tuple[tuple, object]] = {}

    def get(self, key: str, params: Sequence[torch.Tensor], build: Callable[[], object]):
        sig = _sig(params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._store[key] = (sig, val)
        return val

    def clear(self) -> None:
        self._store.clear()


def bf16(t: torch.Tensor) -> torch.Tensor:
    """bf16 contiguous device copy (no copy when already so)."""
    t = t.detach()
    if t.dtype == torch.bfloat16:
        return t.contiguous()
    return ops.convert(t, torch.bfloat16)


def f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype == torch.float32:
        return t.contiguous()
    return ops.convert(t, torch.float32)


def cat_bf16(ts: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([bf16(t) for t in ts], dim=0).contiguous()


def cat_f32(ts: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([f32(t) for t in ts], dim=0).contiguous()


def pad_k(w: torch.Tensor, kpad: int) -> torch.Tensor:
    """(N, K) -> (N, kpad) bf16 zero padded along K."""
    w = bf16(w)
    n, k = w.shape
    if k == kpad:
        return w
    out = torch.zeros((n, kpad), dtype=torch.bfloat16, device=w.device)
    out[:, :k] = w
    return out


def transposed(w_bf16: torch.Tensor) -> torch.Tensor:
    """(N, K) bf16 -> (K, N) bf16 with a 16-byte aligned row stride (dgrad operand)."""
    return ops.transpose(w_bf16)


def ln_fold(w: torch.Tensor, b: torch.Tensor | None, gamma: torch.Tensor, beta: torch.Tensor):
    """Weights of ``LayerNorm(x) @ w.T + b`` for the GEMM that consumes the UN-normalised x
    (vb_gemm_args.ln_stats): LN(x) w^T + b = rstd (x (w*gamma)^T) - rstd mean colsum(w*gamma) + (b + w beta).
    Returns (w*gamma as the bf16 operand, b + w beta f32, colsum f32 of the bf16-rounded operand, so the
    mean term cancels against exactly what the tensor cores multiply).  One-time weight preparation."""
    wf = w.detach().float()
    wg = (wf * gamma.detach().float()[None, :]).to(torch.bfloat16).contiguous()
    bias = wf @ beta.detach().float()
    if b is not None:
        bias = bias + b.detach().float()
    return wg, bias.contiguous(), wg.float().sum(dim=1).contiguous()
