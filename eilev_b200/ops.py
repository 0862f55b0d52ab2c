"""Tensor-level wrappers over the C ABI (``include/videoblip_b200.h``).

PyTorch is used only for device memory and stream handles: each function takes CUDA
tensors, enqueues one native kernel family on torch's current stream and returns.
There is no CPU path — a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import (EPI_GELU, EPI_GELU_BWD, EPI_NONE, EPI_RELU, EPI_RELU_BWD, GEMM_AUTO, GEMM_GENERIC,  # noqa: F401
                   GEMM_TCGEN05, VB_BF16, VB_F16, VB_F32, AttnArgs, AttnBwdArgs, GemmArgs, check)

_DT = {torch.bfloat16: VB_BF16, torch.float32: VB_F32, torch.float16: VB_F16}


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: torch.Tensor, dtype: torch.dtype | None, name: str) -> None:
    if not t.is_cuda:
        raise _lib.VbError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.VbError(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *,
         residual: torch.Tensor | None = None, out: torch.Tensor | None = None,
         epilogue: int = EPI_NONE, alpha: float = 1.0, alpha_cols: int = 0, beta: float = 0.0,
         out_dtype: torch.dtype = torch.bfloat16, row_group: int = 0, out_rows: int | None = None,
         backend: int = GEMM_AUTO, block_n: int = 0, dropout=None, ln_fold=None,
         stats_out: torch.Tensor | None = None, stats_zero: torch.Tensor | None = None) -> torch.Tensor:
    """``out = act(alpha * (a @ w.T + bias)) + residual (+ beta*out)``.

    a: (M, K) bf16 (last dim contiguous), w: (N, K) bf16 — an ``nn.Linear`` weight.
    ln_fold = (stats (M, 2) f64, colsum (N) f32, eps): ``a`` is the UN-normalised input of a LayerNorm
    whose gamma is already folded into ``w`` and whose beta into ``bias`` (pack_ln_fold); the epilogue
    applies the per-row mean / rstd.  stats_out (rows, 2) f64, zeroed by the caller: receives the row
    [sum, sum of squares] of the stored output for the next folded LayerNorm.  stats_zero (M, 2) f64: a
    statistics buffer the stream has finished reading, cleared by this launch for its next producer.
    """
    _need(a, torch.bfloat16, "gemm.a")
    _need(w, torch.bfloat16, "gemm.w")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1
    m, k = a.shape
    n = w.shape[0]
    if out is None:
        rows = out_rows if out_rows is not None else m
        out = torch.empty((rows, n), dtype=out_dtype, device=a.device)
    else:
        assert out.stride(-1) == 1
        out_dtype = out.dtype
    if bias is not None:
        _need(bias, torch.float32, "gemm.bias")
        assert bias.numel() == n
    if residual is not None:
        _need(residual, torch.bfloat16, "gemm.residual")
        assert residual.stride(-1) == 1
    args = GemmArgs()
    args.a, args.b, args.c = a.data_ptr(), w.data_ptr(), out.data_ptr()
    args.bias, args.residual = _ptr(bias), _ptr(residual)
    args.m, args.n, args.k = m, n, k
    args.lda, args.ldb = a.stride(0), w.stride(0)
    args.ldc = out.stride(-2) if out.dim() >= 2 else n
    args.ldr = residual.stride(-2) if residual is not None and residual.dim() >= 2 else 0
    args.alpha, args.beta = alpha, beta
    args.alpha_cols, args.row_group = alpha_cols, row_group
    args.epilogue, args.out_dtype, args.backend = epilogue, _DT[out_dtype], backend
    args.reserved = block_n
    if dropout is not None and dropout[0] > 0.0:  # (p, seed tensor (uint64/int64 on device), salt)
        args.dropout_p, args.dropout_seed, args.dropout_salt = float(dropout[0]), dropout[1].data_ptr(), int(dropout[2])
    if ln_fold is not None:
        stats, colsum, eps = ln_fold
        _need(stats, torch.float64, "gemm.ln_fold.stats")
        _need(colsum, torch.float32, "gemm.ln_fold.colsum")
        assert stats.is_contiguous() and stats.shape == (m, 2) and colsum.numel() == n
        args.ln_stats, args.ln_colsum, args.ln_eps = stats.data_ptr(), colsum.data_ptr(), float(eps)
    if stats_out is not None:
        _need(stats_out, torch.float64, "gemm.stats_out")
        assert stats_out.is_contiguous() and stats_out.shape == (out.shape[0], 2)
        args.stats_out = stats_out.data_ptr()
    if stats_zero is not None:
        _need(stats_zero, torch.float64, "gemm.stats_zero")
        assert stats_zero.is_contiguous() and stats_zero.shape == (m, 2)
        args.stats_zero = stats_zero.data_ptr()
    check(_lib.lib().vb_gemm(C.byref(args), _stream()), "vb_gemm")
    return out


def gemm_tn(dy: torch.Tensor, x: torch.Tensor, *, out: torch.Tensor | None = None, beta: float = 0.0,
            out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``out = dy.T @ x (+ beta * out)``: the weight-gradient product of a linear layer, dy: (tokens, N_out),
    x: (tokens, N_in) bf16, out: (N_out, N_in).  The activations enter the tcgen05 instruction as MN-major tiles
    (vb_gemm_args.operand_layout = 1), so nothing is transposed; shapes the tensor maps cannot address (dims
    that are not multiples of 8, unaligned views) go through two transpose kernels and the plain GEMM."""
    _need(dy, torch.bfloat16, "gemm_tn.dy")
    _need(x, torch.bfloat16, "gemm_tn.x")
    assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0]
    m, n, k = dy.shape[1], x.shape[1], dy.shape[0]
    ok = (m % 8 == 0 and n % 8 == 0 and dy.stride(1) == 1 and x.stride(1) == 1 and dy.stride(0) % 8 == 0
          and x.stride(0) % 8 == 0 and dy.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0
          # (VB_GEMM_TN_MAXK: A/B knob; long reductions such as the cross-K|V weight gradient over 34 952 tokens
          # used to keep two transposes + the CTA-pair kernel: 0.5 ms per step slower, 740 MB of scratch)
          and k <= int(os.environ.get("VB_GEMM_TN_MAXK", str(1 << 30)))
          and os.environ.get("VB_GEMM_TN", "1") != "0")
    if not ok:
        return gemm(transpose(dy), transpose(x), out=out, beta=beta, out_dtype=out_dtype)
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=dy.device)
    else:
        assert out.shape == (m, n) and out.stride(-1) == 1
    if out.data_ptr() % 16 != 0 or out.stride(0) % (8 if out.dtype == torch.bfloat16 else 4) != 0:
        return gemm(transpose(dy), transpose(x), out=out, beta=beta, out_dtype=out_dtype)
    args = GemmArgs()
    args.a, args.b, args.c = dy.data_ptr(), x.data_ptr(), out.data_ptr()
    args.m, args.n, args.k = m, n, k
    args.lda, args.ldb, args.ldc = dy.stride(0), x.stride(0), out.stride(0)
    args.alpha, args.beta = 1.0, beta
    args.epilogue, args.out_dtype, args.backend = EPI_NONE, _DT[out.dtype], GEMM_TCGEN05
    args.reserved2 = 1
    check(_lib.lib().vb_gemm(C.byref(args), _stream()), "vb_gemm")
    return out


def row_stats(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """(rows, 2) f64 [sum, sum of squares] of every row of a 2-D bf16 tensor (see gemm(ln_fold=...))."""
    _need(x, torch.bfloat16, "row_stats.x")
    assert x.dim() == 2 and x.stride(1) == 1
    if out is None:
        out = torch.empty((x.shape[0], 2), dtype=torch.float64, device=x.device)
    assert out.dtype == torch.float64 and out.is_contiguous()
    check(_lib.lib().vb_row_stats(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], x.stride(0), _stream()),
          "vb_row_stats")
    return out


def gemm_act_bwd(dy: torch.Tensor, w_t: torch.Tensor, saved: torch.Tensor, act: int) -> torch.Tensor:
    """d_pre = (dy @ w_t.T) * act'(saved): the dgrad GEMM of the layer AFTER an activation with the activation's
    backward in its epilogue (saved = the pre-activation for GELU, the activation output for ReLU)."""
    epi = {EPI_GELU: EPI_GELU_BWD, EPI_RELU: EPI_RELU_BWD}[act]
    assert saved.shape == (dy.shape[0], w_t.shape[0]) and saved.dtype == torch.bfloat16
    if os.environ.get("VB_ACT_BWD_FUSED", "1") == "0":  # measurement knob: the two-kernel form
        return act_bwd(gemm(dy, w_t), saved, act)
    return gemm(dy, w_t, residual=saved, epilogue=epi)


def gemm_uses_tcgen05(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor) -> bool:
    args = GemmArgs()
    args.a, args.b, args.c = a.data_ptr(), w.data_ptr(), out.data_ptr()
    args.m, args.k = a.shape
    args.n = w.shape[0]
    args.lda, args.ldb, args.ldc = a.stride(0), w.stride(0), out.stride(0)
    args.out_dtype = _DT[out.dtype]
    return bool(_lib.lib().vb_gemm_uses_tcgen05(C.byref(args)))


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *,
              residual: torch.Tensor | None = None, save_stats: bool = False,
              out: torch.Tensor | None = None):
    """LayerNorm over the last dim of a 2-D bf16 tensor (optionally of x + residual)."""
    _need(x, torch.bfloat16, "layernorm.x")
    _need(gamma, torch.float32, "layernorm.gamma")
    _need(beta, torch.float32, "layernorm.beta")
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    if out is None:
        y = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    else:
        y = out
        assert y.shape == (rows, cols) and y.stride(1) == 1 and y.dtype == torch.bfloat16
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    if residual is not None:
        _need(residual, torch.bfloat16, "layernorm.residual")
        assert residual.shape == x.shape and residual.stride(1) == 1
    check(_lib.lib().vb_layernorm(x.data_ptr(), _ptr(residual), gamma.data_ptr(), beta.data_ptr(),
                                  y.data_ptr(), _ptr(mean), _ptr(rstd), rows, cols, x.stride(0),
                                  residual.stride(0) if residual is not None else 0, y.stride(0), eps,
                                  _stream()), "vb_layernorm")
    return (y, mean, rstd) if save_stats else y


def layernorm_bwd(dy, xin, gamma, mean, rstd, *, dx_add=None, dgamma=None, dbeta=None, dropout=None):
    """dx of a LayerNorm (+ dx_add).  dropout = (p, seed tensor, salt): also returns dropout(dx) with that mask,
    written in the same pass -> (dx, dx_dropped)."""
    _need(dy, torch.bfloat16, "layernorm_bwd.dy")
    _need(xin, torch.bfloat16, "layernorm_bwd.xin")
    assert dy.is_contiguous() and xin.is_contiguous()
    rows, cols = dy.shape
    dx = torch.empty_like(dy)
    if dx_add is not None:
        assert dx_add.is_contiguous()
    if dropout is not None and dropout[0] > 0.0:
        assert dgamma is None and dbeta is None
        dx_drop = torch.empty_like(dy)
        check(_lib.lib().vb_layernorm_bwd_dropout(dy.data_ptr(), xin.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
                                                  rstd.data_ptr(), _ptr(dx_add), dx.data_ptr(), dx_drop.data_ptr(),
                                                  float(dropout[0]), dropout[1].data_ptr(), int(dropout[2]), rows, cols,
                                                  _stream()), "vb_layernorm_bwd_dropout")
        return dx, dx_drop
    check(_lib.lib().vb_layernorm_bwd(dy.data_ptr(), xin.data_ptr(), gamma.data_ptr(),
                                      mean.data_ptr(), rstd.data_ptr(), _ptr(dx_add), dx.data_ptr(),
                                      _ptr(dgamma), _ptr(dbeta), rows, cols, 0.0, _stream()),
          "vb_layernorm_bwd")
    return dx


def _attn_args(q, k, v, o, lse, key_mask, heads, d, scale, causal, dropout=None, rel_bias=None) -> AttnArgs:
    """q: (B, Sq, >=H*D) view, k/v: (B, Skv, ...) views; last dim contiguous."""
    a = AttnArgs()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    a.lse, a.key_mask = _ptr(lse), _ptr(key_mask)
    a.batch, a.sq, a.skv = q.shape[0], q.shape[1], k.shape[1]
    a.heads, a.d = heads, d
    a.q_bs, a.q_rs = q.stride(0), q.stride(1)
    a.k_bs, a.k_rs = k.stride(0), k.stride(1)
    a.v_bs, a.v_rs = v.stride(0), v.stride(1)
    a.o_bs, a.o_rs = o.stride(0), o.stride(1)
    a.scale, a.causal = scale, 1 if causal else 0
    if dropout is not None and dropout[0] > 0.0:
        a.dropout_p, a.dropout_seed, a.dropout_salt = float(dropout[0]), dropout[1].data_ptr(), int(dropout[2])
    if rel_bias is not None:  # (heads, Sq + Skv - 1) f32 relative-position table (T5)
        _need(rel_bias, torch.float32, "attention.rel_bias")
        assert rel_bias.dim() == 2 and rel_bias.shape[0] == heads and (rel_bias.shape[1] == 1 or rel_bias.stride(1) == 1)
        assert rel_bias.shape[1] == q.shape[1] + k.shape[1] - 1
        a.rel_bias, a.rel_bias_stride = rel_bias.data_ptr(), rel_bias.stride(0)
    return a


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: float, *,
              causal: bool = False, key_mask: torch.Tensor | None = None,
              need_lse: bool = False, dropout=None, rel_bias: torch.Tensor | None = None):
    """Softmax attention.  q: (B, Sq, H*D), k/v: (B, Skv, H*D) bf16 views whose last dim is
    contiguous (they may be slices of one fused QKV buffer).  Returns o: (B, Sq, H*D)."""
    for t, nme in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, torch.bfloat16, f"attention.{nme}")
        assert t.dim() == 3 and t.stride(2) == 1
    hd = q.shape[2]
    d = hd // heads
    o = torch.empty((q.shape[0], q.shape[1], hd), dtype=torch.bfloat16, device=q.device)
    lse = None
    if need_lse:
        lse = torch.empty((q.shape[0], heads, q.shape[1]), dtype=torch.float32, device=q.device)
    if key_mask is not None:
        _need(key_mask, torch.uint8, "attention.key_mask")
        assert key_mask.is_contiguous() and key_mask.shape == (k.shape[0], k.shape[1])
    a = _attn_args(q, k, v, o, lse, key_mask, heads, d, scale, causal, dropout, rel_bias)
    check(_lib.lib().vb_attention_fwd(C.byref(a), _stream()), "vb_attention_fwd")
    return (o, lse) if need_lse else o


def attention_probs(q: torch.Tensor, k: torch.Tensor, heads: int, scale: float, *, causal: bool = False,
                    key_mask: torch.Tensor | None = None, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """The attention maps softmax(scale * q k^T + masks), (B, H, Sq, Skv): the optional ``attentions`` output
    (output_attentions=True), computed by a plain kernel off the hot path."""
    for t, nme in ((q, "q"), (k, "k")):
        _need(t, torch.bfloat16, f"attention_probs.{nme}")
        assert t.dim() == 3 and t.stride(2) == 1
    hd = q.shape[2]
    probs = torch.empty((q.shape[0], heads, q.shape[1], k.shape[1]), dtype=dtype, device=q.device)
    a = _attn_args(q, k, k, q, None, key_mask, heads, hd // heads, scale, causal)
    check(_lib.lib().vb_attention_probs(C.byref(a), probs.data_ptr(), _DT[dtype], _stream()), "vb_attention_probs")
    return probs


def attention_uses_tcgen05(q, k, v, heads: int, *, causal=False, key_mask=None, need_lse=False) -> bool:
    hd = q.shape[2]
    o = torch.empty((q.shape[0], q.shape[1], hd), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty(1, device=q.device) if need_lse else None
    a = _attn_args(q, k, v, o, lse, key_mask, heads, hd // heads, 1.0, causal)
    return _lib.lib().vb_attention_uses_tcgen05(C.byref(a)) == 1


def attention_kernel(q, k, v, heads: int, *, causal=False, key_mask=None, need_lse=False) -> str:
    """The kernel ``attention`` runs for these arguments: "tcgen05_vit", "tcgen05_flash" or "mma_sync"."""
    hd = q.shape[2]
    o = torch.empty((q.shape[0], q.shape[1], hd), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty(1, device=q.device) if need_lse else None
    a = _attn_args(q, k, v, o, lse, key_mask, heads, hd // heads, 1.0, causal)
    return ("mma_sync", "tcgen05_vit", "tcgen05_flash")[_lib.lib().vb_attention_uses_tcgen05(C.byref(a))]


def attention_bwd(q, k, v, o, lse, d_o, heads: int, scale: float, *, causal: bool = False,
                  key_mask=None, dq_scale: float = 1.0, dq=None, dk=None, dv=None, dropout=None, rel_bias=None,
                  _probe: bool = False):
    """Returns (dq, dk, dv) with the shapes of q, k, v (contiguous unless views are given)."""
    hd = q.shape[2]
    d = hd // heads
    assert d_o.stride() == o.stride() and d_o.dtype == torch.bfloat16
    dq = torch.empty(q.shape, dtype=torch.bfloat16, device=q.device) if dq is None else dq
    dk = torch.empty(k.shape, dtype=torch.bfloat16, device=q.device) if dk is None else dk
    dv = torch.empty(v.shape, dtype=torch.bfloat16, device=q.device) if dv is None else dv
    delta = torch.empty((q.shape[0], heads, q.shape[1]), dtype=torch.float32, device=q.device)
    dq_acc = torch.empty((q.shape[0], q.shape[1], hd), dtype=torch.float32, device=q.device)
    b = AttnBwdArgs()
    b.fwd = _attn_args(q, k, v, o, lse, key_mask, heads, d, scale, causal, dropout, rel_bias)
    b.d_o, b.dq, b.dk, b.dv = d_o.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    b.dq_bs, b.dq_rs = dq.stride(0), dq.stride(1)
    b.dk_bs, b.dk_rs = dk.stride(0), dk.stride(1)
    b.dv_bs, b.dv_rs = dv.stride(0), dv.stride(1)
    b.delta, b.dq_acc, b.dq_scale = delta.data_ptr(), dq_acc.data_ptr(), dq_scale
    if _probe:
        return bool(_lib.lib().vb_attention_bwd_uses_tcgen05(C.byref(b)))
    check(_lib.lib().vb_attention_bwd(C.byref(b), _stream()), "vb_attention_bwd")
    return dq, dk, dv


def attention_bwd_uses_tcgen05(*args, **kwargs) -> bool:
    """True when ``attention_bwd`` with these arguments runs the tcgen05 / TMEM kernels."""
    return attention_bwd(*args, _probe=True, **kwargs)


def attention_merge(o1: torch.Tensor, lse1: torch.Tensor, o2: torch.Tensor, lse2: torch.Tensor,
                    heads: int) -> torch.Tensor:
    """Merges two attention partials over disjoint key sets.  o_i: (B_i, S_i, H*D) bf16
    contiguous with B_1*S_1 == B_2*S_2 rows in the same order; lse_i: (B_i, H, S_i) f32."""
    _need(o1, torch.bfloat16, "attention_merge.o1")
    _need(o2, torch.bfloat16, "attention_merge.o2")
    assert o1.is_contiguous() and o2.is_contiguous() and lse1.is_contiguous() and lse2.is_contiguous()
    rows = o1.shape[0] * o1.shape[1]
    assert rows == o2.shape[0] * o2.shape[1] and o1.shape[2] == o2.shape[2]
    hd = o1.shape[2]
    out = torch.empty_like(o2)
    check(_lib.lib().vb_attention_merge(o1.data_ptr(), lse1.data_ptr(), o1.shape[1], o2.data_ptr(),
                                        lse2.data_ptr(), o2.shape[1], out.data_ptr(), rows, heads,
                                        hd // heads, _stream()), "vb_attention_merge")
    return out


def token_logprob(logits: torch.Tensor, targets: torch.Tensor, row_index: torch.Tensor | None = None) -> torch.Tensor:
    """log_softmax(logits[row])[target] per target (0 for targets outside the vocabulary)."""
    _need(logits, None, "token_logprob.logits")
    _need(targets, torch.int64, "token_logprob.targets")
    assert logits.dim() == 2 and logits.stride(1) == 1 and logits.dtype in (torch.bfloat16, torch.float32)
    targets = targets.contiguous().view(-1)
    if row_index is not None:
        _need(row_index, torch.int64, "token_logprob.row_index")
        row_index = row_index.contiguous().view(-1)
        assert row_index.numel() == targets.numel()
    else:
        assert targets.numel() == logits.shape[0]
    out = torch.empty(targets.numel(), dtype=torch.float32, device=logits.device)
    check(_lib.lib().vb_token_logprob(logits.data_ptr(), _DT[logits.dtype], _ptr(row_index),
                                      targets.data_ptr(), out.data_ptr(), targets.numel(), logits.shape[1],
                                      logits.stride(0), _stream()), "vb_token_logprob")
    return out


def patch_gather(pixels: torch.Tensor, patch: int, kpad: int) -> torch.Tensor:
    """(NV, C, T, H, W) -> (NV*T*gh*gw, kpad) bf16 patch matrix."""
    _need(pixels, None, "patch_gather.pixels")
    assert pixels.dim() == 5 and pixels.is_contiguous() and pixels.dtype in _DT
    nv, c, t, h, w = pixels.shape
    gh, gw = h // patch, w // patch
    out = torch.empty((nv * t * gh * gw, kpad), dtype=torch.bfloat16, device=pixels.device)
    check(_lib.lib().vb_patch_gather(pixels.data_ptr(), _DT[pixels.dtype], out.data_ptr(), nv, c, t,
                                     h, w, patch, kpad, _stream()), "vb_patch_gather")
    return out


def patch_gather_u8(frames: torch.Tensor, patch: int, kpad: int, rescale: float, mean, std) -> torch.Tensor:
    """uint8 (NV, C, T, H, W) -> (NV*T*gh*gw, kpad) bf16 patch matrix of the normalised frames:
    (u8 * rescale - mean[c]) / std[c] in fp32, fused into the gather."""
    _need(frames, torch.uint8, "patch_gather_u8.frames")
    assert frames.dim() == 5 and frames.is_contiguous()
    nv, c, t, h, w = frames.shape
    if len(mean) != c or len(std) != c or c > 4:
        raise ValueError(f"patch_gather_u8: {c} channels need {c} (<= 4) mean / std values, got {len(mean)} / {len(std)}")
    gh, gw = h // patch, w // patch
    out = torch.empty((nv * t * gh * gw, kpad), dtype=torch.bfloat16, device=frames.device)
    arr = C.c_float * c
    check(_lib.lib().vb_patch_gather_u8(frames.data_ptr(), out.data_ptr(), nv, c, t, h, w, patch, kpad,
                                        float(rescale), arr(*[float(m) for m in mean]),
                                        arr(*[float(v) for v in std]), _stream()), "vb_patch_gather_u8")
    return out


def crop_resize_normalize_u8(frames: torch.Tensor, box: tuple[int, int, int, int], out_size: tuple[int, int],
                             rescale: float, mean, std, *, flip: bool = False,
                             out: torch.Tensor | None = None, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Training-time frame transform of one decoded clip (train_v2.py:143-167): uint8 (C, T, H, W) ->
    (C, T, out_h, out_w) f32 | bf16 = flip(bicubic_interpolate(crop(frames, box))) * rescale, normalised.
    box = (top, left, height, width), drawn on the host (data/preprocess.py::resized_crop_params)."""
    _need(frames, torch.uint8, "crop_resize_normalize_u8.frames")
    assert frames.dim() == 4 and frames.is_contiguous()
    c, t, h, w = frames.shape
    top, left, ch, cw = (int(v) for v in box)
    oh, ow = int(out_size[0]), int(out_size[1])
    if len(mean) != c or len(std) != c or c > 4:
        raise ValueError(f"crop_resize_normalize_u8: {c} channels need {c} (<= 4) mean / std values")
    if out is None:
        out = torch.empty((c, t, oh, ow), dtype=dtype, device=frames.device)
    else:
        assert out.shape == (c, t, oh, ow) and out.is_contiguous() and out.dtype in (torch.float32, torch.bfloat16)
    arr = C.c_float * c
    check(_lib.lib().vb_crop_resize_normalize_u8(frames.data_ptr(), c, t, h, w, top, left, ch, cw, int(bool(flip)),
                                                 out.data_ptr(), _DT[out.dtype], oh, ow, float(rescale),
                                                 arr(*[float(m) for m in mean]), arr(*[float(v) for v in std]),
                                                 _stream()), "vb_crop_resize_normalize_u8")
    return out


_RESIZE_TABLES: dict = {}


def resize_coeffs(in_size: int, out_size: int):
    """Host tables of one axis of Pillow's antialiased bicubic: (ksize, bounds (out, 2) int32 =
    [first tap, tap count], kk (out, ksize) int32 22-bit fixed-point weights).  Host arithmetic only
    (``vb_resize_bicubic_coeffs``): no device work, no launch."""
    lib = _lib.lib()
    ksize = int(lib.vb_resize_bicubic_ksize(in_size, out_size))
    if ksize <= 0:
        raise ValueError(f"resize_coeffs: bad sizes {in_size} -> {out_size}")
    bounds = torch.empty((out_size, 2), dtype=torch.int32)
    kk = torch.empty((out_size, ksize), dtype=torch.int32)
    rc = lib.vb_resize_bicubic_coeffs(in_size, out_size, C.cast(bounds.data_ptr(), C.POINTER(C.c_int32)),
                                      C.cast(kk.data_ptr(), C.POINTER(C.c_int32)), kk.numel())
    if rc != 0:
        raise _lib.VbError(f"vb_resize_bicubic_coeffs failed: {lib.vb_last_error().decode()}")
    return ksize, bounds, kk


def resize_plan(in_h: int, in_w: int, out_h: int, out_w: int) -> list[dict]:
    """The passes of Pillow's two-pass resample (ImagingResampleInner) as ``vb_resize_u8_pass``
    argument sets — pure integer bookkeeping, shared by the device path and its CPU emulation test.
    Horizontal pass first (when the width changes), restricted to source rows [first, last) that
    the vertical pass reads; then the vertical pass with its windows shifted by ``first``."""
    need_h, need_v = out_w != in_w, out_h != in_h
    passes: list[dict] = []
    first, rows = 0, in_h
    if need_v:
        _, vb, _ = resize_coeffs(in_h, out_h)
        if need_h:
            first = int(vb[0, 0])
            rows = int(vb[-1, 0] + vb[-1, 1]) - first
    if need_h:
        passes.append(dict(axis=(in_w, out_w), shift=0, in_offset=first * in_w, lines=rows, out_len=out_w,
                           in_strides=(in_h * in_w, in_w, 1), out_strides=(rows * out_w, out_w, 1),
                           out_shape=(rows, out_w), lines_fastest=0))
    if need_v:
        passes.append(dict(axis=(in_h, out_h), shift=first, in_offset=0, lines=out_w, out_len=out_h,
                           in_strides=(rows * out_w, 1, out_w), out_strides=(out_h * out_w, 1, out_w),
                           out_shape=(out_h, out_w), lines_fastest=1))
    return passes


def _resize_tables(in_size: int, out_size: int, device, shift: int):
    key = (in_size, out_size, shift, str(device))
    hit = _RESIZE_TABLES.get(key)
    if hit is None:
        ksize, bounds, kk = resize_coeffs(in_size, out_size)
        bounds = bounds.clone()
        bounds[:, 0] -= shift
        hit = (ksize, bounds.to(device), kk.to(device))
        if len(_RESIZE_TABLES) >= 64:  # videos of many resolutions: keep the table cache bounded
            _RESIZE_TABLES.pop(next(iter(_RESIZE_TABLES)))
        _RESIZE_TABLES[key] = hit
    return hit


def resize_bicubic_u8(frames: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """uint8 (..., H, W) -> (..., out_h, out_w) uint8: every plane resized as
    ``PIL.Image.resize((out_w, out_h), BICUBIC)`` resizes it, bit for bit (Pillow's two-pass 8-bit
    resample: horizontal pass over the source rows the vertical pass reads, then the vertical pass)."""
    _need(frames, torch.uint8, "resize_bicubic_u8.frames")
    assert frames.dim() >= 2
    frames = frames.contiguous()
    in_h, in_w = frames.shape[-2:]
    lead = frames.shape[:-2]
    planes = frames.numel() // (in_h * in_w) if in_h * in_w else 0
    passes = resize_plan(in_h, in_w, out_h, out_w)
    if not passes or planes == 0:
        return frames.clone()
    cur = frames
    for ps in passes:
        ksize, bounds, kk = _resize_tables(*ps["axis"], frames.device, ps["shift"])
        out = torch.empty((planes, *ps["out_shape"]), dtype=torch.uint8, device=frames.device)
        check(_lib.lib().vb_resize_u8_pass(cur.data_ptr() + ps["in_offset"], out.data_ptr(), bounds.data_ptr(),
                                           kk.data_ptr(), planes, ps["lines"], ps["out_len"], ksize,
                                           *ps["in_strides"], *ps["out_strides"], ps["lines_fastest"], _stream()),
              "vb_resize_u8_pass")
        cur = out
    return cur.view(*lead, out_h, out_w)


def cls_rows(cls: torch.Tensor, pos: torch.Tensor, hidden: torch.Tensor) -> None:
    frames, tokens, dim = hidden.shape
    check(_lib.lib().vb_cls_rows(cls.data_ptr(), pos.data_ptr(), hidden.data_ptr(), frames, tokens,
                                 dim, _stream()), "vb_cls_rows")


def embed_splice(input_ids, attention_mask, video_mask, embed_tokens, video_features, pos_table,
                 pos_offset: int, *, want_embeds: bool = True, want_hidden: bool = True):
    """Returns (inputs_embeds|None, hidden|None, slot_index, pos_ids, status)."""
    _need(input_ids, torch.int64, "embed_splice.input_ids")
    batch, seq = input_ids.shape
    vocab, dim = embed_tokens.shape
    dev = input_ids.device
    input_ids = input_ids.contiguous()
    if attention_mask is not None:
        attention_mask = attention_mask.to(torch.int64).contiguous()
    if video_mask is not None:
        video_mask = video_mask.to(torch.int64).contiguous()
    n_feat = 0 if video_features is None else video_features.shape[0]
    if n_feat == 0:
        video_mask = None  # nothing to splice: the reference ignores the mask without pixel values (v2.py:205-213)
    if video_features is not None:
        _need(video_features, torch.bfloat16, "embed_splice.video_features")
        assert video_features.is_contiguous() and video_features.shape[1] == dim
    emb = torch.empty((batch, seq, dim), dtype=torch.bfloat16, device=dev) if want_embeds else None
    hid = torch.empty((batch, seq, dim), dtype=torch.bfloat16, device=dev) if want_hidden else None
    slot = torch.empty(batch * seq, dtype=torch.int32, device=dev)
    pos = torch.empty(batch * seq, dtype=torch.int32, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    check(_lib.lib().vb_embed_splice(input_ids.data_ptr(), _ptr(attention_mask), _ptr(video_mask),
                                     embed_tokens.data_ptr(), _ptr(video_features), _ptr(pos_table),
                                     pos_offset, _ptr(emb), _ptr(hid), slot.data_ptr(),
                                     pos.data_ptr(), status.data_ptr(), batch, seq, dim, vocab,
                                     n_feat, _stream()), "vb_embed_splice")
    return emb, hid, slot, pos, status


def splice_bwd(d_embeds: torch.Tensor, slot_index: torch.Tensor, n_features: int) -> torch.Tensor:
    dim = d_embeds.shape[-1]
    positions = d_embeds.numel() // dim
    out = torch.zeros((n_features, dim), dtype=torch.bfloat16, device=d_embeds.device)
    check(_lib.lib().vb_splice_bwd(d_embeds.data_ptr(), slot_index.data_ptr(), out.data_ptr(),
                                   positions, dim, n_features, _stream()), "vb_splice_bwd")
    return out


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, shift: int = 1):
    """Mean CE with ignore_index -100.  shift=1: causal-LM loss (position l is scored against
    labels[l+1]); shift=0: seq2seq loss (T5, labels aligned with the logits).
    logits (B, L, V) bf16|f32; labels (B, L) int64.  Returns (loss f32 scalar, row_lse, n_valid)."""
    assert logits.dim() == 3 and logits.stride(2) == 1 and logits.dtype in (torch.bfloat16, torch.float32)
    b, l, v = logits.shape
    assert logits.stride(0) == l * logits.stride(1)
    labels = labels.contiguous()
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    row_lse = torch.empty(b * l, dtype=torch.float32, device=logits.device)
    n_valid = torch.empty((), dtype=torch.int32, device=logits.device)
    check(_lib.lib().vb_cross_entropy(logits.data_ptr(), _DT[logits.dtype], labels.data_ptr(),
                                      loss.data_ptr(), row_lse.data_ptr(), n_valid.data_ptr(), b, l,
                                      v, logits.stride(1), int(shift), _stream()), "vb_cross_entropy")
    return loss, row_lse, n_valid


def cross_entropy_bwd(logits, labels, row_lse, n_valid, grad_scale: torch.Tensor | None, shift: int = 1):
    b, l, v = logits.shape
    vpad = (v + 7) // 8 * 8
    d = torch.empty((b * l, vpad), dtype=torch.bfloat16, device=logits.device)
    if vpad != v:
        d[:, v:].zero_()
    check(_lib.lib().vb_cross_entropy_bwd(logits.data_ptr(), _DT[logits.dtype],
                                          labels.contiguous().data_ptr(), row_lse.data_ptr(),
                                          n_valid.data_ptr(), _ptr(grad_scale), d.data_ptr(), b, l,
                                          v, logits.stride(1), vpad, int(shift), _stream()),
          "vb_cross_entropy_bwd")
    return d[:, :v]


def rmsnorm(x: torch.Tensor, gamma: torch.Tensor, eps: float, *, save_stats: bool = False):
    """T5LayerNorm over the last dim of a 2-D bf16 tensor: gamma * x * rsqrt(mean(x^2) + eps)."""
    _need(x, torch.bfloat16, "rmsnorm.x")
    _need(gamma, torch.float32, "rmsnorm.gamma")
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    y = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    check(_lib.lib().vb_rmsnorm(x.data_ptr(), gamma.data_ptr(), y.data_ptr(), _ptr(rstd), rows, cols,
                                x.stride(0), y.stride(0), float(eps), _stream()), "vb_rmsnorm")
    return (y, rstd) if save_stats else y


def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, rstd: torch.Tensor, *, dx_add=None):
    assert dy.is_contiguous() and x.is_contiguous() and dy.shape == x.shape
    if dx_add is not None:
        assert dx_add.is_contiguous() and dx_add.shape == x.shape
    dx = torch.empty_like(dy)
    check(_lib.lib().vb_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), rstd.data_ptr(),
                                    _ptr(dx_add), dx.data_ptr(), x.shape[0], x.shape[1], _stream()),
          "vb_rmsnorm_bwd")
    return dx


def gated_gelu(h01: torch.Tensor) -> torch.Tensor:
    """(rows, 2*dff) [wi_0 x | wi_1 x] -> gelu_new(first half) * second half, (rows, dff)."""
    _need(h01, torch.bfloat16, "gated_gelu.h01")
    assert h01.dim() == 2 and h01.is_contiguous() and h01.shape[1] % 2 == 0
    rows, dff = h01.shape[0], h01.shape[1] // 2
    out = torch.empty((rows, dff), dtype=torch.bfloat16, device=h01.device)
    check(_lib.lib().vb_gated_gelu(h01.data_ptr(), out.data_ptr(), rows, dff, _stream()), "vb_gated_gelu")
    return out


def gated_gelu_bwd(d_out: torch.Tensor, h01: torch.Tensor) -> torch.Tensor:
    assert d_out.is_contiguous() and h01.is_contiguous() and h01.shape[1] == 2 * d_out.shape[1]
    d = torch.empty_like(h01)
    check(_lib.lib().vb_gated_gelu_bwd(d_out.data_ptr(), h01.data_ptr(), d.data_ptr(), d_out.shape[0],
                                       d_out.shape[1], _stream()), "vb_gated_gelu_bwd")
    return d


def embedding(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """table[ids] for a bf16 table; ids int64 of any shape -> (*ids.shape, dim)."""
    _need(ids, torch.int64, "embedding.ids")
    _need(table, torch.bfloat16, "embedding.table")
    assert table.is_contiguous()
    flat = ids.contiguous().view(-1)
    out = torch.empty((flat.numel(), table.shape[1]), dtype=torch.bfloat16, device=table.device)
    check(_lib.lib().vb_embedding(flat.data_ptr(), table.data_ptr(), out.data_ptr(), flat.numel(),
                                  table.shape[1], table.shape[0], _stream()), "vb_embedding")
    return out.view(*ids.shape, table.shape[1])


def transpose(x: torch.Tensor) -> torch.Tensor:
    """(rows, cols) bf16 -> contiguous (cols, rows_padded-to-8)[:, :rows]."""
    _need(x, torch.bfloat16, "transpose.x")
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    ld = (rows + 7) // 8 * 8
    out = torch.empty((cols, ld), dtype=torch.bfloat16, device=x.device)
    if ld != rows:
        out[:, rows:].zero_()
    check(_lib.lib().vb_transpose(x.data_ptr(), out.data_ptr(), rows, cols, x.stride(0), ld,
                                  _stream()), "vb_transpose")
    return out[:, :rows]


def convert(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _need(x, None, "convert.x")
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(_lib.lib().vb_convert(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _DT[dtype], x.numel(),
                                _stream()), "vb_convert")
    return out


def act_bwd(dy: torch.Tensor, saved: torch.Tensor, epilogue: int) -> torch.Tensor:
    assert dy.is_contiguous() and saved.is_contiguous() and dy.shape == saved.shape
    dx = torch.empty_like(dy)
    check(_lib.lib().vb_act_bwd(dy.data_ptr(), saved.data_ptr(), dx.data_ptr(), epilogue,
                                dy.numel(), _stream()), "vb_act_bwd")
    return dx


def colsum(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """f32 column sums of a 2-D bf16 tensor (bias gradient); accumulates into `out` if given."""
    _need(x, torch.bfloat16, "colsum.x")
    assert x.dim() == 2 and x.stride(1) == 1
    acc = 1
    if out is None:
        out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
        acc = 0
    check(_lib.lib().vb_colsum(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], x.stride(0),
                               acc, _stream()), "vb_colsum")
    return out


def dropout(x: torch.Tensor, p: float, seed: torch.Tensor, salt: int) -> torch.Tensor:
    """y = x * mask / (1 - p) with the counter-hash mask of (seed + salt, flat index); 2-D bf16."""
    _need(x, torch.bfloat16, "dropout.x")
    assert x.dim() == 2 and x.stride(1) == 1
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().vb_dropout(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], x.stride(0), y.stride(0),
                                float(p), seed.data_ptr(), int(salt), _stream()), "vb_dropout")
    return y


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape
    y = torch.empty_like(a)
    check(_lib.lib().vb_add(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream()), "vb_add")
    return y


def adamw_(param, grad, exp_avg, exp_avg_sq, *, lr, beta1, beta2, eps, weight_decay, step,
           grad_scale: torch.Tensor | None = None) -> None:
    for t in (param, grad, exp_avg, exp_avg_sq):
        _need(t, torch.float32, "adamw")
        assert t.is_contiguous()
    check(_lib.lib().vb_adamw(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(),
                              exp_avg_sq.data_ptr(), param.numel(), lr, beta1, beta2, eps,
                              weight_decay, step, _ptr(grad_scale), _stream()), "vb_adamw")


def sumsq(x: torch.Tensor, out: torch.Tensor) -> None:
    _need(x, torch.float32, "sumsq.x")
    check(_lib.lib().vb_sumsq(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "vb_sumsq")


def gemv(x: torch.Tensor, w: torch.Tensor, bias=None, *, residual=None, epilogue: int = EPI_NONE,
         alpha: float = 1.0, alpha_cols: int = 0, out_dtype=torch.bfloat16, ln=None) -> torch.Tensor:
    """Decode-time projection for M <= 16 rows (weight-streaming).  ln = (gamma, beta, eps)
    fuses the LayerNorm of x into the kernel's staging pass; beta = None selects RMSNorm (T5)."""
    _need(x, torch.bfloat16, "gemv.x")
    _need(w, torch.bfloat16, "gemv.w")
    m, k = x.shape
    n = w.shape[0]
    if ln is not None and (k % 64 != 0 or m * (k + 32) * 2 + 8 * k + 16384 > 190 * 1024):
        # shapes the staged kernel does not take
        x = layernorm(x, ln[0], ln[1], ln[2]) if ln[1] is not None else rmsnorm(x, ln[0], ln[2])
        ln = None
    y = torch.empty((m, n), dtype=out_dtype, device=x.device)
    check(_lib.lib().vb_gemv(x.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(),
                             m, n, k, x.stride(0), w.stride(0), n,
                             residual.stride(0) if residual is not None else 0, alpha, alpha_cols,
                             epilogue, _DT[out_dtype], _ptr(ln[0]) if ln else None,
                             _ptr(ln[1]) if ln else None, float(ln[2]) if ln else 0.0, _stream()), "vb_gemv")
    return y


def decode_embed(tokens: torch.Tensor, embed: torch.Tensor, pos_table: torch.Tensor,
                 n_valid: torch.Tensor, ctx_len: torch.Tensor, pos_offset: int = 2) -> torch.Tensor:
    """x[b] = embed[tokens[b]] + pos_table[n_valid[b] + pos_offset]; advances n_valid / ctx_len
    in place (HF:opt/modeling_opt.py:45-70,350-354)."""
    _need(embed, torch.bfloat16, "decode_embed.embed")
    _need(pos_table, torch.bfloat16, "decode_embed.pos_table")
    _need(tokens, torch.int64, "decode_embed.tokens")
    _need(n_valid, torch.int32, "decode_embed.n_valid")
    _need(ctx_len, torch.int32, "decode_embed.ctx_len")
    assert embed.is_contiguous() and pos_table.is_contiguous() and tokens.is_contiguous()
    b, dim = tokens.shape[0], embed.shape[1]
    x = torch.empty((b, dim), dtype=torch.bfloat16, device=tokens.device)
    check(_lib.lib().vb_decode_embed(tokens.data_ptr(), embed.data_ptr(), pos_table.data_ptr(),
                                     n_valid.data_ptr(), ctx_len.data_ptr(), x.data_ptr(), b, dim,
                                     embed.shape[0], pos_table.shape[0], pos_offset, _stream()),
          "vb_decode_embed")
    return x


# ---- decode-step programs (vb_decode_op records, include/videoblip_b200.h) ----------------
OP_GEMV, OP_ATTN, OP_EMBED = 1, 2, 3
DECODE_STEP_WS_BYTES = 4096 + 1008 * 512  # VB_DECODE_STEP_WS_BYTES
_OP_DTYPE = None


def op_dtype():
    """numpy mirror of ``vb_decode_op`` (192 bytes)."""
    global _OP_DTYPE
    if _OP_DTYPE is None:
        import numpy as np
        _OP_DTYPE = np.dtype([("type", "<i4"), ("i32", "<i4", (7,)), ("ptr", "<u8", (10,)),
                              ("i64", "<i8", (8,)), ("f32", "<f4", (4,))])
        assert _OP_DTYPE.itemsize == 192
    return _OP_DTYPE


def decode_step(ops_host, ops_dev: torch.Tensor, m: int, barrier: torch.Tensor) -> None:
    """Runs a decode-step program (numpy record array + its device copy) as one persistent
    cooperative launch.  barrier: the program's int32 workspace (DECODE_STEP_WS_BYTES)."""
    _need(ops_dev, torch.uint8, "decode_step.ops_dev")
    _need(barrier, torch.int32, "decode_step.barrier")
    assert barrier.numel() * 4 >= DECODE_STEP_WS_BYTES
    check(_lib.lib().vb_decode_step(ops_host.ctypes.data, ops_dev.data_ptr(), int(ops_host.shape[0]), int(m),
                                    barrier.data_ptr(), _stream()), "vb_decode_step")


def decode_cross_attention(q, k, v, seq_ids, ctx_len, first_valid, heads: int, scale: float, *,
                           workspace, counters, splits: int) -> torch.Tensor:
    """One query per sequence over dense keys / values: q (B, H*D) bf16; k, v (B, L, H*D) bf16 views
    (e.g. column slices of one projection output) with the same strides."""
    _need(q, torch.bfloat16, "decode_cross_attention.q")
    assert k.dim() == 3 and k.stride(2) == 1 and k.stride() == v.stride() and k.stride(0) == k.shape[1] * k.stride(1)
    b, hd = q.shape
    out = torch.empty((b, hd), dtype=torch.bfloat16, device=q.device)
    check(_lib.lib().vb_decode_cross_attention(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(1),
                                               seq_ids.data_ptr(), ctx_len.data_ptr(), first_valid.data_ptr(),
                                               out.data_ptr(), workspace.data_ptr(), counters.data_ptr(), splits,
                                               b, heads, hd // heads, k.shape[1], scale, _stream()),
          "vb_decode_cross_attention")
    return out


def paged_kv_write(k, v, k_cache, v_cache, page_table, page_size: int) -> None:
    """k, v: (B, L, H*D) views with a common row stride."""
    b, l, hd = k.shape
    assert k.stride(2) == 1 and k.stride() == v.stride() and k.stride(0) == l * k.stride(1)
    check(_lib.lib().vb_paged_kv_write(k.data_ptr(), v.data_ptr(), k.stride(1), k_cache.data_ptr(),
                                       v_cache.data_ptr(), page_table.data_ptr(), b, l, hd,
                                       page_size, page_table.shape[1], _stream()),
          "vb_paged_kv_write")


def paged_decode_attention(qkv, k_cache, v_cache, page_table, ctx_len, first_valid, heads: int,
                           page_size: int, scale: float, *, workspace=None, counters=None,
                           splits: int = 8, rel_bias=None, rel_center: int = 0) -> torch.Tensor:
    """rel_bias (heads, n) f32 + rel_center: T5 decoder self-attention bias of cached token l seen
    from the newest position, rel_bias[h, rel_center + l - (ctx - 1)]."""
    b = qkv.shape[0]
    hd = qkv.shape[1] // 3
    d = hd // heads
    out = torch.empty((b, hd), dtype=torch.bfloat16, device=qkv.device)
    if workspace is None:
        workspace = torch.empty(b * heads * splits * (d + 2), dtype=torch.float32, device=qkv.device)
    if counters is None:
        counters = torch.zeros(b * heads, dtype=torch.int32, device=qkv.device)
    check(_lib.lib().vb_paged_decode_attention(qkv.data_ptr(), k_cache.data_ptr(),
                                               v_cache.data_ptr(), page_table.data_ptr(),
                                               ctx_len.data_ptr(), _ptr(first_valid),
                                               out.data_ptr(), workspace.data_ptr(),
                                               counters.data_ptr(), splits, b, heads, d, page_size,
                                               page_table.shape[1], scale, _ptr(rel_bias),
                                               rel_bias.stride(0) if rel_bias is not None else 0,
                                               int(rel_center), _stream()),
          "vb_paged_decode_attention")
    return out
