"""flan-T5 encoder-decoder LM over the interleaved (32 video queries per clip + text) sequence
— the seq2seq branch of eilev/model/v2.py:228-238.

Restates T5ForConditionalGeneration.forward (HF:t5/modeling_t5.py: T5Stack, T5Block,
T5Attention — no 1/sqrt(d) scaling, no projection biases, bucketed relative-position bias
owned by block 0 of each stack —, T5LayerNorm = RMSNorm, T5DenseGatedActDense with the tanh
GELU, ``_shift_right`` and the unshifted cross entropy) as sm_100a launches:

    encoder: embed_splice, N x [RMSNorm, fused QKV GEMM, flash attention with the bias table
             and the padding mask, o GEMM(+residual), RMSNorm, fused wi_0|wi_1 GEMM,
             gelu_new(h0)*h1, wo GEMM(+residual)], RMSNorm
    decoder: shared-embedding gather of shift_right(labels), ONE GEMM for the cross-attention
             K|V of all layers off the encoder output, M x [causal self-attention block,
             cross-attention block, gated FFN], RMSNorm, (untied) head GEMM, fused CE.

The LM is frozen in the recipe (train_v2.py:126-127): the backward is dgrad only, down through
the decoder, the cross-attention K/V projection, the encoder and the splice to the video slots.
"""
from __future__ import annotations

import math

import torch

from .. import ops
from .packing import PackCache, bf16, cat_bf16, f32

T = ops.transpose


def _check_cfg(cfg) -> None:
    gated_gelu = cfg.is_gated_act and cfg.dense_act_fn == "gelu_new"
    plain_relu = (not cfg.is_gated_act) and cfg.dense_act_fn == "relu"
    if not (gated_gelu or plain_relu):
        raise NotImplementedError(
            f"T5 feed_forward_proj {cfg.feed_forward_proj!r}: only gated-gelu (flan-T5) and relu (T5 v1.0) are built")
    if cfg.num_heads * cfg.d_kv % 8 != 0:
        raise NotImplementedError("T5 inner dim must be a multiple of 8")


def scale_decoder_outputs(cfg) -> bool:
    """transformers 4.33.1 scales the decoder output by d_model**-0.5 iff tie_word_embeddings;
    5.x froze that decision into config.scale_decoder_outputs (flan-T5: untied, unscaled)."""
    return bool(getattr(cfg, "scale_decoder_outputs", cfg.tie_word_embeddings))


def pack_t5(lm, cache: PackCache, need_backward: bool):
    params = list(lm.parameters())

    def ff(layer_ff):
        d = layer_ff.DenseReluDense
        gated = bool(lm.config.is_gated_act)
        wi = cat_bf16([d.wi_0.weight, d.wi_1.weight]) if gated else bf16(d.wi.weight)
        return dict(ln=f32(layer_ff.layer_norm.weight), wi_w=wi, wo_w=bf16(d.wo.weight), gated=gated)

    def att(a):
        return dict(qkv_w=cat_bf16([a.q.weight, a.k.weight, a.v.weight]), o_w=bf16(a.o.weight))

    def build():
        w = {"shared": bf16(lm.shared.weight), "head": bf16(lm.lm_head.weight), "enc": [], "dec": [],
             "enc_ln": f32(lm.encoder.final_layer_norm.weight), "dec_ln": f32(lm.decoder.final_layer_norm.weight),
             "enc_rel": f32(lm.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight),
             "dec_rel": f32(lm.decoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight)}
        for blk in lm.encoder.block:
            e = dict(ln1=f32(blk.layer[0].layer_norm.weight), **att(blk.layer[0].SelfAttention))
            e["ff"] = ff(blk.layer[1])
            w["enc"].append(e)
        ckv = []
        for blk in lm.decoder.block:
            ca = blk.layer[1].EncDecAttention
            d = dict(ln1=f32(blk.layer[0].layer_norm.weight), **att(blk.layer[0].SelfAttention),
                     ln2=f32(blk.layer[1].layer_norm.weight), cq_w=bf16(ca.q.weight), co_w=bf16(ca.o.weight))
            d["ff"] = ff(blk.layer[2])
            ckv += [ca.k.weight, ca.v.weight]
            w["dec"].append(d)
        w["ckv_w"] = cat_bf16(ckv)  # (layers * 2 * inner, d_model): every layer's cross K|V in one GEMM
        return w

    w = cache.get("t5", params, build)
    if need_backward and "head_t" not in w:
        with torch.no_grad():  # dgrad operands, packed once (the LM is frozen)
            w["head_t"] = T(w["head"])
            w["ckv_wt"] = T(w["ckv_w"])
            for lw in w["enc"] + w["dec"]:
                for k in ("qkv_w", "o_w", "cq_w", "co_w"):
                    if k in lw:
                        lw[k + "t"] = T(lw[k])
                for k in ("wi_w", "wo_w"):
                    lw["ff"][k + "t"] = T(lw["ff"][k])
    return w


_REL_CACHE: dict = {}


def rel_bias_table(weight: torch.Tensor, sq: int, skv: int, bidirectional: bool, cfg) -> torch.Tensor:
    """(heads, sq + skv - 1) f32: entry (j - i) + (sq - 1) = bias of key j seen from query i
    (T5Attention.compute_bias / _relative_position_bucket; the bias depends on j - i only).
    The bucket indices are integer bookkeeping computed once per shape ON THE HOST with the
    reference's own float32 expression: the log-spaced bucket edges (distance 16, 32, 64 ...)
    sit exactly on integers, where a device log() one ulp lower would pick the other bucket."""
    nb_all, maxd = int(cfg.relative_attention_num_buckets), int(cfg.relative_attention_max_distance)
    key = (sq, skv, bidirectional, nb_all, maxd, str(weight.device))
    bucket = _REL_CACHE.get(key)
    if bucket is None:
        rel = torch.arange(-(sq - 1), skv, dtype=torch.long)
        nb = nb_all
        ret = torch.zeros_like(rel)
        if bidirectional:
            nb //= 2
            ret = ret + (rel > 0).long() * nb
            rel = rel.abs()
        else:
            rel = -torch.minimum(rel, torch.zeros_like(rel))
        max_exact = nb // 2
        large = max_exact + (torch.log(rel.float() / max_exact) / math.log(maxd / max_exact) * (nb - max_exact)).long()
        large = torch.minimum(large, torch.full_like(large, nb - 1))
        bucket = (ret + torch.where(rel < max_exact, rel, large)).to(weight.device)
        _REL_CACHE[key] = bucket
    return weight[bucket].t().contiguous()


def shift_right(labels: torch.Tensor, cfg) -> torch.Tensor:
    """T5ForConditionalGeneration._shift_right (token bookkeeping)."""
    out = torch.empty_like(labels)
    out[:, 1:] = labels[:, :-1]
    out[:, 0] = cfg.decoder_start_token_id
    return out.masked_fill(out == -100, cfg.pad_token_id)


# Dropout sites (T5's own ``dropout_rate``, active in train mode exactly as in the reference — the frozen
# LM runs in train() under HF Trainer, eilev/model/v2.py:228-238 + scripts/general/train_v2.py:123-130):
#   per block   0 self-attention probabilities (T5Attention)       1 self-attention output (T5LayerSelfAttention)
#               2 feed-forward inner activation (T5Dense*ActDense)  3 feed-forward output (T5LayerFF)
#               4 cross-attention probabilities                     5 cross-attention output (T5LayerCrossAttention)
#   per stack   embeddings and the final-layer-norm output (T5Stack)
# Masks are counter hashes of (seed, salt, element index), regenerated in the backward pass.
_SALT_T5 = 1 << 18
_T5_DEC = 1 << 12          # decoder blocks
_T5_STACK = 1 << 13        # + 0 encoder embeddings, 1 encoder output, 2 decoder embeddings, 3 decoder output


def _drop(p: float, seed, salt: int):
    return (p, seed, _SALT_T5 + salt) if (seed is not None and p > 0.0) else None


def _masked(t, d):
    """dropout(t) with the mask of site d (forward activations and, with the same site, their gradients)."""
    return t if d is None else ops.dropout(t, d[0], d[1], d[2])


def _ff_fwd(x, lw, eps, d_inner=None, d_out=None):
    y, r = ops.rmsnorm(x, lw["ln"], eps, save_stats=True)
    if lw["gated"]:  # T5DenseGatedActDense: gelu_new(wi_0 y) * (wi_1 y)
        h01 = ops.gemm(y, lw["wi_w"])
        act = ops.gated_gelu(h01)
    else:  # T5DenseActDense: relu(wi y) in the GEMM epilogue; the output doubles as the backward mask
        h01 = act = ops.gemm(y, lw["wi_w"], epilogue=ops.EPI_RELU)
    out = ops.gemm(_masked(act, d_inner), lw["wo_w"], residual=x, dropout=d_out)
    return out, dict(x=x, r=r, h01=h01)


def _ff_bwd(dx, lw, s, d_inner=None, d_out=None):
    d_act = _masked(ops.gemm(_masked(dx, d_out), lw["wo_wt"]), d_inner)
    d_h01 = ops.gated_gelu_bwd(d_act, s["h01"]) if lw["gated"] else ops.act_bwd(d_act, s["h01"], ops.EPI_RELU)
    return ops.rmsnorm_bwd(ops.gemm(d_h01, lw["wi_wt"]), s["x"], lw["ln"], s["r"], dx_add=dx)


def t5_forward(lm, cache: PackCache, input_ids, attention_mask, video_mask, video_features, labels=None,
               decoder_input_ids=None, save: bool = False, seed=None):
    """Returns dict(logits (B, Ld, V), loss, encoder_last_hidden_state, status, ctx).
    seed (device uint64 tensor): train mode — T5's dropout_rate is applied at every site HF applies it."""
    cfg = lm.config
    _check_cfg(cfg)
    p = float(cfg.dropout_rate) if seed is not None else 0.0

    def d(salt):
        return _drop(p, seed, salt)

    w = pack_t5(lm, cache, need_backward=save)
    dm, heads, dkv = cfg.d_model, cfg.num_heads, cfg.d_kv
    inner = heads * dkv
    eps = float(cfg.layer_norm_epsilon)
    b, l = input_ids.shape
    rows = b * l
    emb, _, slot, _, status = ops.embed_splice(input_ids, attention_mask, video_mask, w["shared"],
                                               video_features, None, 0, want_hidden=False)
    key_mask = attention_mask.to(torch.uint8).contiguous()
    enc_bias = rel_bias_table(w["enc_rel"], l, l, True, cfg)
    x = _masked(emb.view(rows, dm), d(_T5_STACK + 0))
    enc_saved = []
    for li, lw in enumerate(w["enc"]):
        y, r1 = ops.rmsnorm(x, lw["ln1"], eps, save_stats=True)
        qkv = ops.gemm(y, lw["qkv_w"]).view(b, l, 3 * inner)
        o, lse = ops.attention(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], heads, 1.0,
                               key_mask=key_mask, need_lse=True, rel_bias=enc_bias, dropout=d(li * 8 + 0))
        x_mid = ops.gemm(o.view(rows, inner), lw["o_w"], residual=x, dropout=d(li * 8 + 1))
        x_out, sff = _ff_fwd(x_mid, lw["ff"], eps, d(li * 8 + 2), d(li * 8 + 3))
        if save:
            enc_saved.append(dict(x=x, r1=r1, qkv=qkv, o=o, lse=lse, ff=sff))
        x = x_out
    enc_out, r_enc = ops.rmsnorm(x, w["enc_ln"], eps, save_stats=True)
    enc_out = _masked(enc_out, d(_T5_STACK + 1))
    x_enc_last = x

    if decoder_input_ids is None:
        if labels is None:
            raise ValueError("You have to specify either decoder_input_ids or labels")
        decoder_input_ids = shift_right(labels, cfg)
    ld = decoder_input_ids.shape[1]
    rows_d = b * ld
    n_dec = len(w["dec"])
    xd = _masked(ops.embedding(decoder_input_ids, w["shared"]).view(rows_d, dm), d(_T5_STACK + 2))
    ckv = ops.gemm(enc_out, w["ckv_w"]).view(b, l, n_dec * 2 * inner)
    dec_bias = rel_bias_table(w["dec_rel"], ld, ld, False, cfg)
    dec_saved = []
    for li, lw in enumerate(w["dec"]):
        sl = _T5_DEC + li * 8
        y, r1 = ops.rmsnorm(xd, lw["ln1"], eps, save_stats=True)
        qkv = ops.gemm(y, lw["qkv_w"]).view(b, ld, 3 * inner)
        o, lse = ops.attention(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], heads, 1.0,
                               causal=True, need_lse=True, rel_bias=dec_bias, dropout=d(sl + 0))
        x1 = ops.gemm(o.view(rows_d, inner), lw["o_w"], residual=xd, dropout=d(sl + 1))
        y2, r2 = ops.rmsnorm(x1, lw["ln2"], eps, save_stats=True)
        cq = ops.gemm(y2, lw["cq_w"]).view(b, ld, inner)
        ck = ckv[:, :, (2 * li) * inner:(2 * li + 1) * inner]
        cv = ckv[:, :, (2 * li + 1) * inner:(2 * li + 2) * inner]
        co, clse = ops.attention(cq, ck, cv, heads, 1.0, key_mask=key_mask, need_lse=True, dropout=d(sl + 4))
        x2 = ops.gemm(co.view(rows_d, inner), lw["co_w"], residual=x1, dropout=d(sl + 5))
        x3, sff = _ff_fwd(x2, lw["ff"], eps, d(sl + 2), d(sl + 3))
        if save:
            dec_saved.append(dict(x=xd, r1=r1, qkv=qkv, o=o, lse=lse, x1=x1, r2=r2, cq=cq, co=co, clse=clse, ff=sff))
        xd = x3
    final, r_dec = ops.rmsnorm(xd, w["dec_ln"], eps, save_stats=True)
    final = _masked(final, d(_T5_STACK + 3))
    alpha = dm ** -0.5 if scale_decoder_outputs(cfg) else 1.0
    logits = ops.gemm(final, w["head"], alpha=alpha).view(b, ld, -1)
    out = dict(logits=logits, loss=None, status=status, ctx=None,
               encoder_last_hidden_state=enc_out.view(b, l, dm))
    if labels is not None:
        loss, row_lse, n_valid = ops.cross_entropy(logits, labels, shift=0)
        out["loss"] = loss
        if save:
            out["ctx"] = dict(enc=enc_saved, dec=dec_saved, x_enc_last=x_enc_last, r_enc=r_enc, x_dec_last=xd,
                              r_dec=r_dec, ckv=ckv, logits=logits, labels=labels, row_lse=row_lse,
                              n_valid=n_valid, slot=slot, key_mask=key_mask, enc_bias=enc_bias, dec_bias=dec_bias,
                              alpha=alpha, b=b, l=l, ld=ld, seed=seed, p=p,
                              n_features=0 if video_features is None else video_features.shape[0])
    return out


# --------------------------------------------------------------------------- generate
def t5_encode(lm, cache: PackCache, input_ids, attention_mask, video_mask, video_features) -> dict:
    """Encoder pass + the cross-attention K|V of every decoder layer (computed once per prompt;
    HF keeps them in the EncoderDecoderCache).  Returns the state consumed by t5_decode_logits."""
    cfg = lm.config
    _check_cfg(cfg)
    w = pack_t5(lm, cache, need_backward=False)
    dm, heads, dkv = cfg.d_model, cfg.num_heads, cfg.d_kv
    inner = heads * dkv
    eps = float(cfg.layer_norm_epsilon)
    b, l = input_ids.shape
    rows = b * l
    emb, _, _, _, status = ops.embed_splice(input_ids, attention_mask, video_mask, w["shared"],
                                            video_features, None, 0, want_hidden=False)
    key_mask = attention_mask.to(torch.uint8).contiguous()
    enc_bias = rel_bias_table(w["enc_rel"], l, l, True, cfg)
    x = emb.view(rows, dm)
    for lw in w["enc"]:
        y = ops.rmsnorm(x, lw["ln1"], eps)
        qkv = ops.gemm(y, lw["qkv_w"]).view(b, l, 3 * inner)
        o = ops.attention(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], heads, 1.0,
                          key_mask=key_mask, rel_bias=enc_bias)
        x_mid = ops.gemm(o.view(rows, inner), lw["o_w"], residual=x)
        x, _ = _ff_fwd(x_mid, lw["ff"], eps)
    enc_out = ops.rmsnorm(x, w["enc_ln"], eps)
    ckv = ops.gemm(enc_out, w["ckv_w"]).view(b, l, len(w["dec"]) * 2 * inner)
    # valid encoder tokens as [first, end) per row for the one-query cross-attention of decode
    # steps (padding at either end; a mask with holes falls back to the flash kernel)
    am = attention_mask.bool()
    n_valid = am.sum(dim=1)
    first = am.int().argmax(dim=1)
    idx = torch.arange(l, device=am.device)[None, :]
    contiguous = bool(((idx >= first[:, None]) & (idx < (first + n_valid)[:, None]) == am).all())
    return dict(ckv=ckv, key_mask=key_mask, status=status, b=b, l=l, enc_out=enc_out.view(b, l, dm),
                enc_first=first.to(torch.int32).contiguous(), enc_end=(first + n_valid).to(torch.int32).contiguous(),
                enc_contiguous=contiguous)


def t5_decode_logits(lm, cache: PackCache, enc: dict, decoder_input_ids: torch.Tensor) -> torch.Tensor:
    """Next-token logits f32 (B, V) after the decoder prefix (B, t), re-running the whole prefix
    (the path for more than 16 rows; up to 16 rows use the KV-cached t5_decode_step)."""
    cfg = lm.config
    w = pack_t5(lm, cache, need_backward=False)
    dm, heads, dkv = cfg.d_model, cfg.num_heads, cfg.d_kv
    inner = heads * dkv
    eps = float(cfg.layer_norm_epsilon)
    b, t = decoder_input_ids.shape
    rows_d = b * t
    ckv, key_mask = enc["ckv"], enc["key_mask"]
    xd = ops.embedding(decoder_input_ids, w["shared"]).view(rows_d, dm)
    dec_bias = rel_bias_table(w["dec_rel"], t, t, False, cfg)
    for li, lw in enumerate(w["dec"]):
        y = ops.rmsnorm(xd, lw["ln1"], eps)
        qkv = ops.gemm(y, lw["qkv_w"]).view(b, t, 3 * inner)
        o = ops.attention(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], heads, 1.0,
                          causal=True, rel_bias=dec_bias)
        x1 = ops.gemm(o.view(rows_d, inner), lw["o_w"], residual=xd)
        cq = ops.gemm(ops.rmsnorm(x1, lw["ln2"], eps), lw["cq_w"]).view(b, t, inner)
        co = ops.attention(cq, ckv[:, :, (2 * li) * inner:(2 * li + 1) * inner],
                           ckv[:, :, (2 * li + 1) * inner:(2 * li + 2) * inner], heads, 1.0, key_mask=key_mask)
        x2 = ops.gemm(co.view(rows_d, inner), lw["co_w"], residual=x1)
        xd, _ = _ff_fwd(x2, lw["ff"], eps)
    last = xd.view(b, t, dm)[:, -1, :].contiguous()
    final = ops.rmsnorm(last, w["dec_ln"], eps)
    alpha = dm ** -0.5 if scale_decoder_outputs(cfg) else 1.0
    return ops.gemm(final, w["head"], alpha=alpha, out_dtype=torch.float32)


def t5_decode_init(lm, cache: PackCache, enc: dict, max_new: int) -> dict:
    """State of the KV-cached decoder step: a paged self-attention cache per decoder layer
    (the cross-attention K|V already live in `enc`), the per-sequence length counter and the
    unidirectional relative-position table for up to max_new + 1 positions."""
    from .opt import PagedKV

    cfg = lm.config
    w = pack_t5(lm, cache, need_backward=False)
    heads, inner = cfg.num_heads, cfg.num_heads * cfg.d_kv
    b = enc["b"]
    dev = enc["ckv"].device
    tmax = max_new + 1
    kv = PagedKV(len(w["dec"]), b, tmax, inner, dev)
    l = enc["l"]
    csplits = max(1, min(8, (l + 127) // 128))
    return dict(kv=kv, tmax=tmax, bias=rel_bias_table(w["dec_rel"], tmax, tmax, False, cfg),
                csplits=csplits, seq_ids=torch.arange(b, dtype=torch.int32, device=dev),
                cws=torch.empty(b * heads * csplits * (cfg.d_kv + 2), dtype=torch.float32, device=dev),
                ccnt=torch.zeros(b * heads, dtype=torch.int32, device=dev),
                ctx_len=torch.zeros(b, dtype=torch.int32, device=dev),
                first_valid=torch.zeros(b, dtype=torch.int32, device=dev),
                ws=torch.empty(b * heads * (cfg.d_kv + 2), dtype=torch.float32, device=dev),
                cnt=torch.zeros(b * heads, dtype=torch.int32, device=dev))


def t5_decode_step(lm, cache: PackCache, enc: dict, st: dict, tokens: torch.Tensor) -> torch.Tensor:
    """One decoder token per sequence on the weight-streaming kernels: tokens (B,) int64 ->
    next-token logits f32 (B, V).  Self-attention reads / appends the paged cache with the
    relative bias of the newest position; cross-attention is a one-query pass over the encoder
    K|V.  Stream-ordered and allocation-stable (CUDA-graph capturable)."""
    cfg = lm.config
    w = pack_t5(lm, cache, need_backward=False)
    dm, heads, dkv = cfg.d_model, cfg.num_heads, cfg.d_kv
    inner = heads * dkv
    eps = float(cfg.layer_norm_epsilon)
    kv, ckv, key_mask = st["kv"], enc["ckv"], enc["key_mask"]
    b = tokens.shape[0]
    st["ctx_len"].add_(1)
    x = ops.embedding(tokens, w["shared"])
    for li, lw in enumerate(w["dec"]):
        qkv = ops.gemv(x, lw["qkv_w"], ln=(lw["ln1"], None, eps))  # RMSNorm fused into the staging
        o = ops.paged_decode_attention(qkv, kv.k[li], kv.v[li], kv.table, st["ctx_len"], st["first_valid"], heads,
                                       kv.page_size, 1.0, workspace=st["ws"], counters=st["cnt"], splits=1,
                                       rel_bias=st["bias"], rel_center=st["tmax"] - 1)
        x1 = ops.gemv(o, lw["o_w"], residual=x)
        cq = ops.gemv(x1, lw["cq_w"], ln=(lw["ln2"], None, eps))
        ck = ckv[:, :, (2 * li) * inner:(2 * li + 1) * inner]
        cv = ckv[:, :, (2 * li + 1) * inner:(2 * li + 2) * inner]
        if enc["enc_contiguous"]:
            co = ops.decode_cross_attention(cq, ck, cv, st["seq_ids"], enc["enc_end"], enc["enc_first"], heads, 1.0,
                                            workspace=st["cws"], counters=st["ccnt"], splits=st["csplits"])
        else:
            co = ops.attention(cq.view(b, 1, inner), ck, cv, heads, 1.0, key_mask=key_mask).view(b, inner)
        x2 = ops.gemv(co, lw["co_w"], residual=x1)
        ff = lw["ff"]
        if ff["gated"]:
            act = ops.gated_gelu(ops.gemv(x2, ff["wi_w"], ln=(ff["ln"], None, eps)))
        else:
            act = ops.gemv(x2, ff["wi_w"], epilogue=ops.EPI_RELU, ln=(ff["ln"], None, eps))
        x = ops.gemv(act, ff["wo_w"], residual=x2)
    alpha = dm ** -0.5 if scale_decoder_outputs(cfg) else 1.0
    return ops.gemv(x, w["head"], alpha=alpha, out_dtype=torch.float32, ln=(w["dec_ln"], None, eps))


def t5_backward(lm, cache: PackCache, ctx: dict, grad_loss: torch.Tensor | None):
    """dgrad-only backward: returns d(video_features) (n_features, d_model) bf16."""
    cfg = lm.config
    w = pack_t5(lm, cache, need_backward=True)
    heads, dkv = cfg.num_heads, cfg.d_kv
    inner = heads * dkv
    b, l, ld = ctx["b"], ctx["l"], ctx["ld"]
    rows, rows_d = b * l, b * ld
    gs = None
    if grad_loss is not None:
        gs = grad_loss.detach().to(torch.float32).reshape(()).contiguous()
    seed, p = ctx.get("seed"), ctx.get("p", 0.0)

    def d(salt):
        return _drop(p, seed, salt)

    dlogits = ops.cross_entropy_bwd(ctx["logits"], ctx["labels"], ctx["row_lse"], ctx["n_valid"], gs, shift=0)
    d_final = _masked(ops.gemm(dlogits, w["head_t"], alpha=ctx["alpha"]), d(_T5_STACK + 3))
    dx = ops.rmsnorm_bwd(d_final, ctx["x_dec_last"], w["dec_ln"], ctx["r_dec"])
    del dlogits, d_final
    d_ckv = torch.empty_like(ctx["ckv"])  # every layer writes its own K | V slice
    ckv = ctx["ckv"]
    for li in range(len(w["dec"]) - 1, -1, -1):
        lw, s = w["dec"][li], ctx["dec"][li]
        sl = _T5_DEC + li * 8
        d_x2 = _ff_bwd(dx, lw["ff"], s["ff"], d(sl + 2), d(sl + 3))
        d_co = ops.gemm(_masked(d_x2, d(sl + 5)), lw["co_wt"]).view(b, ld, inner)
        ck = ckv[:, :, (2 * li) * inner:(2 * li + 1) * inner]
        cv = ckv[:, :, (2 * li + 1) * inner:(2 * li + 2) * inner]
        dcq, _, _ = ops.attention_bwd(s["cq"], ck, cv, s["co"], s["clse"], d_co, heads, 1.0,
                                      key_mask=ctx["key_mask"],
                                      dk=d_ckv[:, :, (2 * li) * inner:(2 * li + 1) * inner],
                                      dv=d_ckv[:, :, (2 * li + 1) * inner:(2 * li + 2) * inner],
                                      dropout=d(sl + 4))
        d_x1 = ops.rmsnorm_bwd(ops.gemm(dcq.view(rows_d, inner), lw["cq_wt"]), s["x1"], lw["ln2"], s["r2"],
                               dx_add=d_x2)
        d_o = ops.gemm(_masked(d_x1, d(sl + 1)), lw["o_wt"]).view(b, ld, inner)
        qkv = s["qkv"]
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], s["o"], s["lse"],
                          d_o, heads, 1.0, causal=True, rel_bias=ctx["dec_bias"],
                          dq=dqkv[:, :, :inner], dk=dqkv[:, :, inner:2 * inner], dv=dqkv[:, :, 2 * inner:],
                          dropout=d(sl + 0))
        dx = ops.rmsnorm_bwd(ops.gemm(dqkv.view(rows_d, 3 * inner), lw["qkv_wt"]), s["x"], lw["ln1"], s["r1"],
                             dx_add=d_x1)
    # encoder output <- all cross-attention K / V projections at once
    d_enc = _masked(ops.gemm(d_ckv.view(rows, -1), w["ckv_wt"]), d(_T5_STACK + 1))
    dx = ops.rmsnorm_bwd(d_enc, ctx["x_enc_last"], w["enc_ln"], ctx["r_enc"])
    for li in range(len(w["enc"]) - 1, -1, -1):
        lw, s = w["enc"][li], ctx["enc"][li]
        d_mid = _ff_bwd(dx, lw["ff"], s["ff"], d(li * 8 + 2), d(li * 8 + 3))
        d_o = ops.gemm(_masked(d_mid, d(li * 8 + 1)), lw["o_wt"]).view(b, l, inner)
        qkv = s["qkv"]
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], s["o"], s["lse"],
                          d_o, heads, 1.0, key_mask=ctx["key_mask"], rel_bias=ctx["enc_bias"],
                          dq=dqkv[:, :, :inner], dk=dqkv[:, :, inner:2 * inner], dv=dqkv[:, :, 2 * inner:],
                          dropout=d(li * 8 + 0))
        dx = ops.rmsnorm_bwd(ops.gemm(dqkv.view(rows, 3 * inner), lw["qkv_wt"]), s["x"], lw["ln1"], s["r1"],
                             dx_add=d_mid)
    if ctx["n_features"] == 0:
        return None
    return ops.splice_bwd(_masked(dx, d(_T5_STACK + 0)), ctx["slot"], ctx["n_features"])
