"""Q-Former (+ language_projection) forward and backward on sm_100a kernels.

Restates Blip2QFormerModel.forward for the query-only path used by VideoBLIP
(HF:blip_2/modeling_blip_2.py:962-1036, layers :729-780, attention :579-633, output
blocks :644-648 / :686-689 / :700-704) and ``language_projection`` (eilev/model/v2.py:199-203).

B200-first choices:
  * the K/V projections of the image tokens of ALL cross-attention layers run as ONE GEMM
    [N*T*257, 1408] x [1408, n_cross*2*768] straight off the ViT output (88 % of the
    Q-Former's FLOPs in a single tcgen05 launch);
  * q/k/v of the self-attention are one fused GEMM; attention reads/writes the fused
    buffers in place through strides;
  * this is the only trainable block of the recipe (train_v2.py:123-130), so its backward
    is written by hand: dgrad and wgrad GEMMs on the same tcgen05 kernel (operands staged
    by a transpose kernel), flash-attention backward, LayerNorm backward with
    deterministic dgamma/dbeta.
Dropout (hidden_dropout_prob / attention_probs_dropout_prob, active in train mode exactly as
in the reference) uses counter-hash masks regenerated in the backward pass; nothing is stored.
"""
from __future__ import annotations

import math

import torch

from .. import ops
from .packing import PackCache, bf16, cat_bf16, cat_f32, f32

T = ops.transpose


def qformer_param_list(model):
    """Ordered (name, parameter) list: query_tokens, every qformer.* parameter, projection."""
    out = [("query_tokens", model.query_tokens)]
    out += [("qformer." + n, p) for n, p in model.qformer.named_parameters()]
    out += [("language_projection.weight", model.language_projection.weight),
            ("language_projection.bias", model.language_projection.bias)]
    return out


def _act(cfg):
    act = {"gelu": ops.EPI_GELU, "relu": ops.EPI_RELU}.get(cfg.hidden_act)
    if act is None:
        raise ValueError(f"unsupported qformer hidden_act {cfg.hidden_act!r}")
    return act


def pack_qformer(model, cache: PackCache):
    qf = model.qformer
    params = [p for _, p in qformer_param_list(model)]

    def build():
        w = {"ln0_g": f32(qf.layernorm.weight), "ln0_b": f32(qf.layernorm.bias), "layers": [],
             "proj_w": bf16(model.language_projection.weight),
             "proj_b": f32(model.language_projection.bias)}
        w["proj_wt"] = T(w["proj_w"])
        cross_w, cross_b = [], []
        for layer in qf.encoder.layer:
            sa = layer.attention
            lw = dict(
                qkv_w=cat_bf16([sa.attention.query.weight, sa.attention.key.weight, sa.attention.value.weight]),
                qkv_b=cat_f32([sa.attention.query.bias, sa.attention.key.bias, sa.attention.value.bias]),
                o_w=bf16(sa.output.dense.weight), o_b=f32(sa.output.dense.bias),
                ln1_g=f32(sa.output.LayerNorm.weight), ln1_b=f32(sa.output.LayerNorm.bias),
                i_w=bf16(layer.intermediate_query.dense.weight), i_b=f32(layer.intermediate_query.dense.bias),
                o2_w=bf16(layer.output_query.dense.weight), o2_b=f32(layer.output_query.dense.bias),
                ln3_g=f32(layer.output_query.LayerNorm.weight), ln3_b=f32(layer.output_query.LayerNorm.bias),
                cross=None,
            )
            for k in ("qkv_w", "o_w", "i_w", "o2_w"):
                lw[k + "t"] = T(lw[k])
            if getattr(layer, "has_cross_attention", False):
                ca = layer.crossattention
                lw["cross"] = len(cross_w) // 2
                lw.update(
                    cq_w=bf16(ca.attention.query.weight), cq_b=f32(ca.attention.query.bias),
                    co_w=bf16(ca.output.dense.weight), co_b=f32(ca.output.dense.bias),
                    ln2_g=f32(ca.output.LayerNorm.weight), ln2_b=f32(ca.output.LayerNorm.bias),
                )
                lw["cq_wt"] = T(lw["cq_w"])
                lw["co_wt"] = T(lw["co_w"])
                cross_w += [ca.attention.key.weight, ca.attention.value.weight]
                cross_b += [ca.attention.key.bias, ca.attention.value.bias]
            w["layers"].append(lw)
        if cross_w:
            w["ckv_w"] = cat_bf16(cross_w)  # (n_cross*2*Dq, Dv)
            w["ckv_b"] = cat_f32(cross_b)
        return w

    return cache.get("qformer", params, build)


_SALT_QF = 1 << 16  # salts of the Q-Former dropout sites: _SALT_QF + layer*8 + site


def _drop(p: float, seed, layer: int, site: int):
    return (p, seed, _SALT_QF + layer * 8 + site) if (seed is not None and p > 0.0) else None


def qformer_forward(model, cache: PackCache, image_embeds: torch.Tensor, save: bool, seed=None):
    """image_embeds (N, Skv, Dv) bf16 -> (video_features (N*Q, Dt) bf16, query_output (N, Q, Dq)
    bf16, ctx for backward | None).  seed: device int64 tensor -> dropout is applied."""
    cfg = model.qformer.config
    w = pack_qformer(model, cache)
    n, skv, dv = image_embeds.shape
    nq = model.query_tokens.shape[1]
    dq = cfg.hidden_size
    heads = cfg.num_attention_heads
    scale = 1.0 / math.sqrt(dq // heads)
    eps = cfg.layer_norm_eps
    act = _act(cfg)
    rows = n * nq
    img2 = image_embeds.reshape(n * skv, dv)

    p_h = float(cfg.hidden_dropout_prob) if seed is not None else 0.0
    p_a = float(cfg.attention_probs_dropout_prob) if seed is not None else 0.0
    x0 = bf16(model.query_tokens).reshape(1, nq, dq).expand(n, nq, dq).reshape(rows, dq).contiguous()
    x, m0, r0 = ops.layernorm(x0, w["ln0_g"], w["ln0_b"], eps, save_stats=True)
    if p_h > 0.0:
        x = ops.dropout(x, p_h, seed, _SALT_QF + 7)
    ckv = None
    if "ckv_w" in w:
        ckv = ops.gemm(img2, w["ckv_w"], w["ckv_b"]).view(n, skv, -1)
    saved = []
    for li, lw in enumerate(w["layers"]):
        s = {"x": x}
        qkv = ops.gemm(x, lw["qkv_w"], lw["qkv_b"]).view(n, nq, 3 * dq)
        ctx, lse = ops.attention(qkv[:, :, :dq], qkv[:, :, dq:2 * dq], qkv[:, :, 2 * dq:], heads, scale,
                                 need_lse=True, dropout=_drop(p_a, seed, li, 0))
        a = ops.gemm(ctx.view(rows, dq), lw["o_w"], lw["o_b"], residual=x,
                     dropout=_drop(p_h, seed, li, 1))  # dropout(dense(ctx)) + x
        x1, m1, r1 = ops.layernorm(a, lw["ln1_g"], lw["ln1_b"], eps, save_stats=True)
        s.update(qkv=qkv, ctx=ctx, lse=lse, xin1=a, m1=m1, r1=r1, x1=x1)
        if lw["cross"] is not None:
            c = lw["cross"]
            kc = ckv[:, :, (2 * c) * dq:(2 * c + 1) * dq]
            vc = ckv[:, :, (2 * c + 1) * dq:(2 * c + 2) * dq]
            qc = ops.gemm(x1, lw["cq_w"], lw["cq_b"]).view(n, nq, dq)
            cctx, clse = ops.attention(qc, kc, vc, heads, scale, need_lse=True, dropout=_drop(p_a, seed, li, 2))
            a2 = ops.gemm(cctx.view(rows, dq), lw["co_w"], lw["co_b"], residual=x1,
                          dropout=_drop(p_h, seed, li, 3))
            x2, m2, r2 = ops.layernorm(a2, lw["ln2_g"], lw["ln2_b"], eps, save_stats=True)
            s.update(qc=qc, cctx=cctx, clse=clse, xin2=a2, m2=m2, r2=r2)
        else:
            x2 = x1
        s["x2"] = x2
        inter = ops.gemm(x2, lw["i_w"], lw["i_b"], epilogue=act)
        a3 = ops.gemm(inter, lw["o2_w"], lw["o2_b"], residual=x2, dropout=_drop(p_h, seed, li, 4))
        x, m3, r3 = ops.layernorm(a3, lw["ln3_g"], lw["ln3_b"], eps, save_stats=True)
        s.update(inter=inter, xin3=a3, m3=m3, r3=r3)
        saved.append(s)
    feats = ops.gemm(x, w["proj_w"], w["proj_b"])
    ctx_out = None
    if save:
        ctx_out = dict(saved=saved, x0=x0, m0=m0, r0=r0, ckv=ckv, img2=img2, qout=x, n=n, nq=nq,
                       skv=skv, seed=seed, p_h=p_h, p_a=p_a)
    return feats, x.view(n, nq, dq), ctx_out


def _wgrad(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW (N_out, K_in) f32 = dy^T (N_out, M) . x (M, K_in)."""
    return ops.gemm_tn(dy, x, out_dtype=torch.float32)


class _GradOut(dict):
    """name -> gradient.  Parameters listed in `sink` (their pre-allocated f32 ``.grad`` views of
    the trainer's flat buffer) are accumulated IN PLACE by the producing kernel — wgrad GEMM with
    beta = 1, column sums / LayerNorm parameter gradients in accumulate mode — and are not
    returned, so autograd issues no ``grad += g`` kernel for them (257 small launches per
    micro-step otherwise)."""

    def __init__(self, sink: dict | None, device) -> None:
        super().__init__()
        self.sink = sink or {}
        self.device = device

    def weight(self, name: str, dy: torch.Tensor, x: torch.Tensor) -> None:
        dst = self.sink.get(name)
        if dst is not None:
            ops.gemm_tn(dy, x, out=dst, beta=1.0)
        else:
            self[name] = _wgrad(dy, x)

    def bias(self, name: str, dy: torch.Tensor) -> None:
        dst = self.sink.get(name)
        if dst is not None:
            ops.colsum(dy, out=dst)
        else:
            self[name] = ops.colsum(dy)

    def _fused_view(self, names):
        """One (sum of rows, cols) view over the sink gradients of `names` if they sit back to back in the trainer's
        flat buffer (query / key / value weights do: FlatBuffers keeps the parameter order inside its decay and
        no-decay segments), else None."""
        dsts = [self.sink.get(n) for n in names]
        if any(d is None or not d.is_contiguous() for d in dsts):
            return None
        for a, b in zip(dsts, dsts[1:]):
            if b.data_ptr() != a.data_ptr() + a.numel() * a.element_size() or b.shape[1:] != a.shape[1:]:
                return None
        first = dsts[0]
        rows = sum(d.shape[0] for d in dsts)
        if first.dim() == 1:
            return torch.as_strided(first, (rows,), (1,))
        return torch.as_strided(first, (rows, first.shape[1]), (first.stride(0), 1))

    def weights_fused(self, names, dy: torch.Tensor, x: torch.Tensor) -> bool:
        """The weight gradients of a fused projection (dy: (tokens, sum of N_out)) straight into their adjacent sink
        views with ONE accumulating GEMM; False (nothing done) when the views are not adjacent."""
        out = self._fused_view(names)
        if out is None or out.shape[0] != dy.shape[1]:
            return False
        ops.gemm_tn(dy, x, out=out, beta=1.0)
        return True

    def biases_fused(self, names, dy: torch.Tensor) -> bool:
        out = self._fused_view(names)
        if out is None or out.shape[0] != dy.shape[1]:
            return False
        ops.colsum(dy, out=out)
        return True

    def ln(self, wname: str, bname: str, cols: int):
        """(dgamma, dbeta) buffers for vb_layernorm_bwd, which accumulates into them."""
        dg, db = self.sink.get(wname), self.sink.get(bname)
        if dg is not None and db is not None:
            return dg, db
        dg = torch.zeros(cols, dtype=torch.float32, device=self.device)
        db = torch.zeros(cols, dtype=torch.float32, device=self.device)
        self[wname], self[bname] = dg, db
        return dg, db


def qformer_backward(model, cache: PackCache, ctx: dict, d_feats: torch.Tensor, sink: dict | None = None) -> dict:
    """Returns {parameter name: f32 gradient} for the entries of qformer_param_list that were
    not accumulated straight into `sink` (see _GradOut)."""
    cfg = model.qformer.config
    w = pack_qformer(model, cache)
    n, nq, skv = ctx["n"], ctx["nq"], ctx["skv"]
    dq = cfg.hidden_size
    heads = cfg.num_attention_heads
    scale = 1.0 / math.sqrt(dq // heads)
    act = _act(cfg)
    rows = n * nq
    dev = d_feats.device
    g = _GradOut(sink, dev)

    seed, p_h, p_a = ctx["seed"], ctx["p_h"], ctx["p_a"]

    def masked(t, layer, site):
        """gradient entering a dense layer whose OUTPUT was dropped in the forward"""
        d = _drop(p_h, seed, layer, site)
        return t if d is None else ops.dropout(t, d[0], d[1], d[2])

    d_feats = d_feats.contiguous()
    g.weight("language_projection.weight", d_feats, ctx["qout"])
    g.bias("language_projection.bias", d_feats)
    dx = ops.gemm(d_feats, w["proj_wt"])

    # every cross layer's attention backward WRITES its dK / dV slice (plain stores, no accumulation) and the slices
    # tile the buffer: no 644 MB zero fill per micro-step
    d_ckv = torch.empty_like(ctx["ckv"]) if ctx["ckv"] is not None else None
    n_layers = len(w["layers"])
    for i in range(n_layers - 1, -1, -1):
        lw, s = w["layers"][i], ctx["saved"][i]
        p = f"qformer.encoder.layer.{i}."
        # ---- output_query: x = LN3(inter W2^T + b2 + x2)
        dg, db = g.ln(p + "output_query.LayerNorm.weight", p + "output_query.LayerNorm.bias", dq)
        ds = ops.layernorm_bwd(dx, s["xin3"], lw["ln3_g"], s["m3"], s["r3"], dgamma=dg, dbeta=db)
        dsm = masked(ds, i, 4)
        g.weight(p + "output_query.dense.weight", dsm, s["inter"])
        g.bias(p + "output_query.dense.bias", dsm)
        # dgrad of output_query.dense with the activation's backward in its epilogue
        saved_act = ops.gemm(s["x2"], lw["i_w"], lw["i_b"]) if act == ops.EPI_GELU else s["inter"]  # GELU: recompute
        d_pre = ops.gemm_act_bwd(dsm, lw["o2_wt"], saved_act, act)
        g.weight(p + "intermediate_query.dense.weight", d_pre, s["x2"])
        g.bias(p + "intermediate_query.dense.bias", d_pre)
        dx2 = ops.gemm(d_pre, lw["i_wt"], residual=ds)
        # ---- cross attention: x2 = LN2(cctx Wo^T + bo + x1)
        if lw["cross"] is not None:
            c = lw["cross"]
            dg, db = g.ln(p + "crossattention.output.LayerNorm.weight", p + "crossattention.output.LayerNorm.bias", dq)
            ds2 = ops.layernorm_bwd(dx2, s["xin2"], lw["ln2_g"], s["m2"], s["r2"], dgamma=dg, dbeta=db)
            ds2m = masked(ds2, i, 3)
            g.weight(p + "crossattention.output.dense.weight", ds2m, s["cctx"].view(rows, dq))
            g.bias(p + "crossattention.output.dense.bias", ds2m)
            d_cctx = ops.gemm(ds2m, lw["co_wt"]).view(n, nq, dq)
            kc = ctx["ckv"][:, :, (2 * c) * dq:(2 * c + 1) * dq]
            vc = ctx["ckv"][:, :, (2 * c + 1) * dq:(2 * c + 2) * dq]
            dkc = d_ckv[:, :, (2 * c) * dq:(2 * c + 1) * dq]
            dvc = d_ckv[:, :, (2 * c + 1) * dq:(2 * c + 2) * dq]
            dqc, _, _ = ops.attention_bwd(s["qc"], kc, vc, s["cctx"], s["clse"], d_cctx, heads, scale,
                                          dk=dkc, dv=dvc, dropout=_drop(p_a, seed, i, 2))
            dqc2 = dqc.view(rows, dq)
            g.weight(p + "crossattention.attention.query.weight", dqc2, s["x1"])
            g.bias(p + "crossattention.attention.query.bias", dqc2)
            dx1 = ops.gemm(dqc2, lw["cq_wt"], residual=ds2)
        else:
            dx1 = dx2
        # ---- self attention: x1 = LN1(ctx Wo^T + bo + x)
        dg, db = g.ln(p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias", dq)
        ds1 = ops.layernorm_bwd(dx1, s["xin1"], lw["ln1_g"], s["m1"], s["r1"], dgamma=dg, dbeta=db)
        ds1m = masked(ds1, i, 1)
        g.weight(p + "attention.output.dense.weight", ds1m, s["ctx"].view(rows, dq))
        g.bias(p + "attention.output.dense.bias", ds1m)
        d_ctx = ops.gemm(ds1m, lw["o_wt"]).view(n, nq, dq)
        qkv = s["qkv"]
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :, :dq], qkv[:, :, dq:2 * dq], qkv[:, :, 2 * dq:], s["ctx"], s["lse"],
                          d_ctx, heads, scale, dq=dqkv[:, :, :dq], dk=dqkv[:, :, dq:2 * dq],
                          dv=dqkv[:, :, 2 * dq:], dropout=_drop(p_a, seed, i, 0))
        dqkv2 = dqkv.view(rows, 3 * dq)
        qkv_names = [p + f"attention.attention.{nm}" for nm in ("query", "key", "value")]
        if not g.weights_fused([nm + ".weight" for nm in qkv_names], dqkv2, s["x"]):
            dw = _wgrad(dqkv2, s["x"])
            for j, nm in enumerate(qkv_names):
                g[nm + ".weight"] = dw[j * dq:(j + 1) * dq]
        if not g.biases_fused([nm + ".bias" for nm in qkv_names], dqkv2):
            dbias = ops.colsum(dqkv2)
            for j, nm in enumerate(qkv_names):
                g[nm + ".bias"] = dbias[j * dq:(j + 1) * dq]
        dx = ops.gemm(dqkv2, lw["qkv_wt"], residual=ds1)

    # ---- cross K/V projections of every cross layer in one wgrad GEMM
    if d_ckv is not None:
        d_ckv2 = d_ckv.view(n * skv, -1)
        dw = ops.gemm_tn(d_ckv2, ctx["img2"], out_dtype=torch.float32)  # (n_cross*2*Dq, Dv)
        dbias = ops.colsum(d_ckv2)
        for i, lw in enumerate(w["layers"]):
            if lw["cross"] is None:
                continue
            c = lw["cross"]
            p = f"qformer.encoder.layer.{i}.crossattention.attention."
            g[p + "key.weight"] = dw[(2 * c) * dq:(2 * c + 1) * dq]
            g[p + "key.bias"] = dbias[(2 * c) * dq:(2 * c + 1) * dq]
            g[p + "value.weight"] = dw[(2 * c + 1) * dq:(2 * c + 2) * dq]
            g[p + "value.bias"] = dbias[(2 * c + 1) * dq:(2 * c + 2) * dq]

    # ---- initial LayerNorm over the expanded query tokens
    dg, db = g.ln("qformer.layernorm.weight", "qformer.layernorm.bias", dq)
    if p_h > 0.0:
        dx = ops.dropout(dx, p_h, seed, _SALT_QF + 7)
    d0 = ops.layernorm_bwd(dx, ctx["x0"], w["ln0_g"], ctx["m0"], ctx["r0"], dgamma=dg, dbeta=db)
    g["query_tokens"] = ops.colsum(d0.view(n, nq * dq)).view(1, nq, dq)
    return g
