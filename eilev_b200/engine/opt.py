"""OPT causal LM over the interleaved (32 video queries per clip + text) sequence.

Restates OPTForCausalLM.forward (HF:opt/modeling_opt.py:464-524), OPTDecoder.forward
(:321-396), OPTDecoderLayer (:202-253), OPTAttention (:135-181; q is scaled by
head_dim**-0.5 right after q_proj and the softmax scale is 1.0) and the shifted causal-LM
loss (HF:loss/loss_utils.py:28-67) as sm_100a launches:

    embed_splice (token gather + video-feature splice + learned positions, offset 2)
    32 x [LN, fused QKV GEMM (alpha on the q columns), causal flash attention with the
          padding mask, out_proj GEMM(+residual), LN, fc1 GEMM(+ReLU), fc2 GEMM(+residual)]
    final LN, tied lm_head GEMM, fused cross entropy.

The LM is frozen in the reference recipe (train_v2.py:126-127): its backward is dgrad
only — gradients flow through activations to the spliced video positions, no wgrad.
Decode (``generate``) runs token-by-token on weight-streaming GEMV kernels over a paged KV
cache.  Dropout (config.dropout after out_proj and fc2, config.attention_dropout on the
probabilities; active in train mode as in the reference — the frozen LM is still in
train mode under HF Trainer) uses counter-hash masks regenerated in the backward pass.
"""
from __future__ import annotations

import os

import torch

from .. import ops
from .packing import PackCache, bf16, cat_bf16, cat_f32, f32

T = ops.transpose


def _check_cfg(cfg) -> None:
    if not cfg.do_layer_norm_before:
        raise NotImplementedError("OPT with do_layer_norm_before=False (opt-350m) is not supported")
    if cfg.word_embed_proj_dim != cfg.hidden_size:
        raise NotImplementedError("OPT with word_embed_proj_dim != hidden_size is not supported")
    if getattr(cfg, "_remove_final_layer_norm", False):
        raise NotImplementedError("OPT without the final layer norm is not supported")
    if cfg.activation_function not in ("relu", "gelu"):
        raise NotImplementedError(f"OPT activation {cfg.activation_function!r} is not supported")


def pack_opt(lm, cache: PackCache, need_backward: bool):
    dec = lm.model.decoder
    params = list(lm.parameters())

    def build():
        w = {
            "embed": bf16(dec.embed_tokens.weight),
            "pos": bf16(dec.embed_positions.weight),
            "lnf_g": f32(dec.final_layer_norm.weight), "lnf_b": f32(dec.final_layer_norm.bias),
            "layers": [],
        }
        for layer in dec.layers:
            sa = layer.self_attn
            has_b = sa.q_proj.bias is not None
            w["layers"].append(dict(
                ln1_g=f32(layer.self_attn_layer_norm.weight), ln1_b=f32(layer.self_attn_layer_norm.bias),
                qkv_w=cat_bf16([sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight]),
                qkv_b=cat_f32([sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias]) if has_b else None,
                out_w=bf16(sa.out_proj.weight), out_b=f32(sa.out_proj.bias) if has_b else None,
                ln2_g=f32(layer.final_layer_norm.weight), ln2_b=f32(layer.final_layer_norm.bias),
                fc1_w=bf16(layer.fc1.weight), fc1_b=f32(layer.fc1.bias) if layer.fc1.bias is not None else None,
                fc2_w=bf16(layer.fc2.weight), fc2_b=f32(layer.fc2.bias) if layer.fc2.bias is not None else None,
            ))
        return w

    w = cache.get("opt", params, build)
    if need_backward and "embed_t" not in w:
        with torch.no_grad():  # dgrad operands: W^T, packed once (the LM is frozen)
            w["embed_t"] = T(w["embed"])
            for lw in w["layers"]:
                for k in ("qkv_w", "out_w", "fc1_w", "fc2_w"):
                    lw[k + "t"] = T(lw[k])
    return w


def _dims(cfg):
    heads = cfg.num_attention_heads
    return cfg.hidden_size, heads, cfg.hidden_size // heads


_SALT_OPT = 1 << 17


def _drop(p: float, seed, layer: int, site: int):
    return (p, seed, _SALT_OPT + layer * 8 + site) if (seed is not None and p > 0.0) else None


def opt_forward(lm, cache: PackCache, input_ids, attention_mask, video_mask, video_features,
                labels=None, save: bool = False, want_logits: bool = True,
                output_hidden_states: bool = False, kv_sink=None, seed=None):
    """Full-sequence forward.  Returns dict(inputs_embeds, logits, loss, hidden_states, ctx)."""
    cfg = lm.config
    _check_cfg(cfg)
    w = pack_opt(lm, cache, need_backward=save)
    dim, heads, hd = _dims(cfg)
    act = ops.EPI_RELU if cfg.activation_function == "relu" else ops.EPI_GELU
    scaling = hd ** -0.5
    b, l = input_ids.shape
    rows = b * l
    p_h = float(cfg.dropout) if seed is not None else 0.0
    p_a = float(cfg.attention_dropout) if seed is not None else 0.0
    if seed is not None and float(getattr(cfg, "layerdrop", 0.0)) > 0.0:
        raise NotImplementedError("OPT layerdrop > 0 is not supported")
    emb, hidden, slot, pos_ids, status = ops.embed_splice(
        input_ids, attention_mask, video_mask, w["embed"], video_features, w["pos"], 2)
    key_mask = None
    if attention_mask is not None:
        key_mask = attention_mask.to(torch.uint8).contiguous()
    x = hidden.view(rows, dim)
    saved = []
    all_hidden = [hidden] if output_hidden_states else None
    for li, lw in enumerate(w["layers"]):
        y, m1, r1 = ops.layernorm(x, lw["ln1_g"], lw["ln1_b"], 1e-5, save_stats=True)
        qkv = ops.gemm(y, lw["qkv_w"], lw["qkv_b"], alpha=scaling, alpha_cols=dim).view(b, l, 3 * dim)
        q, k, v = qkv[:, :, :dim], qkv[:, :, dim:2 * dim], qkv[:, :, 2 * dim:]
        if kv_sink is not None:
            kv_sink(li, k, v)
        o, lse = ops.attention(q, k, v, heads, 1.0, causal=True, key_mask=key_mask, need_lse=True,
                               dropout=_drop(p_a, seed, li, 0))
        x_mid = ops.gemm(o.view(rows, dim), lw["out_w"], lw["out_b"], residual=x, dropout=_drop(p_h, seed, li, 1))
        y2, m2, r2 = ops.layernorm(x_mid, lw["ln2_g"], lw["ln2_b"], 1e-5, save_stats=True)
        f1 = ops.gemm(y2, lw["fc1_w"], lw["fc1_b"], epilogue=act)
        x_out = ops.gemm(f1, lw["fc2_w"], lw["fc2_b"], residual=x_mid, dropout=_drop(p_h, seed, li, 2))
        if save:
            s = dict(x_in=x, m1=m1, r1=r1, qkv=qkv, o=o, lse=lse, x_mid=x_mid, m2=m2, r2=r2, f1=f1)
            if act == ops.EPI_GELU:
                s["y2"] = y2
            saved.append(s)
        x = x_out
        if output_hidden_states:
            all_hidden.append(x.view(b, l, dim))
    final, mf, rf = ops.layernorm(x, w["lnf_g"], w["lnf_b"], 1e-5, save_stats=True)
    if output_hidden_states:
        all_hidden[-1] = final.view(b, l, dim)  # HF reports the normalised last state
    out = dict(inputs_embeds=emb, status=status, hidden_states=all_hidden, final=final.view(b, l, dim),
               logits=None, loss=None, ctx=None)
    if want_logits or labels is not None:
        logits = ops.gemm(final, w["embed"]).view(b, l, -1)  # tied lm_head, no bias
        out["logits"] = logits
        if labels is not None:
            loss, row_lse, n_valid = ops.cross_entropy(logits, labels)
            out["loss"] = loss
            if save:
                out["ctx"] = dict(saved=saved, x_last=x, mf=mf, rf=rf, logits=logits, labels=labels,
                                  row_lse=row_lse, n_valid=n_valid, slot=slot, key_mask=key_mask,
                                  seed=seed, p_h=p_h, p_a=p_a,
                                  b=b, l=l, n_features=0 if video_features is None else video_features.shape[0])
    return out


def opt_backward(lm, cache: PackCache, ctx: dict, grad_loss: torch.Tensor | None):
    """dgrad-only backward: returns d(video_features) (n_features, D) bf16."""
    cfg = lm.config
    w = pack_opt(lm, cache, need_backward=True)
    dim, heads, hd = _dims(cfg)
    act = ops.EPI_RELU if cfg.activation_function == "relu" else ops.EPI_GELU
    scaling = hd ** -0.5
    b, l = ctx["b"], ctx["l"]
    rows = b * l
    gs = None
    if grad_loss is not None:
        gs = grad_loss.detach().to(torch.float32).reshape(()).contiguous()
    dlogits = ops.cross_entropy_bwd(ctx["logits"], ctx["labels"], ctx["row_lse"], ctx["n_valid"], gs)
    d_final = ops.gemm(dlogits, w["embed_t"])
    seed, p_h, p_a = ctx["seed"], ctx["p_h"], ctx["p_a"]
    n_layers = len(w["layers"])

    def ln_bwd(dy, xin, gamma, mean, rstd, dx_add, layer, site):
        """dx of a LayerNorm and, in train mode, the same gradient through the residual dropout of (layer, site):
        the input of the next dgrad GEMM, written by the same kernel."""
        d = _drop(p_h, seed, layer, site) if layer >= 0 else None
        if d is None:
            dx_ = ops.layernorm_bwd(dy, xin, gamma, mean, rstd, dx_add=dx_add)
            return dx_, dx_
        if os.environ.get("VB_OPT_FUSED_LN_DROP", "1") == "0":  # measurement knob: the two-kernel form
            dx_ = ops.layernorm_bwd(dy, xin, gamma, mean, rstd, dx_add=dx_add)
            return dx_, ops.dropout(dx_, d[0], d[1], d[2])
        return ops.layernorm_bwd(dy, xin, gamma, mean, rstd, dx_add=dx_add, dropout=d)

    # gradient of the residual stream entering the top layer, and its copy through fc2's dropout (site 2)
    dx, dx_m = ln_bwd(d_final, ctx["x_last"], w["lnf_g"], ctx["mf"], ctx["rf"], None, n_layers - 1, 2)
    del dlogits, d_final

    for li in range(n_layers - 1, -1, -1):
        lw, s = w["layers"][li], ctx["saved"][li]
        # fc2 dgrad with the activation's backward in its epilogue (ReLU: mask by the saved output; GELU: the
        # pre-activation is recomputed)
        saved_act = s["f1"] if act == ops.EPI_RELU else ops.gemm(s["y2"], lw["fc1_w"], lw["fc1_b"])
        d_pre = ops.gemm_act_bwd(dx_m, lw["fc2_wt"], saved_act, act)
        d_y2 = ops.gemm(d_pre, lw["fc1_wt"])
        # ... and through the attention output projection's dropout (site 1)
        d_mid, d_mid_m = ln_bwd(d_y2, s["x_mid"], lw["ln2_g"], s["m2"], s["r2"], dx, li, 1)
        d_o = ops.gemm(d_mid_m, lw["out_wt"]).view(b, l, dim)
        qkv = s["qkv"]
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :, :dim], qkv[:, :, dim:2 * dim], qkv[:, :, 2 * dim:], s["o"], s["lse"],
                          d_o, heads, 1.0, causal=True, key_mask=ctx["key_mask"], dq_scale=scaling,
                          dq=dqkv[:, :, :dim], dk=dqkv[:, :, dim:2 * dim], dv=dqkv[:, :, 2 * dim:],
                          dropout=_drop(p_a, seed, li, 0))
        d_y = ops.gemm(dqkv.view(rows, 3 * dim), lw["qkv_wt"])
        # the layer below: its fc2 dropout (site 2)
        dx, dx_m = ln_bwd(d_y, s["x_in"], lw["ln1_g"], s["m1"], s["r1"], d_mid, li - 1, 2)
    if ctx["n_features"] == 0:
        return None
    return ops.splice_bwd(dx, ctx["slot"], ctx["n_features"])


# --------------------------------------------------------------------------- classify
def opt_classify(lm, cache: PackCache, prompt_ids, prompt_mask, video_mask, video_features,
                 class_ids, class_mask=None, class_batch_size: int | None = None) -> torch.Tensor:
    """Mean log-likelihood of every class continuation given the (left-padded) prompt:
    eilev/model/v2.py:326-501 (``classify`` step 4 + ``_calc_class_log_likelihood``).

    The prompt runs once; its per-layer K/V stay where the fused QKV GEMM wrote them.  Class
    tokens of all (sequence, class) pairs form one (B*C*Lc)-row batch; in every layer they
    attend (i) to the shared prompt K/V — one non-causal launch per layer with all C*Lc
    queries of a sequence stacked, instead of the reference's ``repeat_interleave`` of the
    cache (:457-460) — and (ii) causally to their own continuation; the two partial
    softmaxes are merged by their log-sum-exps.  Returns (B, num_classes) f32."""
    cfg = lm.config
    _check_cfg(cfg)
    w = pack_opt(lm, cache, need_backward=False)
    dim, heads, hd = _dims(cfg)
    act = ops.EPI_RELU if cfg.activation_function == "relu" else ops.EPI_GELU
    scaling = hd ** -0.5
    b, lp = prompt_ids.shape
    n_cls, lc = class_ids.shape
    dev = prompt_ids.device
    if prompt_mask is None:
        prompt_mask = torch.ones_like(prompt_ids)
    if class_mask is None:
        class_mask = torch.ones_like(class_ids)
    class_ids = class_ids.to(dev).contiguous()
    class_mask = class_mask.to(dev).to(torch.long).contiguous()

    prompt_kv: list = []
    out = opt_forward(lm, cache, prompt_ids, prompt_mask, video_mask, video_features, want_logits=False,
                      kv_sink=lambda li, k, v: prompt_kv.append((k, v)))
    status = out["status"]
    last = out["final"][:, -1, :].contiguous()
    prompt_logits = (ops.gemv(last, w["embed"], out_dtype=torch.float32) if b <= 16
                     else ops.gemm(last, w["embed"], out_dtype=torch.float32))  # (B, V)
    key_mask = prompt_mask.to(torch.uint8).contiguous()
    n_valid = prompt_mask.sum(dim=1).tolist()  # positions of the continuation start here

    # labels: class token j is scored by the logits of position j-1 (the prompt's last
    # position for j == 0); padded class tokens are ignored (v2.py:474-494)
    labels = torch.where(class_mask != 0, class_ids, torch.full_like(class_ids, -100))
    step = n_cls if class_batch_size is None else int(class_batch_size)
    chunks = []
    for c0 in range(0, n_cls, step):
        ids_c, mask_c, lab_c = class_ids[c0:c0 + step], class_mask[c0:c0 + step], labels[c0:c0 + step]
        ncc = ids_c.shape[0]
        rows = b * ncc * lc
        hidden = torch.cat([
            ops.embed_splice(ids_c, mask_c, None, w["embed"], None, w["pos"], 2 + int(n_valid[bi]),
                             want_embeds=False)[1].view(1, ncc * lc, dim)
            for bi in range(b)], dim=0)
        x = hidden.view(rows, dim)
        for li, lw in enumerate(w["layers"]):
            y = ops.layernorm(x, lw["ln1_g"], lw["ln1_b"], 1e-5)
            qkv = ops.gemm(y, lw["qkv_w"], lw["qkv_b"], alpha=scaling, alpha_cols=dim)
            pk, pv = prompt_kv[li]
            q_all = qkv.view(b, ncc * lc, 3 * dim)[:, :, :dim]
            o1, lse1 = ops.attention(q_all, pk, pv, heads, 1.0, causal=False, key_mask=key_mask, need_lse=True)
            qc = qkv.view(b * ncc, lc, 3 * dim)
            o2, lse2 = ops.attention(qc[:, :, :dim], qc[:, :, dim:2 * dim], qc[:, :, 2 * dim:], heads, 1.0,
                                     causal=True, need_lse=True)
            o = ops.attention_merge(o1, lse1, o2, lse2, heads)
            x_mid = ops.gemm(o.view(rows, dim), lw["out_w"], lw["out_b"], residual=x)
            y2 = ops.layernorm(x_mid, lw["ln2_g"], lw["ln2_b"], 1e-5)
            f1 = ops.gemm(y2, lw["fc1_w"], lw["fc1_b"], epilogue=act)
            x = ops.gemm(f1, lw["fc2_w"], lw["fc2_b"], residual=x_mid)
        final = ops.layernorm(x, w["lnf_g"], w["lnf_b"], 1e-5)
        logits = ops.gemm(final, w["embed"], out_dtype=torch.float32)  # (rows, V), tied head
        # row (b, c, j) scores class token j+1
        nxt = torch.cat([lab_c[:, 1:], torch.full_like(lab_c[:, :1], -100)], dim=1)
        lp_rest = ops.token_logprob(logits, nxt.unsqueeze(0).expand(b, -1, -1).contiguous())
        first_rows = torch.arange(b, device=dev, dtype=torch.long).repeat_interleave(ncc)
        lp_first = ops.token_logprob(prompt_logits, lab_c[:, 0].repeat(b).contiguous(), first_rows)
        total = lp_rest.view(b, ncc, lc).sum(dim=-1) + lp_first.view(b, ncc)
        lengths = mask_c.sum(dim=-1).to(torch.float32).unsqueeze(0)
        chunks.append(total / lengths)
    res = torch.cat(chunks, dim=1)
    return res, status


# --------------------------------------------------------------------------- decode
class PagedKV:
    """Paged KV cache: per layer K and V pools of (n_pages, page_size, H*D) bf16 and one page
    table (B, max_pages) shared by all layers.  Beam reordering permutes table rows and copies
    only partially filled tail pages."""

    def __init__(self, n_layers: int, batch: int, max_len: int, hd: int, device, page_size: int = 64):
        self.page_size = page_size
        self.max_pages = (max_len + page_size - 1) // page_size
        n_pages = batch * self.max_pages
        self.k = [torch.empty((n_pages, page_size, hd), dtype=torch.bfloat16, device=device) for _ in range(n_layers)]
        self.v = [torch.empty((n_pages, page_size, hd), dtype=torch.bfloat16, device=device) for _ in range(n_layers)]
        self.table = torch.arange(n_pages, dtype=torch.int32, device=device).view(batch, self.max_pages).contiguous()
        self.batch = batch

    def reorder(self, beam_idx: torch.Tensor) -> None:
        """Row i of the cache becomes old row beam_idx[i] (physical copy of the page contents;
        pages stay owned by their row so no reference counting is needed)."""
        idx = beam_idx.to(torch.long)
        if bool((idx == torch.arange(idx.numel(), device=idx.device)).all()):
            return
        mp = self.max_pages
        for pool in (self.k, self.v):
            for li in range(len(pool)):
                view = pool[li].view(self.batch, mp, self.page_size, -1)
                pool[li].copy_(view[idx].reshape(pool[li].shape))  # in place: pointers stay valid


_STATE_PAGES_ROUND = 4   # reusable decode states round their capacity up to 4 pages (256 tokens)
_STATE_CACHE_MAX = 2     # generation states kept alive per LM (each owns its KV pools)


def opt_prefill(lm, cache: PackCache, input_ids, attention_mask, video_mask, video_features,
                max_new_tokens: int, reuse_slot: int | None = None):
    """Runs the prompt, fills a paged KV cache, returns (last-position logits f32 (B, V), state).

    reuse_slot: generate() passes an integer so that consecutive calls with the same batch and (rounded)
    capacity write into the SAME KV pools / counters; the CUDA graph of the decode step captured for that
    state (``state["_graph"]``) then survives from one generate() call to the next instead of being
    re-captured (~20 ms) for every prompt.  Steppers alive at the same time use different slots."""
    cfg = lm.config
    dim, heads, hd = _dims(cfg)
    b, l = input_ids.shape
    prev = None
    capacity = l + max_new_tokens
    if reuse_slot is not None:
        pages = (capacity + 63) // 64
        pages = (pages + _STATE_PAGES_ROUND - 1) // _STATE_PAGES_ROUND * _STATE_PAGES_ROUND
        capacity = pages * 64
        states = lm.__dict__.setdefault("_decode_states", {})
        key = (int(reuse_slot), b, pages, str(input_ids.device))
        prev = states.get(key)
    kv = prev["kv"] if prev is not None else PagedKV(cfg.num_hidden_layers, b, capacity, dim, input_ids.device)

    def sink(li, k, v):
        ops.paged_kv_write(k, v, kv.k[li], kv.v[li], kv.table, kv.page_size)

    out = opt_forward(lm, cache, input_ids, attention_mask, video_mask, video_features,
                      want_logits=False, kv_sink=sink)
    w = pack_opt(lm, cache, need_backward=False)
    last = out["final"][:, -1, :].contiguous()
    logits = ops.gemv(last, w["embed"], out_dtype=torch.float32) if b <= 16 else \
        ops.gemm(last, w["embed"], out_dtype=torch.float32)
    am = attention_mask if attention_mask is not None else torch.ones_like(input_ids)
    first_valid = (am != 0).to(torch.int32).argmax(dim=1).to(torch.int32).contiguous()
    n_valid = am.sum(dim=1).to(torch.int32).contiguous()
    if prev is not None:  # same buffers (the captured graph holds their addresses), new contents
        prev["ctx_len"].fill_(l)
        prev["first_valid"].copy_(first_valid)
        prev["n_valid"].copy_(n_valid)
        prev["attn_cnt"].zero_()
        prev["status"] = out["status"]
        return logits, prev
    state = dict(kv=kv, ctx_len=torch.full((b,), l, dtype=torch.int32, device=input_ids.device),
                 first_valid=first_valid, n_valid=n_valid, status=out["status"])
    if reuse_slot is not None:
        while len(states) >= _STATE_CACHE_MAX:
            states.pop(next(iter(states)))
        states[key] = state
    # flash-decoding splits: enough (head, sequence, split) units to cover the SMs, no more —
    # every extra split adds to the merge (measured at ctx ~ 1000: batch 1 best at 8, batch 8 at 4)
    total = capacity
    hb = heads * b
    splits = 8 if hb <= 64 else (4 if hb <= 256 else 2)
    if total <= 256:
        splits = min(splits, 2)
    if os.environ.get("VB_ATTN_SPLITS"):  # measurement knob
        splits = int(os.environ["VB_ATTN_SPLITS"])
    state["attn_splits"] = splits
    state["attn_ws"] = torch.empty(b * heads * splits * (hd + 2), dtype=torch.float32, device=input_ids.device)
    state["attn_cnt"] = torch.zeros(b * heads, dtype=torch.int32, device=input_ids.device)
    return logits, state


class DecodeProgram:
    """One decode step as a ``vb_decode_op`` program for the persistent kernel
    (``vb_decode_step``): embed, 32 x [LN+qkv, paged attention, out_proj+residual,
    LN+fc1+act, fc2+residual], final LN + tied head.  All activation buffers are owned by the
    program, so the step is allocation-free and CUDA-graph replayable."""

    def __init__(self, lm, cache: PackCache, state: dict, batch: int, device) -> None:
        import numpy as np

        cfg = lm.config
        w = pack_opt(lm, cache, need_backward=False)
        dim, heads, hd = _dims(cfg)
        act = ops.EPI_RELU if cfg.activation_function == "relu" else ops.EPI_GELU
        kv: PagedKV = state["kv"]
        ffn = w["layers"][0]["fc1_w"].shape[0]
        vocab = w["embed"].shape[0]
        bf = dict(dtype=torch.bfloat16, device=device)
        self.tokens = torch.zeros(batch, dtype=torch.long, device=device)
        self.xa, self.xb = torch.empty((batch, dim), **bf), torch.empty((batch, dim), **bf)
        self.qkv = torch.empty((batch, 3 * dim), **bf)
        self.att = torch.empty((batch, dim), **bf)
        self.f1 = torch.empty((batch, ffn), **bf)
        self.logits = torch.empty((batch, vocab), dtype=torch.float32, device=device)
        # (head, sequence, split) units: at most two 128-thread units per SM in one round
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        splits = max(1, min(16, (2 * sms) // max(1, heads * batch)))
        max_ctx = kv.page_size * kv.max_pages
        chunk_cap = (max_ctx + splits - 1) // splits
        self.ws = torch.empty(batch * heads * splits * (hd + 2), dtype=torch.float32, device=device)
        self.cnt = torch.zeros(batch * heads, dtype=torch.int32, device=device)
        self.barrier = torch.zeros(ops.DECODE_STEP_WS_BYTES // 4, dtype=torch.int32, device=device)
        self.batch = batch
        self._keep = (w, kv, state["n_valid"], state["ctx_len"], state["first_valid"])

        recs = []

        def rec(kind, ptr=(), i64=(), i32=(), f32=()):
            r = np.zeros((), dtype=ops.op_dtype())
            r["type"] = kind
            for name, vals in (("ptr", ptr), ("i64", i64), ("i32", i32), ("f32", f32)):
                for j, v in enumerate(vals):
                    r[name][j] = 0 if v is None else (v.data_ptr() if isinstance(v, torch.Tensor) else v)
            recs.append(r)

        def gemv(x, wt, bias, y, *, residual=None, ln=(None, None), alpha=1.0, alpha_cols=0, epilogue=ops.EPI_NONE):
            n, k = wt.shape
            rec(ops.OP_GEMV, ptr=(wt, bias, residual, x, y, ln[0], ln[1]),
                i64=(n, k, wt.stride(0), x.stride(0), y.stride(0), 0 if residual is None else residual.stride(0),
                     alpha_cols),
                f32=(alpha, 1e-5), i32=(epilogue, 1 if y.dtype == torch.float32 else 0))

        rec(ops.OP_EMBED, ptr=(self.tokens, w["embed"], w["pos"], state["n_valid"], state["ctx_len"], self.xa),
            i64=(dim, vocab, w["pos"].shape[0], 2))
        for li, lw in enumerate(w["layers"]):
            gemv(self.xa, lw["qkv_w"], lw["qkv_b"], self.qkv, ln=(lw["ln1_g"], lw["ln1_b"]),
                 alpha=hd ** -0.5, alpha_cols=dim)
            rec(ops.OP_ATTN, ptr=(self.qkv, kv.k[li], kv.v[li], kv.table, state["ctx_len"], state["first_valid"],
                                  self.att, self.ws, self.cnt),
                i32=(heads, hd, kv.page_size, kv.max_pages, splits, chunk_cap), f32=(1.0,))
            gemv(self.att, lw["out_w"], lw["out_b"], self.xb, residual=self.xa)
            gemv(self.xb, lw["fc1_w"], lw["fc1_b"], self.f1, ln=(lw["ln2_g"], lw["ln2_b"]), epilogue=act)
            gemv(self.f1, lw["fc2_w"], lw["fc2_b"], self.xa, residual=self.xb)
        gemv(self.xa, w["embed"], None, self.logits, ln=(w["lnf_g"], w["lnf_b"]))
        self.host = np.stack(recs)
        self.dev = torch.from_numpy(self.host.view(np.uint8).reshape(-1).copy()).to(device)

    def step(self, tokens: torch.Tensor) -> torch.Tensor:
        if tokens.data_ptr() != self.tokens.data_ptr():
            self.tokens.copy_(tokens)
        ops.decode_step(self.host, self.dev, self.batch, self.barrier)
        return self.logits


def _decode_program(lm, cache: PackCache, state: dict, batch: int, device):
    """The persistent-kernel program of this generation state (opt-in: VB_DECODE_PERSISTENT=1),
    or None when it is off or the shapes are outside what ``vb_decode_step`` takes (batch > 8,
    feature sizes not a multiple of 64).  Measured on B200 (profiles/r01_decode_trace.txt) the
    op-by-op path under programmatic dependent launch is the faster one today: the grid
    barrier + the in-order per-SM load path make every op of the persistent kernel pay
    ~9 us of latency, so it stays an experiment behind the flag."""
    if "program" not in state:
        prog = None
        dim = lm.config.hidden_size
        if (batch <= 8 and dim % 64 == 0 and lm.config.ffn_dim % 64 == 0
                and os.environ.get("VB_DECODE_PERSISTENT", "0") == "1"):
            prog = DecodeProgram(lm, cache, state, batch, device)
        state["program"] = prog
    return state["program"]


def opt_decode_step(lm, cache: PackCache, tokens: torch.Tensor, state: dict) -> torch.Tensor:
    """One token per sequence: tokens (B,) int64 -> next-position logits f32 (B, V).
    Stream-ordered and allocation-stable, so it can be captured into a CUDA graph
    (``DecodeGraph``): the per-sequence counters are advanced in place.  Default: op by op —
    every projection is a tensor-core GEMV launched with programmatic dependent launch, so
    its first weight loads run under the tail of its predecessor; VB_DECODE_PERSISTENT=1
    runs the whole step as ONE persistent launch (``DecodeProgram``) instead."""
    cfg = lm.config
    b = tokens.shape[0]
    if b > 16:
        raise NotImplementedError("decode batch (incl. beams) > 16 is not supported yet")
    prog = _decode_program(lm, cache, state, b, tokens.device)
    if prog is not None:
        return prog.step(tokens)
    w = pack_opt(lm, cache, need_backward=False)
    dim, heads, hd = _dims(cfg)
    act = ops.EPI_RELU if cfg.activation_function == "relu" else ops.EPI_GELU
    scaling = hd ** -0.5
    kv: PagedKV = state["kv"]
    # position of the new token = number of valid tokens so far - 1 + offset 2 (HF :350-354)
    x = ops.decode_embed(tokens, w["embed"], w["pos"], state["n_valid"], state["ctx_len"], 2)
    for li, lw in enumerate(w["layers"]):
        qkv = ops.gemv(x, lw["qkv_w"], lw["qkv_b"], alpha=scaling, alpha_cols=dim,
                       ln=(lw["ln1_g"], lw["ln1_b"], 1e-5))
        o = ops.paged_decode_attention(qkv, kv.k[li], kv.v[li], kv.table, state["ctx_len"],
                                       state["first_valid"], heads, kv.page_size, 1.0,
                                       workspace=state["attn_ws"], counters=state["attn_cnt"],
                                       splits=state["attn_splits"])
        x = ops.gemv(o, lw["out_w"], lw["out_b"], residual=x)
        f1 = ops.gemv(x, lw["fc1_w"], lw["fc1_b"], epilogue=act, ln=(lw["ln2_g"], lw["ln2_b"], 1e-5))
        x = ops.gemv(f1, lw["fc2_w"], lw["fc2_b"], residual=x)
    return ops.gemv(x, w["embed"], out_dtype=torch.float32, ln=(w["lnf_g"], w["lnf_b"], 1e-5))


def decode_graph_for(lm, cache: PackCache, state: dict, batch: int, device) -> "DecodeGraph":
    """The decode-step graph of `state`, captured once per (state, packed weights): a reused state
    (opt_prefill(reuse_slot=...)) keeps its graph across generate() calls as long as the packed LM
    operands the graph points at are still the cached ones."""
    w = pack_opt(lm, cache, need_backward=False)
    g = state.get("_graph")
    if g is not None and state.get("_graph_pack") is w and g.tokens.shape[0] == batch:
        return g
    g = DecodeGraph(lm, cache, state, batch, device)
    state["_graph"], state["_graph_pack"] = g, w
    return g


class DecodeGraph:
    """CUDA graph of one decode step (~230 launches -> one graph launch): static token
    buffer in, static logits buffer out; the KV cache and the counters live in `state`."""

    def __init__(self, lm, cache: PackCache, state: dict, batch: int, device) -> None:
        self.tokens = torch.zeros(batch, dtype=torch.long, device=device)
        snap = {k: state[k].clone() for k in ("n_valid", "ctx_len")}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            opt_decode_step(lm, cache, self.tokens, state)  # warm-up (allocator, attributes)
        torch.cuda.current_stream().wait_stream(side)
        for k, v in snap.items():
            state[k].copy_(v)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits = opt_decode_step(lm, cache, self.tokens, state)
        for k, v in snap.items():  # capture does not execute; restore defensively anyway
            state[k].copy_(v)

    def step(self, tokens: torch.Tensor) -> torch.Tensor:
        self.tokens.copy_(tokens)
        self.graph.replay()
        return self.logits

    # ---- greedy decoding with the token bookkeeping inside the graph -------------------------------------
    # generate()'s greedy loop (HF GenerationMixin greedy search: EOS suppression below min_new_tokens, argmax,
    # finished rows emit pad, stop when every row has emitted EOS) costs ~10 small torch launches and ~0.3 ms of
    # host time per token when written as Python between graph replays.  Here one graph replay does a whole
    # iteration: bookkeeping of the logits at hand -> next token -> decode step -> next logits; the position is
    # a device counter, so nothing but the replay itself is issued per token.
    def _greedy_bookkeeping(self) -> None:
        b = self.tokens.shape[0]
        scores = self.g_logits
        if self.g_eos.numel() > 0:
            blocked = torch.where(self.g_pos < self.g_min_new, float("-inf"), 0.0).to(scores.dtype)  # (1,)
            scores = scores.index_add(1, self.g_eos, blocked.expand(b, self.g_eos.numel()))
        nxt = scores.argmax(dim=-1)
        nxt = torch.where(self.g_unfinished, nxt, self.g_pad.expand(b))
        self.g_out.scatter_(1, self.g_pos.expand(b, 1), nxt[:, None])
        if self.g_eos.numel() > 0:
            self.g_unfinished.logical_and_(~(nxt[:, None] == self.g_eos[None, :]).any(dim=1))
        self.g_alive.scatter_(0, self.g_pos, self.g_unfinished.sum(dtype=torch.int32)[None])
        self.g_pos.add_(1)
        self.tokens.copy_(nxt)

    def greedy_begin(self, lm, cache: PackCache, state: dict, logits: torch.Tensor, max_new: int, min_new: int,
                     eos_ids, pad_id: int) -> None:
        """Arms the greedy graph with the prefill's logits.  (Re)captures when the capacity or the number of EOS
        ids grew; the captured kernels point at the static tensors below."""
        b, dev = self.tokens.shape[0], self.tokens.device
        n_eos = len(eos_ids) if eos_ids else 0
        cap = -(-max_new // 64) * 64
        stale = (getattr(self, "g_graph", None) is None or self.g_out.shape[1] < max_new or self.g_eos.numel() != n_eos)
        if stale:
            self.g_logits = torch.empty_like(self.logits)
            self.g_out = torch.empty((b, cap), dtype=torch.long, device=dev)
            self.g_pos = torch.zeros(1, dtype=torch.long, device=dev)
            self.g_min_new = torch.zeros(1, dtype=torch.long, device=dev)
            self.g_unfinished = torch.ones(b, dtype=torch.bool, device=dev)
            self.g_alive = torch.ones(cap, dtype=torch.int32, device=dev)
            self.g_eos = torch.zeros(n_eos, dtype=torch.long, device=dev)
            self.g_pad = torch.zeros(1, dtype=torch.long, device=dev)
        self.g_logits.copy_(logits)
        self.g_out.fill_(int(pad_id))
        self.g_pos.zero_()
        self.g_min_new.fill_(int(min_new))
        self.g_unfinished.fill_(True)
        self.g_alive.fill_(1)
        self.g_pad.fill_(int(pad_id))
        if n_eos:
            self.g_eos.copy_(torch.tensor(list(eos_ids), dtype=torch.long, device=dev))
        if stale:
            snap = {k: state[k].clone() for k in ("n_valid", "ctx_len")}
            keep = {k: getattr(self, k).clone() for k in ("g_logits", "g_out", "g_pos", "g_unfinished", "g_alive")}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up of the bookkeeping ops (allocator)
                self._greedy_bookkeeping()
            torch.cuda.current_stream().wait_stream(side)
            for k, v in keep.items():
                getattr(self, k).copy_(v)
            self.g_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_graph):
                self._greedy_bookkeeping()
                self.g_logits.copy_(opt_decode_step(lm, cache, self.tokens, state))
            for k, v in snap.items():  # capture does not execute; restore defensively anyway
                state[k].copy_(v)
            for k, v in keep.items():
                getattr(self, k).copy_(v)

    def greedy_iteration(self) -> None:
        """Token of the current position -> g_out, then the decode step that produces the next logits."""
        self.g_graph.replay()

    def greedy_last(self) -> None:
        """The last position: bookkeeping only (its token is never fed back)."""
        self._greedy_bookkeeping()
