"""ORACLE — test infrastructure only.  NOT part of the product path.

A CPU, fp32, plain-PyTorch restatement of the reference's VideoBLIP forward path
(yukw777/EILEV ``eilev/model/v2.py`` and the HuggingFace modules it calls), written as
pure functions over a ``state_dict`` with the reference's parameter names.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker / the CPU baseline.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *real* reference
(``/root/reference/eilev/model/v2.py`` on the installed transformers 5.5.0; the reference
pins 4.33.1, whose equations for this path are the same) on seeded inputs and commits
inputs + outputs under ``tests/golden/`` (``make_golden_classify.py`` does the same for ``classify``
— the reference's own method behind a tuple<->Cache shim —, ``make_golden_t5.py`` for the flan-T5
branch incl. greedy ``generate``, ``make_golden_beams.py`` for HF beam search / repetition penalty, ``make_golden_v1.py`` for
the v1 class (eilev/model/v1.py behind a shim restoring the 4.33.1 prepend contract));
``tests/test_oracle.py`` checks this restatement against those fixtures (and against the live
reference when ``/root/reference`` exists).
``normalize_frames`` (the image processor's rescale + normalize) is pinned bit-exactly to the HF
``image_transforms.rescale`` / ``normalize`` functions; the frame resize lives in
``oracle/pil_resize_ref.py`` (pinned bit-exactly to ``PIL.Image.resize``).
The reference's own tests pin shapes only (tests/model/test_model_v2.py:53-83,185-186).

Citations: ``v2.py`` = eilev/model/v2.py; ``HF:`` = transformers/models/…
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float(), eps)


def _lin(x, sd, prefix, bias=True):
    b = sd.get(prefix + ".bias") if bias else None
    return F.linear(x, sd[prefix + ".weight"].float(), None if b is None else b.float())


# --------------------------------------------------------------------------- vision tower
def vision_embeddings(sd, vcfg, frames, p="vision_model.embeddings."):
    """HF:blip_2/modeling_blip_2.py:243-255 — Conv2d(k=s=patch) + [CLS] + position table."""
    w = sd[p + "patch_embedding.weight"].float()
    b = sd[p + "patch_embedding.bias"].float()
    x = F.conv2d(frames.float(), w, b, stride=vcfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "class_embedding"].float().expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1)
    return x + sd[p + "position_embedding"].float()[:, : x.shape[1]]


def vision_layer(sd, vcfg, x, p):
    """HF:blip_2/modeling_blip_2.py:383-402 (layer), :319-353 (attention), :365-369 (MLP)."""
    h, d = vcfg.num_attention_heads, vcfg.hidden_size // vcfg.num_attention_heads
    b, s, _ = x.shape
    y = _ln(x, sd, p + "layer_norm1", vcfg.layer_norm_eps)
    qkv = _lin(y, sd, p + "self_attn.qkv").reshape(b, s, 3, h, d).permute(2, 0, 3, 1, 4)
    att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * d ** -0.5, dim=-1) @ qkv[2]
    att = att.transpose(1, 2).reshape(b, s, h * d)
    x = x + _lin(att, sd, p + "self_attn.projection")
    y = _ln(x, sd, p + "layer_norm2", vcfg.layer_norm_eps)
    act = F.gelu if vcfg.hidden_act == "gelu" else getattr(F, vcfg.hidden_act)
    y = _lin(act(_lin(y, sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    return x + y


def vision_forward(sd, vcfg, pixel_values):
    """v2.py:24-103 on top of HF:blip_2/modeling_blip_2.py:506-531.

    pixel_values (N, C, T, H, W) -> last_hidden_state (N, T*S, D), pooler_output (N, T, D).
    """
    n, _, t, _, _ = pixel_values.shape
    frames = pixel_values.permute(0, 2, 1, 3, 4).flatten(end_dim=1)  # v2.py:57
    x = vision_embeddings(sd, vcfg, frames)
    for i in range(vcfg.num_hidden_layers):
        x = vision_layer(sd, vcfg, x, f"vision_model.encoder.layers.{i}.")
    x = _ln(x, sd, "vision_model.post_layernorm", vcfg.layer_norm_eps)
    pooled = _ln(x[:, 0], sd, "vision_model.post_layernorm", vcfg.layer_norm_eps)  # LN twice (:525-526)
    s = x.shape[1]
    return x.reshape(n, t * s, -1), pooled.reshape(n, t, -1)  # v2.py:69-75


# --------------------------------------------------------------------------- Q-Former
def _no_drop(site, t):
    return t


def _qf_attention(sd, qcfg, hidden, kv_src, p, drop=_no_drop, sites=(None, None)):
    """HF:blip_2/modeling_blip_2.py:579-633 (eager, scores / sqrt(d); masks are all-ones here)
    followed by Blip2QFormerSelfOutput :644-648 (dense + LN(residual))."""
    h = qcfg.num_attention_heads
    d = qcfg.hidden_size // h
    b, sq, _ = hidden.shape

    def heads(x):
        return x.reshape(b, x.shape[1], h, d).permute(0, 2, 1, 3)

    q = heads(_lin(hidden, sd, p + "attention.query"))
    k = heads(_lin(kv_src, sd, p + "attention.key"))
    v = heads(_lin(kv_src, sd, p + "attention.value"))
    probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
    probs = drop(sites[0], probs)  # attention_probs dropout (:621-624)
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(b, sq, h * d)
    out = drop(sites[1], _lin(ctx, sd, p + "output.dense"))  # hidden dropout (:646)
    return _ln(out + hidden, sd, p + "output.LayerNorm", qcfg.layer_norm_eps)


def qformer_forward(sd, qcfg, query_tokens, image_embeds, drop=_no_drop):
    """HF:blip_2/modeling_blip_2.py:962-1036, layers :729-780 (query-only path).
    `drop(site, tensor)` injects dropout masks (tests only; identity = eval mode); sites are
    ("qf", layer, k) with k = 0 self-attn probs, 1 self-output, 2 cross probs, 3 cross-output,
    4 FFN output, and ("qf", -1, 7) for the embedding dropout (:985)."""
    x = _ln(query_tokens.float(), sd, "qformer.layernorm", qcfg.layer_norm_eps)  # :984-985
    x = drop(("qf", -1, 7), x)
    for i in range(qcfg.num_hidden_layers):
        p = f"qformer.encoder.layer.{i}."
        x = _qf_attention(sd, qcfg, x, x, p + "attention.", drop, (("qf", i, 0), ("qf", i, 1)))
        if i % qcfg.cross_attention_frequency == 0:  # :716-720
            x = _qf_attention(sd, qcfg, x, image_embeds.float(), p + "crossattention.", drop,
                              (("qf", i, 2), ("qf", i, 3)))
        act = F.gelu if qcfg.hidden_act == "gelu" else getattr(F, qcfg.hidden_act)
        inter = act(_lin(x, sd, p + "intermediate_query.dense"))
        x = _ln(drop(("qf", i, 4), _lin(inter, sd, p + "output_query.dense")) + x, sd,
                p + "output_query.LayerNorm", qcfg.layer_norm_eps)
    return x


# --------------------------------------------------------------------------- OPT
def opt_positions(attention_mask):
    """HF:opt/modeling_opt.py:350-354 + offset 2 (:45-70)."""
    am = attention_mask.long()
    return (torch.cumsum(am, dim=1) * am - 1) + 2


def opt_decoder(sd, tcfg, inputs_embeds, attention_mask, p="language_model.model.decoder.", drop=_no_drop):
    """HF:opt/modeling_opt.py:321-396; layers :202-253; attention :135-181 (q scaled first).
    drop sites: ("opt", layer, 0) attention probs, 1 after out_proj (:219), 2 after fc2 (:243)."""
    assert tcfg.do_layer_norm_before and tcfg.word_embed_proj_dim == tcfg.hidden_size
    b, l, _ = inputs_embeds.shape
    h = tcfg.num_attention_heads
    d = tcfg.hidden_size // h
    pos = sd[p + "embed_positions.weight"].float()[opt_positions(attention_mask)]
    x = inputs_embeds.float() + pos
    neg = torch.finfo(torch.float32).min
    causal = torch.ones(l, l, dtype=torch.bool).tril()
    allowed = causal[None, None] & attention_mask.bool()[:, None, None, :]
    bias = torch.zeros(b, 1, l, l).masked_fill(~allowed, neg)
    for i in range(tcfg.num_hidden_layers):
        lp = f"{p}layers.{i}."
        y = _ln(x, sd, lp + "self_attn_layer_norm", 1e-5)

        def heads(t):
            return t.reshape(b, l, h, d).transpose(1, 2)

        q = heads(_lin(y, sd, lp + "self_attn.q_proj") * d ** -0.5)
        k = heads(_lin(y, sd, lp + "self_attn.k_proj"))
        v = heads(_lin(y, sd, lp + "self_attn.v_proj"))
        att = drop(("opt", i, 0), torch.softmax(q @ k.transpose(-1, -2) + bias, dim=-1)) @ v
        att = att.transpose(1, 2).reshape(b, l, h * d)
        x = x + drop(("opt", i, 1), _lin(att, sd, lp + "self_attn.out_proj"))
        y = _ln(x, sd, lp + "final_layer_norm", 1e-5)
        act = F.relu if tcfg.activation_function == "relu" else getattr(F, tcfg.activation_function)
        x = x + drop(("opt", i, 2), _lin(act(_lin(y, sd, lp + "fc1")), sd, lp + "fc2"))
    return _ln(x, sd, p + "final_layer_norm", 1e-5)


def causal_lm_loss(logits, labels):
    """HF:loss/loss_utils.py:28-67 — shift, ignore_index=-100, mean over valid targets."""
    v = logits.shape[-1]
    return F.cross_entropy(logits[:, :-1].float().reshape(-1, v), labels[:, 1:].reshape(-1),
                           ignore_index=-100)


# --------------------------------------------------------------------------- full model
def video_features(sd, config, pixel_values, drop=_no_drop):
    """v2.py:169-203 — ViT -> Q-Former -> language_projection; rows in (clip, query) order."""
    image_embeds, pooled = vision_forward(sd, config.vision_config, pixel_values)
    n = image_embeds.shape[0]
    query = sd["query_tokens"].float().expand(n, -1, -1)
    qout = qformer_forward(sd, config.qformer_config, query, image_embeds, drop)
    feats = _lin(qout.reshape(n * config.num_query_tokens, -1), sd, "language_projection")
    return feats, qout, image_embeds, pooled


def splice(sd, input_ids, video_input_mask, feats):
    """v2.py:205-214 — boolean-mask assignment fills True slots in row-major order."""
    emb = sd["language_model.model.decoder.embed_tokens.weight"].float()[input_ids]
    if feats is not None:
        emb = emb.clone()
        emb[video_input_mask.bool()] = feats
    return emb


def videoblip_forward(sd, config, input_ids, attention_mask=None, pixel_values=None,
                      video_input_mask=None, labels=None, drop=_no_drop):
    """v2.py:132-252 for the decoder-only (OPT) language model.  Returns a dict."""
    out = {}
    feats = None
    if pixel_values is not None:
        assert video_input_mask is not None  # v2.py:154-157
        feats, qout, image_embeds, pooled = video_features(sd, config, pixel_values, drop)
        out.update(video_features=feats, query_output=qout, image_embeds=image_embeds,
                   pooler_output=pooled)
    emb = splice(sd, input_ids, video_input_mask, feats)
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)  # v2.py:216-217
    hidden = opt_decoder(sd, config.text_config, emb, attention_mask, drop=drop)
    logits = F.linear(hidden, sd["language_model.model.decoder.embed_tokens.weight"].float())  # tied
    out.update(inputs_embeds=emb, logits=logits)
    if labels is not None:
        out["loss"] = causal_lm_loss(logits, labels)
    return out


@torch.no_grad()
def greedy_generate(sd, config, input_ids, attention_mask, pixel_values, video_input_mask,
                    max_new_tokens, eos_token_id=None):
    """v2.py:254-324 with greedy search (HF:generation/utils.py:2658-2790); returns only the new
    tokens (decoder-only + inputs_embeds).  No KV cache: recomputes the prefix, small cases only."""
    feats = None
    if pixel_values is not None:
        feats = video_features(sd, config, pixel_values)[0]
    emb = splice(sd, input_ids, video_input_mask, feats)
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    table = sd["language_model.model.decoder.embed_tokens.weight"].float()
    b = emb.shape[0]
    new = []
    done = torch.zeros(b, dtype=torch.bool)
    pad = config.text_config.pad_token_id
    for _ in range(max_new_tokens):
        hidden = opt_decoder(sd, config.text_config, emb, attention_mask)
        nxt = F.linear(hidden[:, -1], table).argmax(-1)
        if eos_token_id is not None:
            nxt = torch.where(done, torch.full_like(nxt, pad), nxt)
            done = done | (nxt == eos_token_id)
        new.append(nxt)
        emb = torch.cat([emb, table[nxt][:, None]], dim=1)
        attention_mask = torch.cat([attention_mask, torch.ones(b, 1, dtype=attention_mask.dtype)], 1)
        if eos_token_id is not None and bool(done.all()):
            break
    return torch.stack(new, dim=1)


# --------------------------------------------------------------------------- flan-T5 (seq2seq LM)
def _rms(x, w, eps):
    """HF:t5/modeling_t5.py T5LayerNorm — scale only, fp32 variance without mean subtraction."""
    return w.float() * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def t5_relative_bucket(rel, bidirectional, num_buckets, max_distance):
    """HF:t5/modeling_t5.py T5Attention._relative_position_bucket (rel = key - query)."""
    ret = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        ret = ret + (rel > 0).long() * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.minimum(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(rel < max_exact, rel, large)


def t5_position_bias(table, lq, lk, bidirectional, tcfg):
    """compute_bias: (1, heads, lq, lk) from the (buckets, heads) embedding of block 0."""
    rel = torch.arange(lk)[None, :] - torch.arange(lq)[:, None]
    bucket = t5_relative_bucket(rel, bidirectional, tcfg.relative_attention_num_buckets,
                                tcfg.relative_attention_max_distance)
    return table.float()[bucket].permute(2, 0, 1)[None]


def _t5_attention(sd, tcfg, x, kv, p, bias, drop=_no_drop, site=None):
    """T5Attention.forward: no 1/sqrt(d) scaling, no projection biases, additive bias; dropout on the
    probabilities (`site`, tests only)."""
    b, lq, _ = x.shape
    lk = kv.shape[1]
    h, d = tcfg.num_heads, tcfg.d_kv
    q = _lin(x, sd, p + "q", bias=False).reshape(b, lq, h, d).transpose(1, 2)
    k = _lin(kv, sd, p + "k", bias=False).reshape(b, lk, h, d).transpose(1, 2)
    v = _lin(kv, sd, p + "v", bias=False).reshape(b, lk, h, d).transpose(1, 2)
    probs = drop(site, torch.softmax(q @ k.transpose(-1, -2) + bias, dim=-1))
    return _lin((probs @ v).transpose(1, 2).reshape(b, lq, h * d), sd, p + "o", bias=False)


def _t5_ff(sd, tcfg, x, p, drop=_no_drop, site=None):
    """T5DenseGatedActDense (gated-gelu: gelu_new(wi_0 x) * wi_1 x) or T5DenseActDense (relu); dropout on
    the inner activation (`site`)."""
    if tcfg.is_gated_act:
        act = F.gelu(_lin(x, sd, p + "wi_0", bias=False), approximate="tanh")
        return _lin(drop(site, act * _lin(x, sd, p + "wi_1", bias=False)), sd, p + "wo", bias=False)
    return _lin(drop(site, F.relu(_lin(x, sd, p + "wi", bias=False))), sd, p + "wo", bias=False)


def t5_encoder(sd, tcfg, inputs_embeds, attention_mask, p="language_model.encoder.", drop=_no_drop):
    """T5Stack (encoder): pre-RMSNorm blocks; block 0 owns the bidirectional relative bias.
    drop sites (HF T5, tests only): ("t5e", layer, 0) attention probabilities, 1 attention output, 2 feed-forward
    inner activation, 3 feed-forward output; ("t5", -1, 0) the embeddings, ("t5", -1, 1) the stack output."""
    b, l, _ = inputs_embeds.shape
    eps = tcfg.layer_norm_epsilon
    neg = torch.finfo(torch.float32).min
    bias = t5_position_bias(sd[p + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"], l, l, True, tcfg)
    bias = bias + (1.0 - attention_mask[:, None, None, :].float()) * neg
    x = drop(("t5", -1, 0), inputs_embeds.float())
    for i in range(tcfg.num_layers):
        bp = f"{p}block.{i}."
        y = _rms(x, sd[bp + "layer.0.layer_norm.weight"], eps)
        x = x + drop(("t5e", i, 1), _t5_attention(sd, tcfg, y, y, bp + "layer.0.SelfAttention.", bias, drop, ("t5e", i, 0)))
        x = x + drop(("t5e", i, 3), _t5_ff(sd, tcfg, _rms(x, sd[bp + "layer.1.layer_norm.weight"], eps),
                                           bp + "layer.1.DenseReluDense.", drop, ("t5e", i, 2)))
    return drop(("t5", -1, 1), _rms(x, sd[p + "final_layer_norm.weight"], eps))


def t5_decoder(sd, tcfg, decoder_input_ids, enc, enc_mask, p="language_model.decoder.", drop=_no_drop):
    """T5Stack (decoder): causal self-attention with the unidirectional relative bias,
    cross-attention over the encoder states (mask only), feed-forward.
    drop sites: ("t5d", layer, 0..3) as in the encoder, 4 cross-attention probabilities, 5 cross-attention
    output; ("t5", -1, 2) the embeddings, ("t5", -1, 3) the stack output."""
    b, l = decoder_input_ids.shape
    eps = tcfg.layer_norm_epsilon
    neg = torch.finfo(torch.float32).min
    x = drop(("t5", -1, 2), sd["language_model.shared.weight"].float()[decoder_input_ids])
    self_bias = t5_position_bias(sd[p + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"], l, l, False, tcfg)
    self_bias = self_bias.masked_fill(~torch.ones(l, l, dtype=torch.bool).tril()[None, None], neg)
    cross_bias = (1.0 - enc_mask[:, None, None, :].float()) * neg
    for i in range(tcfg.num_decoder_layers):
        bp = f"{p}block.{i}."
        y = _rms(x, sd[bp + "layer.0.layer_norm.weight"], eps)
        x = x + drop(("t5d", i, 1), _t5_attention(sd, tcfg, y, y, bp + "layer.0.SelfAttention.", self_bias, drop,
                                                  ("t5d", i, 0)))
        y = _rms(x, sd[bp + "layer.1.layer_norm.weight"], eps)
        x = x + drop(("t5d", i, 5), _t5_attention(sd, tcfg, y, enc, bp + "layer.1.EncDecAttention.", cross_bias, drop,
                                                  ("t5d", i, 4)))
        x = x + drop(("t5d", i, 3), _t5_ff(sd, tcfg, _rms(x, sd[bp + "layer.2.layer_norm.weight"], eps),
                                           bp + "layer.2.DenseReluDense.", drop, ("t5d", i, 2)))
    return drop(("t5", -1, 3), _rms(x, sd[p + "final_layer_norm.weight"], eps))


def t5_shift_right(labels, tcfg):
    """T5ForConditionalGeneration._shift_right: start token first, -100 -> pad."""
    out = labels.new_zeros(labels.shape)
    out[:, 1:] = labels[:, :-1]
    out[:, 0] = tcfg.decoder_start_token_id
    return out.masked_fill(out == -100, tcfg.pad_token_id)


def videoblip_forward_t5(sd, config, input_ids, attention_mask=None, pixel_values=None,
                         video_input_mask=None, labels=None, decoder_input_ids=None, drop=_no_drop):
    """v2.py:132-252 with the seq2seq branch (:228-238) -> T5ForConditionalGeneration.forward:
    encoder over the interleaved embeddings, decoder over shift_right(labels), untied or tied
    (x d_model**-0.5) head, unshifted CE with ignore_index -100."""
    tcfg = config.text_config
    out = {}
    feats = None
    if pixel_values is not None:
        assert video_input_mask is not None
        feats, qout, image_embeds, pooled = video_features(sd, config, pixel_values, drop)
        out.update(video_features=feats, query_output=qout, image_embeds=image_embeds, pooler_output=pooled)
    emb = sd["language_model.shared.weight"].float()[input_ids]
    if feats is not None:
        emb = emb.clone()
        emb[video_input_mask.bool()] = feats
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    enc = t5_encoder(sd, tcfg, emb, attention_mask, drop=drop)
    if decoder_input_ids is None:
        decoder_input_ids = t5_shift_right(labels, tcfg)
    dec = t5_decoder(sd, tcfg, decoder_input_ids, enc, attention_mask, drop=drop)
    # 4.33.1 scales the decoder output by d_model**-0.5 iff tie_word_embeddings; 5.5.0 froze that
    # decision into config.scale_decoder_outputs (flan-t5: untied head, no scaling, in both)
    if getattr(tcfg, "scale_decoder_outputs", tcfg.tie_word_embeddings):
        dec = dec * tcfg.d_model ** -0.5
    head = sd["language_model.lm_head.weight"]
    logits = F.linear(dec, head.float())
    out.update(inputs_embeds=emb, encoder_last_hidden_state=enc, logits=logits)
    if labels is not None:
        out["loss"] = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    return out


@torch.no_grad()
def greedy_generate_t5(sd, config, input_ids, attention_mask, pixel_values, video_input_mask, max_new_tokens,
                       eos_token_id=None, return_margins=False):
    """v2.py:254-324 with the seq2seq LM and greedy search: encoder once, decoder re-run on the
    growing prefix (no cache; small cases only).  Returns [decoder_start] + new tokens, rows that
    hit eos are padded afterwards (HF:generation/utils.py greedy loop)."""
    tcfg = config.text_config
    feats = None
    if pixel_values is not None:
        feats = video_features(sd, config, pixel_values)[0]
    emb = sd["language_model.shared.weight"].float()[input_ids]
    if feats is not None:
        emb = emb.clone()
        emb[video_input_mask.bool()] = feats
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    enc = t5_encoder(sd, tcfg, emb, attention_mask)
    b = input_ids.shape[0]
    seq = torch.full((b, 1), tcfg.decoder_start_token_id, dtype=torch.long)
    done = torch.zeros(b, dtype=torch.bool)
    scale = tcfg.d_model ** -0.5 if getattr(tcfg, "scale_decoder_outputs", tcfg.tie_word_embeddings) else 1.0
    margins = []  # fp32 top-1 minus top-2 logit per step: tells a bf16 tie from an error in the tests
    for _ in range(max_new_tokens):
        dec = t5_decoder(sd, tcfg, seq, enc, attention_mask)
        step_logits = F.linear(dec[:, -1] * scale, sd["language_model.lm_head.weight"].float())
        top2 = step_logits.topk(2, dim=-1).values
        margins.append(top2[:, 0] - top2[:, 1])
        nxt = step_logits.argmax(-1)
        if eos_token_id is not None:
            nxt = torch.where(done, torch.full_like(nxt, tcfg.pad_token_id), nxt)
            done = done | (nxt == eos_token_id)
        seq = torch.cat([seq, nxt[:, None]], dim=1)
        if eos_token_id is not None and bool(done.all()):
            break
    return (seq, torch.stack(margins)) if return_margins else seq


# --------------------------------------------------------------------------- frame normalisation
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def normalize_frames(frames_u8, rescale=1 / 255, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD):
    """eilev/model/utils.py:5-26 ``process`` -> ``Blip2Processor`` -> ``BlipImageProcessor.preprocess``
    of the pinned transformers 4.33.1 on frames that already have the target size: ``rescale``
    (HF:image_transforms.py ``rescale``: uint8 array * python float in float64, cast to float32)
    then ``normalize`` (HF:image_transforms.py ``normalize``: (image - mean) / std in float32).
    frames_u8: uint8 with the channel axis at dim 1, e.g. (N, C, T, H, W) or (N*T, C, H, W), and
    C = len(mean).  Pinned bit-exactly against those two HF functions in tests/test_oracle.py."""
    shape = [1] * frames_u8.dim()
    shape[1] = len(mean)
    m = torch.tensor(mean, dtype=torch.float32).view(shape)
    sd_ = torch.tensor(std, dtype=torch.float32).view(shape)
    x = (frames_u8.to(torch.float64) * rescale).to(torch.float32)
    return (x - m) / sd_


# --------------------------------------------------------------------------- v1 (HF 4.33.1 Blip2 forward)
def videoblip_forward_v1(sd, config, pixel_values, input_ids, attention_mask=None, labels=None,
                         decoder_input_ids=None):
    """eilev/model/v1.py:95-119 inherits ``Blip2ForConditionalGeneration.forward`` of the pinned
    transformers 4.33.1 (HF 4.33.1 blip_2/modeling_blip_2.py ~:1680-1780; not available offline,
    restated from its published algorithm — transformers 5.5.0 keeps the same loss block,
    HF:blip_2/modeling_blip_2.py:1802-1822): one video per row, the projected query rows are
    CONCATENATED in front of the embedded prompt, the mask is ones(Q) ++ attention_mask, and the
    decoder-only loss is the shifted CE over the last ``labels.size(1)`` logits, which are also the
    logits returned.  Seq2seq: labels / decoder_input_ids go to T5 untouched."""
    tcfg = config.text_config
    b = input_ids.shape[0]
    assert pixel_values.shape[0] == b
    feats, qout, image_embeds, pooled = video_features(sd, config, pixel_values)
    q = config.num_query_tokens
    lm_inputs = feats.view(b, q, -1)
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    full_mask = torch.cat([torch.ones(b, q, dtype=attention_mask.dtype), attention_mask], dim=1)
    out = dict(video_features=feats, query_output=qout, image_embeds=image_embeds, pooler_output=pooled)
    if config.use_decoder_only_language_model:
        table = sd["language_model.model.decoder.embed_tokens.weight"].float()
        emb = torch.cat([lm_inputs, table[input_ids]], dim=1)
        hidden = opt_decoder(sd, tcfg, emb, full_mask)
        logits = F.linear(hidden, table)
        out["full_logits"] = logits
        if labels is not None:
            logits = logits[:, -labels.size(1):, :]
            v = logits.shape[-1]
            out["loss"] = F.cross_entropy(logits[:, :-1].reshape(-1, v), labels[:, 1:].reshape(-1))
        out["logits"] = logits
        return out
    emb = torch.cat([lm_inputs, sd["language_model.shared.weight"].float()[input_ids]], dim=1)
    enc = t5_encoder(sd, tcfg, emb, full_mask)
    if decoder_input_ids is None:
        decoder_input_ids = t5_shift_right(labels, tcfg)
    dec = t5_decoder(sd, tcfg, decoder_input_ids, enc, full_mask)
    if getattr(tcfg, "scale_decoder_outputs", tcfg.tie_word_embeddings):
        dec = dec * tcfg.d_model ** -0.5
    logits = F.linear(dec, sd["language_model.lm_head.weight"].float())
    out.update(encoder_last_hidden_state=enc, logits=logits)
    if labels is not None:
        out["loss"] = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), ignore_index=-100)
    return out


@torch.no_grad()
def greedy_generate_v1(sd, config, pixel_values, input_ids=None, attention_mask=None, max_new_tokens=4):
    """HF 4.33.1 ``Blip2ForConditionalGeneration.generate`` with a decoder-only LM and greedy search:
    prompt defaults to [bos] per row; embeddings = cat([video rows, embed(prompt)]); only the new
    tokens are returned.  No cache (prefix recomputed): small cases only."""
    tcfg = config.text_config
    b = pixel_values.shape[0]
    if input_ids is None:
        input_ids = torch.full((b, 1), tcfg.bos_token_id, dtype=torch.long)
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    q = config.num_query_tokens
    feats = video_features(sd, config, pixel_values)[0]
    table = sd["language_model.model.decoder.embed_tokens.weight"].float()
    emb = torch.cat([feats.view(b, q, -1), table[input_ids]], dim=1)
    mask = torch.cat([torch.ones(b, q, dtype=attention_mask.dtype), attention_mask], dim=1)
    new = []
    for _ in range(max_new_tokens):
        hidden = opt_decoder(sd, tcfg, emb, mask)
        nxt = F.linear(hidden[:, -1], table).argmax(-1)
        new.append(nxt)
        emb = torch.cat([emb, table[nxt][:, None]], dim=1)
        mask = torch.cat([mask, torch.ones(b, 1, dtype=mask.dtype)], dim=1)
    return torch.stack(new, dim=1)


@torch.no_grad()
def classify(sd, config, prompt_input_ids, class_input_ids, prompt_attention_mask=None, pixel_values=None,
             prompt_video_input_mask=None, class_attention_mask=None):
    """v2.py:326-501 — mean log-likelihood (batch, num_classes) of each class continuation
    after the (left-padded) prompt.  The reference feeds the class tokens on top of the
    prompt's KV cache (:462-467); without a cache that is the decoder run on the
    concatenation [prompt ; class] under the concatenated mask (:443-455) — positions come
    from the mask cumsum either way (HF:opt/modeling_opt.py:350-354).  Class token j is
    scored by the logits of the position before it (the prompt's last position for j = 0,
    :469-478); padded class tokens are ignored (:480-481) and the sum is divided by the
    class length (:497-501).  Small cases only (one decoder pass per (row, class))."""
    feats = None
    if pixel_values is not None:
        feats = video_features(sd, config, pixel_values)[0]
    prompt_emb = splice(sd, prompt_input_ids, prompt_video_input_mask, feats)
    if prompt_attention_mask is None:
        prompt_attention_mask = torch.ones_like(prompt_input_ids)
    if class_attention_mask is None:
        class_attention_mask = torch.ones_like(class_input_ids)
    table = sd["language_model.model.decoder.embed_tokens.weight"].float()
    b, lp = prompt_input_ids.shape
    n_cls, lc = class_input_ids.shape
    out = torch.zeros(b, n_cls)
    for bi in range(b):
        for ci in range(n_cls):
            emb = torch.cat([prompt_emb[bi:bi + 1], table[class_input_ids[ci]][None]], dim=1)
            mask = torch.cat([prompt_attention_mask[bi:bi + 1], class_attention_mask[ci:ci + 1]], dim=1)
            hidden = opt_decoder(sd, config.text_config, emb, mask)
            logp = F.log_softmax(F.linear(hidden[0, lp - 1:lp + lc - 1], table), dim=-1)  # (lc, V)
            tok = logp.gather(1, class_input_ids[ci][:, None])[:, 0]
            valid = class_attention_mask[ci] != 0
            out[bi, ci] = (tok * valid).sum() / valid.sum()
    return out


def sane_init_(state_dict, seed=1234, std=0.02):
    """Seeded, numerically sane re-initialisation (HF's default init is degenerate for this
    model: ViT std 1e-10, zero query tokens — SURVEY §0.8).  Linear/conv/embedding weights
    ~ N(0, std); LayerNorm gamma ~ 1 + N(0, 0.1), beta ~ N(0, 0.1); biases ~ N(0, std)."""
    g = torch.Generator().manual_seed(seed)
    for name in sorted(state_dict):
        t = state_dict[name]
        if not t.dtype.is_floating_point:
            continue
        lname = name.lower()
        if "layernorm" in lname or "layer_norm" in lname:
            if name.endswith("weight"):
                t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
            else:
                t.copy_(0.1 * torch.randn(t.shape, generator=g))
        else:
            t.copy_(std * torch.randn(t.shape, generator=g))
    return state_dict
