"""EVA-ViT-g tower on sm_100a kernels.

Restates VideoBlipVisionModel.forward (eilev/model/v2.py:24-103) + Blip2VisionModel
(HF:blip_2/modeling_blip_2.py:243-255, :319-402, :506-531) as a launch sequence:

    patch_gather -> GEMM(+bias +pos, CLS-slot row remap) -> 39 x [LN, QKV GEMM, attention,
    proj GEMM(+residual), LN, fc1 GEMM(+GELU), fc2 GEMM(+residual)] -> post LN (+ LN of CLS)

The ViT is frozen in the reference recipe (scripts/general/train_v2.py:124-125) and has no
autograd graph, so nothing is saved for backward and residual GEMMs run in place.
"""
from __future__ import annotations

import torch

from .. import ops
from .packing import PackCache, bf16, f32, ln_fold, pad_k


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def fold_layernorm_enabled() -> bool:
    """VB_VIT_LN_FOLD=0 keeps the stand-alone LayerNorm kernel in the encoder layers (A/B measurements)."""
    import os
    return os.environ.get("VB_VIT_LN_FOLD", "1") != "0"


def pack_vision(model, cache: PackCache):
    emb = model.embeddings
    params = [p for p in model.parameters()]

    def build():
        cfg = model.config
        k = 3 * cfg.patch_size * cfg.patch_size
        kpad = _round_up(k, 64) if k >= 64 else _round_up(k, 8)
        w = {
            "kpad": kpad,
            "patch_w": pad_k(emb.patch_embedding.weight.reshape(cfg.hidden_size, k), kpad),
            "patch_b": f32(emb.patch_embedding.bias),
            "cls": bf16(emb.class_embedding.reshape(-1)),
            "pos": bf16(emb.position_embedding.reshape(-1, cfg.hidden_size)),
            "post_g": f32(model.post_layernorm.weight),
            "post_b": f32(model.post_layernorm.bias),
            "layers": [],
        }
        # LayerNorm folded into the consuming GEMM (vb_gemm_args.ln_stats): layer_norm1 -> qkv and
        # layer_norm2 -> fc1 never write the normalised activations; the shapes must take the tcgen05 path
        w["fold"] = fold_layernorm_enabled() and cfg.hidden_size % 8 == 0 and cfg.intermediate_size % 8 == 0
        for layer in model.encoder.layers:
            lw = dict(
                proj_w=bf16(layer.self_attn.projection.weight), proj_b=f32(layer.self_attn.projection.bias),
                fc2_w=bf16(layer.mlp.fc2.weight), fc2_b=f32(layer.mlp.fc2.bias),
            )
            if w["fold"]:
                lw["qkv_w"], lw["qkv_b"], lw["qkv_cs"] = ln_fold(layer.self_attn.qkv.weight, layer.self_attn.qkv.bias,
                                                                 layer.layer_norm1.weight, layer.layer_norm1.bias)
                lw["fc1_w"], lw["fc1_b"], lw["fc1_cs"] = ln_fold(layer.mlp.fc1.weight, layer.mlp.fc1.bias,
                                                                 layer.layer_norm2.weight, layer.layer_norm2.bias)
            else:
                lw.update(
                    ln1_g=f32(layer.layer_norm1.weight), ln1_b=f32(layer.layer_norm1.bias),
                    qkv_w=bf16(layer.self_attn.qkv.weight),
                    qkv_b=None if layer.self_attn.qkv.bias is None else f32(layer.self_attn.qkv.bias),
                    ln2_g=f32(layer.layer_norm2.weight), ln2_b=f32(layer.layer_norm2.bias),
                    fc1_w=bf16(layer.mlp.fc1.weight), fc1_b=f32(layer.mlp.fc1.bias))
            w["layers"].append(lw)
        return w

    return cache.get("vision", params, build)


def vision_forward(model, cache: PackCache, pixel_values: torch.Tensor,
                   output_hidden_states: bool = False, max_frames: int = 192, output_attentions: bool = False):
    """pixel_values (N, C, T, H, W) on the GPU -> (last_hidden (N*T, S, D) bf16,
    pooled (N*T, D) bf16, hidden_states list|None[, attentions list of (N*T, H, S, S) f32]).  Frames are
    processed in chunks of `max_frames` to bound activation memory for large eval batches.
    output_attentions: the per-layer attention maps (v2.py:87-95) are computed by a separate plain kernel
    from the same q / k the fused attention consumes; the fused path itself is unchanged."""
    cfg = model.config
    w = pack_vision(model, cache)
    nv, c, t, h, wd = pixel_values.shape
    p = cfg.patch_size
    gh, gw = h // p, wd // p
    tokens = gh * gw + 1
    if tokens != w["pos"].shape[0]:
        raise ValueError(
            f"pixel_values give {tokens - 1} patches but the position table has {w['pos'].shape[0] - 1}; "
            "position interpolation is not supported")
    dim = cfg.hidden_size
    heads = cfg.num_attention_heads
    scale = (dim // heads) ** -0.5
    act = {"gelu": ops.EPI_GELU, "relu": ops.EPI_RELU}.get(cfg.hidden_act)
    if act is None:
        raise ValueError(f"unsupported vision hidden_act {cfg.hidden_act!r}")
    eps = cfg.layer_norm_eps
    frames = nv * t
    pixel_values = pixel_values.contiguous()
    # decoded uint8 frames: the image processor's rescale + normalize run inside the patch gather
    # (eilev/model/utils.py:5-26 -> BlipImageProcessor); float inputs are already normalised
    frame_norm = None
    if pixel_values.dtype == torch.uint8:
        frame_norm = (float(model.rescale_factor), tuple(model.image_mean), tuple(model.image_std))
    elif pixel_values.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        pixel_values = pixel_values.float()

    last = torch.empty((frames, tokens, dim), dtype=torch.bfloat16, device=pixel_values.device)
    pooled = torch.empty((frames, dim), dtype=torch.bfloat16, device=pixel_values.device)
    all_hidden = [] if output_hidden_states else None
    all_attn = [] if output_attentions else None
    if output_hidden_states or output_attentions:
        max_frames = frames  # keep layer outputs aligned
    # whole clips per chunk so patch_gather can address (clip, t) directly
    clips_per = max(1, max_frames // t)
    for v0 in range(0, nv, clips_per):
        v1 = min(nv, v0 + clips_per)
        f0, f1 = v0 * t, v1 * t
        nf = f1 - f0
        if frame_norm is not None:
            patches = ops.patch_gather_u8(pixel_values[v0:v1], p, w["kpad"], *frame_norm)
        else:
            patches = ops.patch_gather(pixel_values[v0:v1], p, w["kpad"])
        hidden = torch.empty((nf, tokens, dim), dtype=torch.bfloat16, device=pixel_values.device)
        hid2 = hidden.view(nf * tokens, dim)
        ops.cls_rows(w["cls"], w["pos"], hidden)
        ops.gemm(patches, w["patch_w"], w["patch_b"], residual=w["pos"], out=hid2, row_group=gh * gw)
        del patches
        if output_hidden_states:
            all_hidden.append(hidden.clone())
        if w["fold"]:
            # row [sum, sum of squares] of the residual stream: st1 feeds layer_norm1 (written by the
            # previous layer's fc2 epilogue; by vb_row_stats for the embeddings), st2 feeds layer_norm2
            # (written by the attention projection's epilogue)
            # (each producer clears the other buffer in its epilogue: no memset launches in the loop)
            st1 = ops.row_stats(hid2)
            st2 = torch.zeros_like(st1)
        for lw in w["layers"]:
            if w["fold"]:
                qkv = ops.gemm(hid2, lw["qkv_w"], lw["qkv_b"], ln_fold=(st1, lw["qkv_cs"], eps)).view(nf, tokens, 3 * dim)
                if output_attentions:
                    all_attn.append(ops.attention_probs(qkv[:, :, :dim], qkv[:, :, dim:2 * dim], heads, scale))
                o = ops.attention(qkv[:, :, :dim], qkv[:, :, dim:2 * dim], qkv[:, :, 2 * dim:], heads, scale)
                ops.gemm(o.view(nf * tokens, dim), lw["proj_w"], lw["proj_b"], residual=hid2, out=hid2,
                         stats_out=st2, stats_zero=st1)  # qkv has consumed st1
                h1 = ops.gemm(hid2, lw["fc1_w"], lw["fc1_b"], epilogue=act, ln_fold=(st2, lw["fc1_cs"], eps))
                ops.gemm(h1, lw["fc2_w"], lw["fc2_b"], residual=hid2, out=hid2,
                         stats_out=st1, stats_zero=st2)  # fc1 has consumed st2
                del qkv, o, h1
            else:
                y = ops.layernorm(hid2, lw["ln1_g"], lw["ln1_b"], eps)
                qkv = ops.gemm(y, lw["qkv_w"], lw["qkv_b"]).view(nf, tokens, 3 * dim)
                if output_attentions:
                    all_attn.append(ops.attention_probs(qkv[:, :, :dim], qkv[:, :, dim:2 * dim], heads, scale))
                o = ops.attention(qkv[:, :, :dim], qkv[:, :, dim:2 * dim], qkv[:, :, 2 * dim:], heads, scale)
                ops.gemm(o.view(nf * tokens, dim), lw["proj_w"], lw["proj_b"], residual=hid2, out=hid2)
                y = ops.layernorm(hid2, lw["ln2_g"], lw["ln2_b"], eps)
                h1 = ops.gemm(y, lw["fc1_w"], lw["fc1_b"], epilogue=act)
                ops.gemm(h1, lw["fc2_w"], lw["fc2_b"], residual=hid2, out=hid2)
                del y, qkv, o, h1
            if output_hidden_states:
                all_hidden.append(hidden.clone())
        ops.layernorm(hid2, w["post_g"], w["post_b"], eps, out=last[f0:f1].view(nf * tokens, dim))
        # pooler = post_layernorm applied a second time to the CLS row (HF :525-526)
        cls_rows = last[f0:f1, 0, :].contiguous()
        pooled[f0:f1] = ops.layernorm(cls_rows, w["post_g"], w["post_b"], eps)
    if output_attentions:
        return last, pooled, all_hidden, all_attn
    return last, pooled, all_hidden
