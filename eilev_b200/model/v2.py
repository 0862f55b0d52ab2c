"""VideoBLIP (EILEV) on B200 — drop-in for ``eilev.model.v2`` of yukw777/EILEV.

Same class names, constructor, parameter names (stock BLIP-2 checkpoint keys), method
signatures, argument meaning and error behaviour as the reference
(eilev/model/v2.py:20-103 VideoBlipVisionModel, :106-501 VideoBlipForConditionalGeneration);
``from_pretrained`` / ``save_pretrained`` / ``config`` come from HuggingFace's
``PreTrainedModel`` exactly as they do for the reference.  The modules below only HOLD the
parameters under the reference's names — every FLOP of ``forward`` / ``generate`` runs in
the hand-written sm_100a kernels of ``libvideoblip_b200.so`` (eilev_b200/engine/*).  There is
no CPU or library fallback: calling ``forward`` without a CUDA device raises.
"""
from __future__ import annotations


import torch
import torch.nn as nn
from transformers import Blip2Config, Blip2VisionConfig, OPTConfig, PreTrainedModel
from transformers import initialization as hf_init
from transformers.modeling_outputs import (
    BaseModelOutputWithPooling,
    BaseModelOutputWithPoolingAndCrossAttentions,
    CausalLMOutputWithPast,
    Seq2SeqLMOutput,
)
from transformers.models.blip_2.modeling_blip_2 import Blip2ForConditionalGenerationModelOutput

from .. import _lib
from ..engine import opt as E_opt
from ..engine import t5 as E_t5
from ..engine import qformer as E_qf
from ..engine import vision as E_vis
from ..engine.packing import PackCache
from . import generation
from .t5 import T5ForConditionalGeneration


# =============================================================================== containers
# Parameter holders mirroring the HuggingFace module tree (names = checkpoint keys).
class _VisionEmbeddings(nn.Module):  # HF:blip_2/modeling_blip_2.py:184-201
    def __init__(self, cfg: Blip2VisionConfig) -> None:
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(1, 1, cfg.hidden_size))
        self.patch_embedding = nn.Conv2d(3, cfg.hidden_size, kernel_size=cfg.patch_size, stride=cfg.patch_size)
        n_pos = (cfg.image_size // cfg.patch_size) ** 2 + 1
        self.position_embedding = nn.Parameter(torch.randn(1, n_pos, cfg.hidden_size))


class _VisionAttention(nn.Module):  # :285-312 (fused qkv; bias = (q_bias, 0, v_bias))
    def __init__(self, cfg: Blip2VisionConfig) -> None:
        super().__init__()
        self.qkv = nn.Linear(cfg.hidden_size, 3 * cfg.hidden_size, bias=bool(cfg.qkv_bias))
        self.projection = nn.Linear(cfg.hidden_size, cfg.hidden_size)


class _VisionMLP(nn.Module):
    def __init__(self, cfg: Blip2VisionConfig) -> None:
        super().__init__()
        self.fc1 = nn.Linear(cfg.hidden_size, cfg.intermediate_size)
        self.fc2 = nn.Linear(cfg.intermediate_size, cfg.hidden_size)


class _VisionLayer(nn.Module):
    def __init__(self, cfg: Blip2VisionConfig) -> None:
        super().__init__()
        self.self_attn = _VisionAttention(cfg)
        self.layer_norm1 = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.mlp = _VisionMLP(cfg)
        self.layer_norm2 = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class _VisionEncoder(nn.Module):
    def __init__(self, cfg: Blip2VisionConfig) -> None:
        super().__init__()
        self.layers = nn.ModuleList([_VisionLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class _QFSelfAttention(nn.Module):  # :536-559
    def __init__(self, cfg, kv_dim: int) -> None:
        super().__init__()
        self.query = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.key = nn.Linear(kv_dim, cfg.hidden_size)
        self.value = nn.Linear(kv_dim, cfg.hidden_size)


class _QFDenseLN(nn.Module):  # SelfOutput :637-648 / Output :693-704
    def __init__(self, in_dim: int, cfg) -> None:
        super().__init__()
        self.dense = nn.Linear(in_dim, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class _QFAttention(nn.Module):
    def __init__(self, cfg, cross: bool) -> None:
        super().__init__()
        self.attention = _QFSelfAttention(cfg, cfg.encoder_hidden_size if cross else cfg.hidden_size)
        self.output = _QFDenseLN(cfg.hidden_size, cfg)


class _QFIntermediate(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        self.dense = nn.Linear(cfg.hidden_size, cfg.intermediate_size)


class _QFLayer(nn.Module):  # :707-727
    def __init__(self, cfg, idx: int) -> None:
        super().__init__()
        self.attention = _QFAttention(cfg, cross=False)
        self.has_cross_attention = idx % cfg.cross_attention_frequency == 0
        if self.has_cross_attention:
            self.crossattention = _QFAttention(cfg, cross=True)
        if getattr(cfg, "use_qformer_text_input", False):  # text branch: held, never executed here
            self.intermediate = _QFIntermediate(cfg)
            self.output = _QFDenseLN(cfg.intermediate_size, cfg)
        self.intermediate_query = _QFIntermediate(cfg)
        self.output_query = _QFDenseLN(cfg.intermediate_size, cfg)


class _QFEncoder(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        self.layer = nn.ModuleList([_QFLayer(cfg, i) for i in range(cfg.num_hidden_layers)])


class Blip2QFormerModel(nn.Module):  # parameter holder for HF Blip2QFormerModel :872-908
    def __init__(self, cfg) -> None:
        super().__init__()
        if isinstance(cfg.hidden_act, str) and cfg.hidden_act not in ("gelu", "relu"):
            raise NotImplementedError(f"Q-Former hidden_act {cfg.hidden_act!r} is not supported")
        self.config = cfg
        self.layernorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.encoder = _QFEncoder(cfg)


class _OPTAttention(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        d, b = cfg.hidden_size, cfg.enable_bias
        self.k_proj = nn.Linear(d, d, bias=b)
        self.v_proj = nn.Linear(d, d, bias=b)
        self.q_proj = nn.Linear(d, d, bias=b)
        self.out_proj = nn.Linear(d, d, bias=b)


class _OPTLayer(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        d = cfg.hidden_size
        self.self_attn = _OPTAttention(cfg)
        self.self_attn_layer_norm = nn.LayerNorm(d, elementwise_affine=cfg.layer_norm_elementwise_affine)
        self.fc1 = nn.Linear(d, cfg.ffn_dim, bias=cfg.enable_bias)
        self.fc2 = nn.Linear(cfg.ffn_dim, d, bias=cfg.enable_bias)
        self.final_layer_norm = nn.LayerNorm(d, elementwise_affine=cfg.layer_norm_elementwise_affine)


class _OPTDecoder(nn.Module):  # HF:opt/modeling_opt.py:272-318
    def __init__(self, cfg) -> None:
        super().__init__()
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.word_embed_proj_dim, cfg.pad_token_id)
        self.embed_positions = nn.Embedding(cfg.max_position_embeddings + 2, cfg.hidden_size)  # offset 2 (:45-70)
        self.final_layer_norm = nn.LayerNorm(cfg.hidden_size, elementwise_affine=cfg.layer_norm_elementwise_affine)
        self.layers = nn.ModuleList([_OPTLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class _OPTModel(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        self.decoder = _OPTDecoder(cfg)


class OPTForCausalLM(PreTrainedModel):
    """Parameter holder with the key layout of HF OPTForCausalLM (``model.decoder.*``,
    ``lm_head`` tied to ``embed_tokens``, HF:opt/modeling_opt.py:399-410); the text-only
    ``forward`` runs on the same engine."""

    config_class = OPTConfig
    config: OPTConfig
    base_model_prefix = "model"
    _tied_weights_keys = {"lm_head.weight": "model.decoder.embed_tokens.weight"}
    _no_split_modules = ["_OPTLayer"]

    def __init__(self, cfg: OPTConfig) -> None:
        super().__init__(cfg)
        E_opt._check_cfg(cfg)
        if not cfg.layer_norm_elementwise_affine:
            raise NotImplementedError("OPT without LayerNorm affine parameters is not supported")
        self.model = _OPTModel(cfg)
        self.lm_head = nn.Linear(cfg.word_embed_proj_dim, cfg.vocab_size, bias=False)
        self._pack = PackCache()
        self.post_init()

    def _init_weights(self, module) -> None:
        _init_weights(self, module)

    def get_input_embeddings(self) -> nn.Embedding:
        return self.model.decoder.embed_tokens

    def get_output_embeddings(self) -> nn.Linear:
        return self.lm_head

    def forward(self, input_ids=None, attention_mask=None, labels=None, inputs_embeds=None, **_):
        if inputs_embeds is not None or input_ids is None:
            raise NotImplementedError(
                "the B200 language model consumes input_ids (+ spliced video features); "
                "call VideoBlipForConditionalGeneration.forward")
        with torch.no_grad():
            out = E_opt.opt_forward(self, self._pack, input_ids, attention_mask, None, None, labels=labels)
        return CausalLMOutputWithPast(loss=out["loss"], logits=out["logits"])


# =============================================================================== vision model
def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.VbError(
            f"{what}: tensors must live on a CUDA device — eilev_b200 runs only on its sm_100a kernels "
            "(no CPU fallback). Move the model and inputs with .to('cuda').")


class VideoBlipVisionModel(PreTrainedModel):
    """Drop-in for eilev.model.v2.VideoBlipVisionModel (v2.py:20-103)."""

    config_class = Blip2VisionConfig
    config: Blip2VisionConfig
    main_input_name = "pixel_values"
    base_model_prefix = "blip"
    _no_split_modules = ["_VisionLayer"]

    # Normalisation applied on the device when ``pixel_values`` are decoded uint8 frames (SURVEY §8f
    # rank 3): BlipImageProcessor's defaults — rescale 1/255, OPENAI_CLIP mean / std
    # (HF:models/blip/image_processing_blip.py; eilev/model/utils.py:5-26).  Float inputs are taken
    # as already processed, exactly as the reference takes them.
    rescale_factor: float = 1 / 255
    image_mean: tuple = (0.48145466, 0.4578275, 0.40821073)
    image_std: tuple = (0.26862954, 0.26130258, 0.27577711)

    def set_frame_normalization(self, image_processor) -> None:
        """Adopt ``rescale_factor`` / ``image_mean`` / ``image_std`` of a HF image processor
        (``processor.image_processor``) for the uint8 frame path."""
        self.rescale_factor = float(image_processor.rescale_factor)
        self.image_mean = tuple(float(v) for v in image_processor.image_mean)
        self.image_std = tuple(float(v) for v in image_processor.image_std)

    def __init__(self, config: Blip2VisionConfig) -> None:
        super().__init__(config)
        self.embeddings = _VisionEmbeddings(config)
        self.encoder = _VisionEncoder(config)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self._pack = PackCache()
        self.post_init()

    def _init_weights(self, module) -> None:
        _init_weights(self, module)

    def get_input_embeddings(self):
        return self.embeddings

    def forward(
        self,
        pixel_values: torch.FloatTensor | None = None,
        output_attentions: bool | None = None,
        output_hidden_states: bool | None = None,
        return_dict: bool | None = None,
    ) -> tuple | BaseModelOutputWithPooling:
        """pixel_values (num_videos, channel, time, height, width) ->
        last_hidden_state (num_videos, time*seq_len, hidden), pooler_output (num_videos, time, hidden),
        hidden_states tuple of (num_videos, time*seq_len, hidden) — v2.py:31-49."""
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")  # v2.py:50-51
        _require_cuda(pixel_values, "VideoBlipVisionModel.forward")
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        num_videos, _, time, _, _ = pixel_values.size()
        attentions = None
        with torch.no_grad():  # frozen tower: never builds an autograd graph (train_v2.py:124-125)
            if output_attentions:
                # v2.py:87-95: one (num_videos, time, heads, seq_len, seq_len) map per layer; computed by a
                # separate plain kernel next to the fused attention (which never materialises them)
                last, pooled, hidden, maps = E_vis.vision_forward(self, self._pack, pixel_values,
                                                                  bool(output_hidden_states), output_attentions=True)
                attentions = tuple(a.view(num_videos, time, a.shape[1], a.shape[2], a.shape[3]).to(self.dtype)
                                   for a in maps)
            else:
                last, pooled, hidden = E_vis.vision_forward(self, self._pack, pixel_values,
                                                            bool(output_hidden_states))
        seq_len = last.size(1)
        dt = self.dtype
        last_hidden_state = last.view(num_videos, time * seq_len, -1).to(dt)
        pooler_output = pooled.view(num_videos, time, -1).to(dt)
        hidden_states = None
        if hidden is not None:
            hidden_states = tuple(h.view(num_videos, time * seq_len, -1).to(dt) for h in hidden)
        if return_dict:
            return BaseModelOutputWithPooling(last_hidden_state=last_hidden_state,
                                              pooler_output=pooler_output,
                                              hidden_states=hidden_states, attentions=attentions)
        return (last_hidden_state, pooler_output, hidden_states, attentions)


def _init_weights(owner: PreTrainedModel, module: nn.Module) -> None:
    """Same distributions as Blip2PreTrainedModel._init_weights (HF :425-446)."""
    std = getattr(owner.config, "initializer_range", None)
    if std is None:
        std = getattr(owner.config, "init_std", 0.02)
    # hf_init.* skip tensors that were just loaded from a checkpoint
    with torch.no_grad():
        if isinstance(module, (nn.Linear, nn.Conv2d, nn.Embedding)):
            hf_init.normal_(module.weight, mean=0.0, std=std)
            if getattr(module, "bias", None) is not None:
                hf_init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            hf_init.ones_(module.weight)
            hf_init.zeros_(module.bias)
        elif isinstance(module, _VisionEmbeddings):
            hf_init.trunc_normal_(module.position_embedding, mean=0.0, std=std)
            hf_init.trunc_normal_(module.class_embedding, mean=0.0, std=std)
        elif isinstance(module, VideoBlipForConditionalGeneration):
            hf_init.zeros_(module.query_tokens)


# =============================================================================== autograd glue
class _QFormerProjectFn(torch.autograd.Function):
    """ViT output -> Q-Former -> language_projection with the hand-written backward."""

    @staticmethod
    def forward(ctx, model, image_embeds, seed, *params):
        feats, qout, saved = E_qf.qformer_forward(model, model._pack, image_embeds, save=True, seed=seed)
        ctx.model, ctx.saved = model, saved
        ctx.param_meta = [(p.dtype, p.requires_grad) for p in params]
        ctx.mark_non_differentiable(qout)
        return feats, qout

    @staticmethod
    def backward(ctx, d_feats, _d_qout):
        model = ctx.model
        # Opt-in (DataParallelTrainer sets model._grad_sink): gradients whose .grad is a
        # pre-allocated contiguous f32 buffer (the trainer's flat views) are accumulated in place by
        # the kernels that produce them.  Off by default: under torch DistributedDataParallel every
        # gradient has to come back through autograd so the reducer's hooks fire.
        sink = None
        if getattr(model, "_grad_sink", False):
            sink = {name: p.grad for name, p in E_qf.qformer_param_list(model)
                    if p.requires_grad and p.grad is not None and p.grad.dtype == torch.float32
                    and p.grad.is_contiguous() and p.grad.is_cuda}
        grads = E_qf.qformer_backward(model, model._pack, ctx.saved, d_feats.to(torch.bfloat16), sink)
        ctx.saved = None
        out = []
        for (name, _), (dtype, need) in zip(E_qf.qformer_param_list(model), ctx.param_meta):
            g = grads.get(name) if need else None
            out.append(None if g is None else g.to(dtype))
        return (None, None, None, *out)


class _LMLossFn(torch.autograd.Function):
    """Splice + OPT + shifted cross entropy; backward is dgrad-only down to the video slots."""

    @staticmethod
    def forward(ctx, model, video_features, input_ids, attention_mask, video_mask, labels, seed):
        lm = model.language_model
        out = E_opt.opt_forward(lm, lm._pack, input_ids, attention_mask, video_mask, video_features,
                                labels=labels, save=True, seed=seed)
        ctx.model, ctx.saved = model, out["ctx"]
        ctx.feat_dtype = video_features.dtype
        ctx.mark_non_differentiable(out["logits"], out["status"])
        return out["loss"], out["logits"], out["status"]

    @staticmethod
    def backward(ctx, grad_loss, _gl, _gs):
        lm = ctx.model.language_model
        d_feats = E_opt.opt_backward(lm, lm._pack, ctx.saved, grad_loss)
        ctx.saved = None
        return None, d_feats.to(ctx.feat_dtype), None, None, None, None, None


class _Seq2SeqLossFn(torch.autograd.Function):
    """Splice + flan-T5 encoder/decoder + cross entropy (v2.py:228-238); backward is dgrad-only
    through decoder, cross-attention K/V, encoder and splice down to the video slots."""

    @staticmethod
    def forward(ctx, model, video_features, input_ids, attention_mask, video_mask, labels, seed):
        lm = model.language_model
        out = E_t5.t5_forward(lm, lm._pack, input_ids, attention_mask, video_mask, video_features,
                              labels=labels, save=True, seed=seed)
        ctx.model, ctx.saved = model, out["ctx"]
        ctx.feat_dtype = video_features.dtype
        ctx.mark_non_differentiable(out["logits"], out["status"], out["encoder_last_hidden_state"])
        return out["loss"], out["logits"], out["status"], out["encoder_last_hidden_state"]

    @staticmethod
    def backward(ctx, grad_loss, _gl, _gs, _ge):
        lm = ctx.model.language_model
        d_feats = E_t5.t5_backward(lm, lm._pack, ctx.saved, grad_loss)
        ctx.saved = None
        return None, d_feats.to(ctx.feat_dtype), None, None, None, None, None


# =============================================================================== full model
class VideoBlipForConditionalGeneration(PreTrainedModel):
    """Drop-in for eilev.model.v2.VideoBlipForConditionalGeneration (v2.py:106-501)."""

    config_class = Blip2Config
    config: Blip2Config
    main_input_name = "pixel_values"
    base_model_prefix = "blip"
    _no_split_modules = ["_VisionLayer", "_QFLayer", "_OPTLayer"]
    _keep_in_fp32_modules: list = []

    def __init__(self, config: Blip2Config) -> None:
        super().__init__(config)
        self.vision_model = VideoBlipVisionModel(config.vision_config)
        self.query_tokens = nn.Parameter(
            torch.zeros(1, config.num_query_tokens, config.qformer_config.hidden_size))
        self.qformer = Blip2QFormerModel(config.qformer_config)
        self.language_projection = nn.Linear(config.qformer_config.hidden_size,
                                             config.text_config.hidden_size)
        if config.use_decoder_only_language_model:
            if config.text_config.model_type != "opt":
                raise NotImplementedError(
                    f"decoder-only language model {config.text_config.model_type!r}: only OPT is built")
            self.language_model = OPTForCausalLM(config.text_config)
        else:
            if config.text_config.model_type != "t5":
                raise NotImplementedError(
                    f"encoder-decoder language model {config.text_config.model_type!r}: only T5 is built")
            self.language_model = T5ForConditionalGeneration(config.text_config)
        self._pack = PackCache()
        self.post_init()

    def _init_weights(self, module) -> None:
        _init_weights(self, module)

    # ------------------------------------------------------------------ HF plumbing
    def get_input_embeddings(self) -> nn.Module:
        return self.language_model.get_input_embeddings()

    def set_input_embeddings(self, value) -> None:
        self.language_model.model.decoder.embed_tokens = value

    def get_output_embeddings(self) -> nn.Module:
        return self.language_model.get_output_embeddings()

    def enable_input_require_grads(self) -> None:
        """train_v2.py:130.  The reference needs the LM embedding output to require grad so
        autograd reaches the Q-Former through the frozen LM; here the hand-written backward
        always delivers d(video_features), so this only records the request."""
        self._input_require_grads = True

    def disable_input_require_grads(self) -> None:
        self._input_require_grads = False

    # ------------------------------------------------------------------ encode
    def _next_dropout_seed(self, device):
        """Device-resident dropout seed, advanced once per training forward (in place, so a
        captured CUDA graph draws fresh masks on every replay).  None in eval mode."""
        if not self.training:
            return None
        seed = getattr(self, "_dropout_seed", None)
        if seed is None or seed.device != device:
            seed = torch.full((1,), int(torch.initial_seed()) & 0x7FFFFFFF, dtype=torch.int64, device=device)
            self._dropout_seed = seed
        seed.add_(1 << 20)  # sites add salts < 2^20
        return seed

    def _video_features(self, pixel_values: torch.Tensor, output_hidden_states: bool, train: bool,
                        seed=None, output_attentions: bool = False):
        vis_last, vis_pooled, vis_hidden, vis_attn = None, None, None, None
        with torch.no_grad():
            res = E_vis.vision_forward(self.vision_model, self.vision_model._pack, pixel_values,
                                       output_hidden_states, output_attentions=output_attentions)
            vis_last, vis_pooled, vis_hidden = res[:3]
            vis_attn = res[3] if output_attentions else None
        n, _, t, _, _ = pixel_values.shape
        s = vis_last.size(1)
        image_embeds = vis_last.view(n, t * s, -1)
        if train:
            params = [p for _, p in E_qf.qformer_param_list(self)]
            feats, qout = _QFormerProjectFn.apply(self, image_embeds, seed, *params)
        else:
            with torch.no_grad():
                feats, qout, _ = E_qf.qformer_forward(self, self._pack, image_embeds, save=False)
        return feats, qout, (image_embeds, vis_pooled.view(n, t, -1), vis_hidden, n, t, s, vis_attn)

    def _pack_vision_outputs(self, vis, return_dict: bool):
        image_embeds, pooled, hidden, n, t, s, maps = vis
        dt = self.dtype
        hs = None if hidden is None else tuple(h.view(n, t * s, -1).to(dt) for h in hidden)
        at = None if maps is None else tuple(a.view(n, t, a.shape[1], s, s).to(dt) for a in maps)
        if return_dict:
            return BaseModelOutputWithPooling(last_hidden_state=image_embeds.to(dt),
                                              pooler_output=pooled.to(dt), hidden_states=hs, attentions=at)
        return (image_embeds.to(dt), pooled.to(dt), hs, at)

    # ------------------------------------------------------------------ forward
    def forward(
        self,
        input_ids: torch.Tensor,
        attention_mask: torch.Tensor | None = None,
        pixel_values: torch.Tensor | None = None,
        video_input_mask: torch.Tensor | None = None,
        decoder_input_ids: torch.Tensor | None = None,
        decoder_attention_mask: torch.Tensor | None = None,
        output_attentions: bool | None = None,
        output_hidden_states: bool | None = None,
        labels: torch.Tensor | None = None,
        return_dict: bool | None = None,
    ) -> tuple | Blip2ForConditionalGenerationModelOutput:
        """Same contract as v2.py:132-252.

        :param pixel_values: (num_videos, channel, time, height, width)
        :param video_input_mask: (batch, seq_len); its ones (row-major) receive the
            num_videos*num_query_tokens projected Q-Former rows in (clip, query) order.
        """
        if pixel_values is not None:
            assert video_input_mask is not None  # v2.py:154-157
            video_input_mask = video_input_mask.bool()
        else:
            video_input_mask = None  # the reference ignores the mask without pixel values (v2.py:205-213)
        # output_attentions: the vision tower's maps are returned in vision_outputs.attentions (v2.py:169-177);
        # the Q-Former's and the LM's fused attention kernels produce none (their `attentions` stay None)
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        _require_cuda(input_ids, "VideoBlipForConditionalGeneration.forward")
        want_hidden = bool(output_hidden_states)

        trainable = any(p.requires_grad for _, p in E_qf.qformer_param_list(self))
        train = torch.is_grad_enabled() and trainable and pixel_values is not None and labels is not None
        if torch.is_grad_enabled() and not getattr(self, "_warned_frozen_towers", False):
            # The reference's recipe freezes the ViT and the LM (scripts/general/train_v2.py:123-130) and this engine
            # differentiates only what that recipe trains: say so once instead of silently returning no gradient.
            loose = [n for n, p in self.named_parameters()
                     if p.requires_grad and (n.startswith("vision_model.") or n.startswith("language_model."))]
            if loose or (trainable and labels is None):
                import warnings
                warnings.warn(
                    "eilev_b200 computes gradients for the Q-Former, query_tokens and language_projection only, and only "
                    "through the returned loss (labels=...): "
                    + (f"{len(loose)} vision_model / language_model parameters have requires_grad=True and will get no "
                       f"gradient (first: {loose[0]}); " if loose else "")
                    + ("labels is None, so the logits carry no graph; " if trainable and labels is None else "")
                    + "freeze the towers as scripts/general/train_v2.py does (eilev_b200.train.freeze_for_recipe).",
                    stacklevel=2)
            self._warned_frozen_towers = True   # checked once, on the first grad-enabled call

        vision_outputs = None
        query_outputs = None
        feats = None
        seed = self._next_dropout_seed(input_ids.device) if train else None
        if pixel_values is not None:
            _require_cuda(pixel_values, "VideoBlipForConditionalGeneration.forward")
            feats, qout, vis = self._video_features(pixel_values, want_hidden, train, seed,
                                                    output_attentions=bool(output_attentions))
            vision_outputs = self._pack_vision_outputs(vis, return_dict)
            q = qout.to(self.dtype)
            query_outputs = (BaseModelOutputWithPoolingAndCrossAttentions(last_hidden_state=q, pooler_output=q[:, 0])
                             if return_dict else (q, q[:, 0]))
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)  # v2.py:216-217

        lm = self.language_model
        if not self.config.use_decoder_only_language_model:
            return self._forward_seq2seq(input_ids, attention_mask, video_input_mask, feats, labels,
                                         decoder_input_ids, decoder_attention_mask, want_hidden, train,
                                         vision_outputs, query_outputs, return_dict, seed)
        if train:
            loss, logits, status = _LMLossFn.apply(self, feats, input_ids, attention_mask,
                                                   video_input_mask, labels, seed)
            hidden_states = None
        else:
            with torch.no_grad():
                out = E_opt.opt_forward(lm, lm._pack, input_ids, attention_mask, video_input_mask, feats,
                                        labels=labels, output_hidden_states=want_hidden)
            loss, logits, status, hidden_states = out["loss"], out["logits"], out["status"], out["hidden_states"]
        self._last_splice_status = status  # device int32[2]; see check_splice()
        if hidden_states is not None:
            hidden_states = tuple(h.to(self.dtype) for h in hidden_states)
        lm_outputs = CausalLMOutputWithPast(loss=loss, logits=logits, hidden_states=hidden_states)
        if not return_dict:
            output = (logits, vision_outputs, query_outputs, tuple(v for v in (loss, logits) if v is not None))
            return ((loss,) + output) if loss is not None else output
        return Blip2ForConditionalGenerationModelOutput(
            loss=loss, logits=logits, vision_outputs=vision_outputs,
            qformer_outputs=query_outputs, language_model_outputs=lm_outputs)

    def _forward_seq2seq(self, input_ids, attention_mask, video_input_mask, feats, labels, decoder_input_ids,
                         decoder_attention_mask, want_hidden, train, vision_outputs, query_outputs, return_dict,
                         seed=None):
        """v2.py:228-238: the interleaved embeddings go to the T5 encoder, the labels
        (shifted right) to the decoder."""
        if want_hidden:
            raise NotImplementedError("output_hidden_states=True is not supported for the seq2seq LM")
        if decoder_attention_mask is not None and not bool(decoder_attention_mask.all()):
            raise NotImplementedError("decoder_attention_mask with zeros is not supported (labels use -100)")
        lm = self.language_model
        if train:
            loss, logits, status, enc = _Seq2SeqLossFn.apply(self, feats, input_ids, attention_mask,
                                                             video_input_mask, labels, seed)
        else:
            with torch.no_grad():
                out = E_t5.t5_forward(lm, lm._pack, input_ids, attention_mask, video_input_mask, feats,
                                      labels=labels, decoder_input_ids=decoder_input_ids)
            loss, logits, status, enc = out["loss"], out["logits"], out["status"], out["encoder_last_hidden_state"]
        self._last_splice_status = status
        lm_outputs = Seq2SeqLMOutput(loss=loss, logits=logits, encoder_last_hidden_state=enc.to(self.dtype))
        if not return_dict:
            output = (logits, vision_outputs, query_outputs, tuple(v for v in (loss, logits) if v is not None))
            return ((loss,) + output) if loss is not None else output
        return Blip2ForConditionalGenerationModelOutput(
            loss=loss, logits=logits, vision_outputs=vision_outputs,
            qformer_outputs=query_outputs, language_model_outputs=lm_outputs)

    def check_splice(self) -> None:
        """Raises like the reference's masked assignment (v2.py:210) if the number of ones in the
        last ``video_input_mask`` differed from num_videos*num_query_tokens.  Costs one
        device->host read, so it is not on the default path."""
        st = getattr(self, "_last_splice_status", None)
        if st is not None:
            bad, count = st.tolist()
            if bad:
                raise RuntimeError(
                    f"shape mismatch: video_input_mask selects {count} positions but the video "
                    "features have a different number of rows")

    # ------------------------------------------------------------------ generate
    @torch.no_grad()
    def generate(
        self,
        input_ids: torch.Tensor,
        pixel_values: torch.Tensor | None = None,
        video_input_mask: torch.Tensor | None = None,
        attention_mask: torch.Tensor | None = None,
        **generate_kwargs,
    ) -> torch.Tensor:
        """Same contract as v2.py:254-324: returns only the newly generated token ids
        (decoder-only LM fed with embeddings)."""
        assert not (input_ids is None and pixel_values is None)  # v2.py:271
        if pixel_values is not None:
            assert video_input_mask is not None  # v2.py:274
            video_input_mask = video_input_mask.bool()
        else:
            video_input_mask = None  # no features to splice: the mask is ignored as in v2.py:300-308
        _require_cuda(input_ids, "VideoBlipForConditionalGeneration.generate")
        feats = None
        if pixel_values is not None:
            _require_cuda(pixel_values, "VideoBlipForConditionalGeneration.generate")
            feats, _, _ = self._video_features(pixel_values, False, train=False)
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        out = generation.generate(self, input_ids, attention_mask, video_input_mask, feats,
                                  **generate_kwargs)
        self.check_splice()  # the decoding loop has already synchronised: the reference's shape error is free here
        return out

    # ------------------------------------------------------------------ classify
    @torch.no_grad()
    def classify(
        self,
        prompt_input_ids: torch.Tensor,
        class_input_ids: torch.Tensor,
        prompt_attention_mask: torch.Tensor | None = None,
        pixel_values: torch.Tensor | None = None,
        prompt_video_input_mask: torch.Tensor | None = None,
        class_attention_mask: torch.Tensor | None = None,
        class_batch_size: int | None = None,
    ) -> torch.Tensor:
        """Same contract as v2.py:326-424: mean log-likelihood (batch, num_classes) of every
        class continuation given the left-padded prompt (decoder-only LM only, :351)."""
        assert self.config.use_decoder_only_language_model  # v2.py:351
        if pixel_values is not None:
            assert prompt_video_input_mask is not None  # v2.py:355
            prompt_video_input_mask = prompt_video_input_mask.bool()
        _require_cuda(prompt_input_ids, "VideoBlipForConditionalGeneration.classify")
        feats = None
        if pixel_values is not None:
            _require_cuda(pixel_values, "VideoBlipForConditionalGeneration.classify")
            feats, _, _ = self._video_features(pixel_values, False, train=False)
        if prompt_attention_mask is None:
            prompt_attention_mask = torch.ones_like(prompt_input_ids)
        lm = self.language_model
        scores, status = E_opt.opt_classify(
            lm, lm._pack, prompt_input_ids, prompt_attention_mask,
            prompt_video_input_mask if pixel_values is not None else None, feats,
            class_input_ids, class_attention_mask, class_batch_size)
        self._last_splice_status = status
        self.check_splice()
        return scores  # f32 (the reference returns the LM dtype; f32 keeps the sums exact)
