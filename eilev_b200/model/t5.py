"""Parameter holder with the key layout of HF ``T5ForConditionalGeneration`` (``shared``,
``encoder.block.N.layer.{0,1}.*``, ``decoder.block.N.layer.{0,1,2}.*``, ``lm_head``;
HF:t5/modeling_t5.py) so flan-T5 / ``eilev-blip2-flan-t5-xl`` checkpoints load unchanged.  The
arithmetic lives in ``eilev_b200/engine/t5.py``."""
from __future__ import annotations

import torch
import torch.nn as nn
from transformers import PreTrainedModel, T5Config
from transformers import initialization as hf_init
from transformers.modeling_outputs import Seq2SeqLMOutput

from ..engine import t5 as E_t5
from ..engine.packing import PackCache


class _T5Norm(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _T5Attention(nn.Module):
    def __init__(self, cfg: T5Config, has_bias: bool) -> None:
        super().__init__()
        inner = cfg.num_heads * cfg.d_kv
        self.q = nn.Linear(cfg.d_model, inner, bias=False)
        self.k = nn.Linear(cfg.d_model, inner, bias=False)
        self.v = nn.Linear(cfg.d_model, inner, bias=False)
        self.o = nn.Linear(inner, cfg.d_model, bias=False)
        if has_bias:
            self.relative_attention_bias = nn.Embedding(cfg.relative_attention_num_buckets, cfg.num_heads)


class _T5Dense(nn.Module):
    """T5DenseGatedActDense (wi_0, wi_1, wo: flan-T5 / T5 v1.1) or T5DenseActDense (wi, wo: the
    original T5, which is what ``T5Config()`` defaults to and the reference's tests build)."""

    def __init__(self, cfg: T5Config) -> None:
        super().__init__()
        if cfg.is_gated_act:
            self.wi_0 = nn.Linear(cfg.d_model, cfg.d_ff, bias=False)
            self.wi_1 = nn.Linear(cfg.d_model, cfg.d_ff, bias=False)
        else:
            self.wi = nn.Linear(cfg.d_model, cfg.d_ff, bias=False)
        self.wo = nn.Linear(cfg.d_ff, cfg.d_model, bias=False)


class _T5SelfAttnLayer(nn.Module):
    def __init__(self, cfg: T5Config, has_bias: bool) -> None:
        super().__init__()
        self.SelfAttention = _T5Attention(cfg, has_bias)
        self.layer_norm = _T5Norm(cfg.d_model)


class _T5CrossAttnLayer(nn.Module):
    def __init__(self, cfg: T5Config) -> None:
        super().__init__()
        self.EncDecAttention = _T5Attention(cfg, False)
        self.layer_norm = _T5Norm(cfg.d_model)


class _T5FFLayer(nn.Module):
    def __init__(self, cfg: T5Config) -> None:
        super().__init__()
        self.DenseReluDense = _T5Dense(cfg)
        self.layer_norm = _T5Norm(cfg.d_model)


class _T5Block(nn.Module):
    def __init__(self, cfg: T5Config, decoder: bool, has_bias: bool) -> None:
        super().__init__()
        layers: list[nn.Module] = [_T5SelfAttnLayer(cfg, has_bias)]
        if decoder:
            layers.append(_T5CrossAttnLayer(cfg))
        layers.append(_T5FFLayer(cfg))
        self.layer = nn.ModuleList(layers)


class _T5Stack(nn.Module):
    def __init__(self, cfg: T5Config, decoder: bool, n_layers: int) -> None:
        super().__init__()
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.d_model)
        self.block = nn.ModuleList([_T5Block(cfg, decoder, i == 0) for i in range(n_layers)])
        self.final_layer_norm = _T5Norm(cfg.d_model)


class T5ForConditionalGeneration(PreTrainedModel):
    config_class = T5Config
    config: T5Config
    base_model_prefix = "transformer"
    _tied_weights_keys = {"encoder.embed_tokens.weight": "shared.weight",
                          "decoder.embed_tokens.weight": "shared.weight"}
    _no_split_modules = ["_T5Block"]

    def __init__(self, cfg: T5Config) -> None:
        super().__init__(cfg)
        E_t5._check_cfg(cfg)
        self.shared = nn.Embedding(cfg.vocab_size, cfg.d_model)
        self.encoder = _T5Stack(cfg, False, cfg.num_layers)
        self.decoder = _T5Stack(cfg, True, cfg.num_decoder_layers)
        self.lm_head = nn.Linear(cfg.d_model, cfg.vocab_size, bias=False)
        # T5 v1.0 ties the head to the embedding (and scales the decoder output by d_model**-0.5);
        # flan-T5 / v1.1 configs carry tie_word_embeddings=False and keep their own head.  That
        # decision is read through scale_decoder_outputs: transformers 5.x's T5Config forces
        # tie_word_embeddings to True and records the configured value there (configuration_t5.py:77-83;
        # it then leaves a head that IS present in a checkpoint alone), 4.33.1 — the reference's pin —
        # simply honours tie_word_embeddings.
        if cfg.tie_word_embeddings and E_t5.scale_decoder_outputs(cfg):
            self._tied_weights_keys = {"lm_head.weight": "shared.weight", **type(self)._tied_weights_keys}
        self._pack = PackCache()
        self.post_init()

    def _init_weights(self, module) -> None:
        std = float(getattr(self.config, "initializer_factor", 1.0))
        with torch.no_grad():
            if isinstance(module, (nn.Linear, nn.Embedding)):
                hf_init.normal_(module.weight, mean=0.0, std=0.02 * std)
            elif isinstance(module, _T5Norm):
                hf_init.ones_(module.weight)

    def get_input_embeddings(self) -> nn.Embedding:
        return self.shared

    def get_output_embeddings(self) -> nn.Linear:
        return self.lm_head

    def forward(self, input_ids=None, attention_mask=None, labels=None, decoder_input_ids=None,
                inputs_embeds=None, **_):
        if inputs_embeds is not None or input_ids is None:
            raise NotImplementedError(
                "the B200 language model consumes input_ids (+ spliced video features); "
                "call VideoBlipForConditionalGeneration.forward")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        with torch.no_grad():
            out = E_t5.t5_forward(self, self._pack, input_ids, attention_mask, None, None, labels=labels,
                                  decoder_input_ids=decoder_input_ids)
        return Seq2SeqLMOutput(loss=out["loss"], logits=out["logits"],
                               encoder_last_hidden_state=out["encoder_last_hidden_state"])
