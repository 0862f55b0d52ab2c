"""VideoBLIP v1 on B200 — drop-in for ``eilev.model.v1`` of yukw777/EILEV.

The reference's v1 (eilev/model/v1.py:95-119) only swaps the vision tower of HuggingFace's
``Blip2ForConditionalGeneration`` and inherits its ``forward`` / ``generate`` from the pinned
transformers 4.33.1: ONE video per batch row, whose ``num_query_tokens`` projected Q-Former
rows are *prepended* to the embedded prompt

    inputs_embeds  = cat([language_projection(qformer(vit(pixel_values))), embed(input_ids)], 1)
    attention_mask = cat([ones(batch, num_query_tokens), attention_mask], 1)

and, for a decoder-only LM, the loss is the shifted cross entropy over the LAST
``labels.size(1)`` logits (the returned ``logits`` are that slice too).

That is the v2 splice with a fixed layout, so v1 runs on the same sm_100a kernels: the rows
are laid out as ``[video slots | text]`` with a ``video_input_mask`` over the first
``num_query_tokens`` positions, and the v2 engine does the rest.  The layout helpers below are
pure integer bookkeeping (``prepend_video_slots``, ``compact_left``) and are tested on CPU;
``VideoBlipVisionModel`` is identical in v1 and v2 (v1.py:14-92 == v2.py:20-103).
"""
from __future__ import annotations

import torch
from transformers.models.blip_2.modeling_blip_2 import Blip2ForConditionalGenerationModelOutput

from . import v2
from .v2 import VideoBlipVisionModel  # noqa: F401  (same class: v1.py:14-92)


def prepend_video_slots(input_ids: torch.Tensor, attention_mask: torch.Tensor | None,
                        labels: torch.Tensor | None, num_query_tokens: int, fill_id: int,
                        decoder_only: bool):
    """(batch, L) prompt -> the (batch, Q+L) v2 layout of HF 4.33.1 ``Blip2ForConditionalGeneration.forward``.

    Returns ``(input_ids, attention_mask, video_input_mask, labels)``.  The Q leading slots hold
    ``fill_id`` (their embeddings are overwritten by the video features), are attended to
    (``language_model_attention_mask`` = ones) and never contribute to the loss.  For a
    decoder-only LM the reference shifts the labels against the last L logits only, so
    ``labels[:, 0]`` is never a target: it becomes -100 here, which makes the full-length shifted
    cross entropy of the v2 engine equal to the reference's sliced one.  Seq2seq labels belong
    to the decoder and pass through untouched.
    """
    b, _ = input_ids.shape
    q = int(num_query_tokens)
    dev = input_ids.device
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    ids = torch.cat([torch.full((b, q), int(fill_id), dtype=input_ids.dtype, device=dev), input_ids], dim=1)
    am = torch.cat([torch.ones((b, q), dtype=attention_mask.dtype, device=dev), attention_mask], dim=1)
    vm = torch.zeros_like(ids)
    vm[:, :q] = 1
    if labels is not None and decoder_only:
        labels = torch.cat([torch.full((b, q + 1), -100, dtype=labels.dtype, device=labels.device),
                            labels[:, 1:]], dim=1)
    return ids, am, vm, labels


def compact_left(input_ids: torch.Tensor, attention_mask: torch.Tensor, video_input_mask: torch.Tensor):
    """Moves every masked-out position of a row to its left end, keeping the attended positions in
    order (stable partition).  With the video slots in front, a left-padded prompt has its padding
    hole in the middle (``[video | pad | text]``); OPT's positions are a cumulative sum of the mask
    and masked keys contribute nothing, so ``[pad | video | text]`` yields the same next-token
    distribution — and is the layout the paged decode kernels read (one ``first_valid`` per row).
    """
    order = torch.sort((attention_mask != 0).to(torch.int8), dim=1, stable=True).indices
    return (input_ids.gather(1, order), attention_mask.gather(1, order), video_input_mask.gather(1, order))


class VideoBlipForConditionalGeneration(v2.VideoBlipForConditionalGeneration):
    """Drop-in for eilev.model.v1.VideoBlipForConditionalGeneration (v1.py:95-119): same module
    tree / checkpoint keys as v2, HF 4.33.1 ``Blip2ForConditionalGeneration`` call signatures."""

    def _fill_id(self) -> int:
        pad = getattr(self.config.text_config, "pad_token_id", None)
        return int(pad) if pad is not None else 0

    def forward(  # type: ignore[override]
        self,
        pixel_values: torch.Tensor,
        input_ids: torch.Tensor,
        attention_mask: torch.Tensor | None = None,
        decoder_input_ids: torch.Tensor | None = None,
        decoder_attention_mask: torch.Tensor | None = None,
        output_attentions: bool | None = None,
        output_hidden_states: bool | None = None,
        labels: torch.Tensor | None = None,
        return_dict: bool | None = None,
    ) -> tuple | Blip2ForConditionalGenerationModelOutput:
        """:param pixel_values: (batch, channel, time, height, width) — one video per row
        :param input_ids: (batch, L) prompt; the video tokens are prepended, not interleaved
        :param labels: decoder-only: (batch, L), aligned with ``input_ids``; seq2seq: decoder targets
        """
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")  # v1.py:40-41
        if pixel_values.shape[0] != input_ids.shape[0]:
            raise ValueError(f"v1 takes one video per row: {pixel_values.shape[0]} videos for "
                             f"{input_ids.shape[0]} prompts")
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        decoder_only = bool(self.config.use_decoder_only_language_model)
        ids, am, vm, full_labels = prepend_video_slots(
            input_ids, attention_mask, labels, self.config.num_query_tokens, self._fill_id(), decoder_only)
        out = super().forward(ids, attention_mask=am, pixel_values=pixel_values, video_input_mask=vm,
                              decoder_input_ids=decoder_input_ids, decoder_attention_mask=decoder_attention_mask,
                              output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                              labels=full_labels, return_dict=True)
        logits = out.logits
        if labels is not None and decoder_only:
            logits = logits[:, -labels.size(1):, :]  # HF 4.33.1: the returned logits are the text slice
        if not return_dict:
            lm_out = out.language_model_outputs
            output = (logits, out.vision_outputs.to_tuple(), out.qformer_outputs.to_tuple(), lm_out.to_tuple())
            return ((out.loss,) + output) if out.loss is not None else output
        return Blip2ForConditionalGenerationModelOutput(
            loss=out.loss, logits=logits, vision_outputs=out.vision_outputs,
            qformer_outputs=out.qformer_outputs, language_model_outputs=out.language_model_outputs)

    @torch.no_grad()
    def generate(  # type: ignore[override]
        self,
        pixel_values: torch.Tensor,
        input_ids: torch.Tensor | None = None,
        attention_mask: torch.Tensor | None = None,
        **generate_kwargs,
    ) -> torch.Tensor:
        """HF 4.33.1 ``Blip2ForConditionalGeneration.generate``: no prompt means ``[bos]`` per row;
        returns the new tokens only (decoder-only LM fed with embeddings) or the decoder sequence
        (seq2seq).  Called by samples/video_blip_generate_action_narration.py:24-32."""
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        batch = pixel_values.shape[0]
        if input_ids is None:
            bos = self.config.text_config.bos_token_id
            input_ids = torch.full((batch, 1), int(bos), dtype=torch.long, device=pixel_values.device)
        ids, am, vm, _ = prepend_video_slots(input_ids, attention_mask, None, self.config.num_query_tokens,
                                             self._fill_id(), True)
        if self.config.use_decoder_only_language_model:
            ids, am, vm = compact_left(ids, am, vm)
        return super().generate(ids, pixel_values=pixel_values, video_input_mask=vm, attention_mask=am,
                                **generate_kwargs)

    def classify(self, *args, **kwargs):  # v1 has no classify (v1.py:95-119)
        raise AttributeError("eilev.model.v1.VideoBlipForConditionalGeneration has no classify(); use eilev_b200.model.v2")
