"""``process()`` — the Blip2Processor pixel/text contract of eilev/model/utils.py:5-26.

CPU pre-processing stays with HuggingFace's ``Blip2Processor`` (resize, rescale, CLIP
normalisation); this helper only folds the time axis into the batch for the image
processor and unfolds it again, so callers get ``pixel_values`` of shape
(batch, channel, time, height, width) exactly as with the reference.
"""
from __future__ import annotations

import torch
from transformers import BatchEncoding


def process(processor, video: torch.Tensor | None = None,
            text: str | list[str] | None = None) -> BatchEncoding:
    """:param video: (batch, channel, time, height, width) or (channel, time, height, width)"""
    dims = None
    frames = None
    if video is not None:
        if video.dim() == 4:
            video = video[None]
        b, c, t = video.shape[:3]
        dims = (b, t, c)
        frames = video.transpose(1, 2).reshape(b * t, c, *video.shape[3:])
    inputs = processor(images=frames, text=text, return_tensors="pt")
    if dims is not None:
        b, t, c = dims
        pv = inputs.pixel_values
        inputs["pixel_values"] = pv.view(b, t, c, pv.shape[-2], pv.shape[-1]).permute(0, 2, 1, 3, 4)
    return inputs
