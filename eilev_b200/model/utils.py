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
            text: str | list[str] | None = None, normalize_on_device: bool = False) -> BatchEncoding:
    """:param video: (batch, channel, time, height, width) or (channel, time, height, width)
    :param normalize_on_device: (extension, default off = the reference's behaviour) keep the
        resized frames as uint8 — ``pixel_values`` is then a uint8 tensor, a quarter of the bytes
        to move to the GPU — and let the model apply the processor's rescale + normalize inside
        its patch-gather kernel (``VideoBlipVisionModel.set_frame_normalization``).
    """
    dims = None
    frames = None
    if video is not None:
        if video.dim() == 4:
            video = video[None]
        b, c, t = video.shape[:3]
        dims = (b, t, c)
        frames = video.transpose(1, 2).reshape(b * t, c, *video.shape[3:])
    if normalize_on_device and frames is not None:
        if frames.dtype != torch.uint8:
            raise ValueError("normalize_on_device=True expects decoded uint8 frames")
        inputs = processor(images=frames, text=text, return_tensors="pt", do_rescale=False, do_normalize=False)
        pv = inputs.pixel_values  # resized only: integral values in [0, 255]
        inputs["pixel_values"] = pv.round().clamp_(0, 255).to(torch.uint8)
    else:
        inputs = processor(images=frames, text=text, return_tensors="pt")
    if dims is not None:
        b, t, c = dims
        pv = inputs.pixel_values
        inputs["pixel_values"] = pv.view(b, t, c, pv.shape[-2], pv.shape[-1]).permute(0, 2, 1, 3, 4)
    return inputs
