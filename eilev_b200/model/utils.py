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


def process_on_device(processor, video: torch.Tensor, text: str | list[str] | None = None) -> BatchEncoding:
    """``process`` for decoded uint8 frames that already live on the GPU: the
    image processor's bicubic resize runs on the device, bit-exact with the PIL resize the reference's
    ``BlipImageProcessor`` performs (``vb_resize_u8_pass``), the frames stay uint8, and the model
    applies rescale + normalize inside its patch gather (``vb_patch_gather_u8``) — no float frame
    tensor is ever materialised.  Text goes through the processor's tokenizer as in ``process``.

    :param video: uint8 CUDA tensor (batch, channel, time, height, width) or (channel, time, height, width)
    :returns: BatchEncoding with uint8 ``pixel_values`` (batch, channel, time, size_h, size_w) on the
        device of ``video`` (+ the tokenizer's fields on the CPU when ``text`` is given)
    """
    from .. import ops

    if video.dtype != torch.uint8:
        raise ValueError("process_on_device expects decoded uint8 frames")
    if video.dim() == 4:
        video = video[None]
    ip = getattr(processor, "image_processor", processor)
    size = ip.size
    height, width = int(size["height"]), int(size["width"])
    if not getattr(ip, "do_resize", True):
        height, width = video.shape[-2:]
    inputs = processor(text=text, return_tensors="pt") if text is not None else BatchEncoding({})
    inputs["pixel_values"] = ops.resize_bicubic_u8(video, height, width)
    return inputs
