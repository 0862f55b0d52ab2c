"""ORACLE — test infrastructure only.  CPU restatement of the floating-point part of the reference's training
video transform (scripts/general/train_v2.py:152-166) for one clip and GIVEN random draws:

    ConvertUint8ToFloat (x / 255) -> Normalize(mean, std) -> crop box -> torch.nn.functional.interpolate(
    size, mode="bicubic") (what pytorchvideo's RandomResizedCrop calls after cropping) -> horizontal flip

written with the very torch functions the reference's transform classes call, in the reference's order.
Parity status: pinned to torch itself (``F.interpolate``) — pytorchvideo is not installed in this image, so its
thin wrappers (crop, ``interpolate``, ``hflip``) are restated from their documentation; the box sampler is
checked draw for draw against torchvision's ``RandomResizedCrop.get_params`` (tests/test_preprocess_cpu.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def train_clip_transform(clip_u8: torch.Tensor, box, size, mean, std, flip: bool) -> torch.Tensor:
    """uint8 (C, T, H, W) -> float32 (C, T, size[0], size[1])."""
    x = clip_u8.float() / 255.0
    m = torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(-1, 1, 1, 1)
    x = (x - m) / s
    top, left, h, w = box
    x = x[:, :, top:top + h, left:left + w]
    x = F.interpolate(x, size=tuple(size), mode="bicubic")  # (C, T, H, W): C is the batch axis, T the channels
    if flip:
        x = x.flip(-1)
    return x
