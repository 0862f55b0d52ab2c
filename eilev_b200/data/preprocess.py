"""Training-time preprocessing of scripts/general/train_v2.py: the ``Preprocessor`` that turns a
datapoint of (video, narration) items into model inputs (:45-75) and the video transform stack (:143-167)

    UniformTemporalSubsample(T) -> RandAugment(magnitude 5) -> /255 -> Normalize -> RandomResizedCrop(224, 224,
    scale (0.5, 1), ratio (3/4, 4/3), bicubic) -> RandomHorizontalFlip

re-designed for the B200: the integer part (frame subsampling, the uint8 augmentation, the random draws of the
crop box and of the flip) stays on the host, where the dataloader workers are; the floating-point part — rescale,
normalise, crop, bicubic resize, flip — is ONE fused device pass over the uint8 clip
(``vb_crop_resize_normalize_u8``), so a clip crosses PCIe once, as uint8 (4x fewer bytes than the reference's
float frames), and no intermediate float clip is ever written.

pytorchvideo (the reference's transform library) is not a dependency: ``UniformTemporalSubsample`` and the
crop-box sampling restate its documented algorithms (the box sampler is torchvision's
``RandomResizedCrop.get_params``, which pytorchvideo's ``_get_param_spatial_crop`` follows, log-uniform ratio,
10 tries, centre-crop fallback); its ``RandAugment`` (an op list of its own) is pluggable —
``rand_augment=None`` uses pytorchvideo's class when it is importable and torchvision's ``RandAugment``
(magnitude 5, the same op applied to every frame of the clip) otherwise.
"""
from __future__ import annotations

import math
import random
from collections.abc import Callable
from dataclasses import dataclass
from typing import Any

import torch

from .utils import clean_narration_text, generate_input_ids_and_labels_from_interleaved

# the recipe's instruction prompts (train_v2.py:29-42, "based on prompts from InstructBLIP"): data, not code
PROMPTS = [
    "What is the camera wearer doing?",
    "Question: What is the camera wearer doing?",
    "What is the camera wearer doing? An answer to the question is",
    "Q: What is the camera wearer doing? A:",
    "Given the video, answer the following question. What is the camera wearer doing?",
    "Based on the video, respond to this question: What is the camera wearer doing? Answer:",
    "Use the provided video to answer the question: What is the camera wearer doing?",
    'What is the answer to the following question? "What is the camera wearer doing?"',
    'The question "What is the camera wearer doing?" can be answered using the video. The answer is',
]


@dataclass
class Preprocessor:
    """train_v2.py:45-75, same fields and behaviour: every in-context item becomes ``random prompt + " " +
    cleaned narration`` with one video, the query item a bare random prompt, the last narration the target;
    ``video_transform`` runs on every clip and the results are stacked into ``pixel_values``."""

    tokenizer: Any
    num_query_tokens: int
    decoder_only_lm: bool
    video_transform: Callable[[torch.Tensor], torch.Tensor] | None = None

    def __call__(self, datapoint: dict[str, Any]) -> dict[str, torch.Tensor]:
        items = datapoint["items"]
        prompts = [(random.choice(PROMPTS) + " " + clean_narration_text(item["narration_text"]), 1)
                   for item in items[:-1]]
        prompts.append((random.choice(PROMPTS), 1))
        out = generate_input_ids_and_labels_from_interleaved(
            self.tokenizer, prompts, clean_narration_text(items[-1]["narration_text"]), self.num_query_tokens,
            self.decoder_only_lm)
        videos = [item["video"] for item in items]
        if self.video_transform is not None:
            videos = [self.video_transform(v) for v in videos]
        out["pixel_values"] = torch.stack(videos)
        return out


def uniform_temporal_subsample(clip: torch.Tensor, num_samples: int, temporal_dim: int = 1) -> torch.Tensor:
    """pytorchvideo ``uniform_temporal_subsample``: ``num_samples`` frame indices equispaced over [0, T-1]
    (``linspace`` then truncation), first and last frame always included."""
    t = clip.shape[temporal_dim]
    idx = torch.linspace(0, t - 1, num_samples).clamp(0, t - 1).long()
    return torch.index_select(clip, temporal_dim, idx.to(clip.device))


def resized_crop_params(height: int, width: int, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                        num_tries: int = 10) -> tuple[int, int, int, int]:
    """(top, left, h, w) of a RandomResizedCrop: area fraction uniform in ``scale``, aspect ratio log-uniform in
    ``ratio``, ``num_tries`` rejection rounds, then the centre crop with the ratio clamped into range —
    torchvision ``RandomResizedCrop.get_params`` draw for draw (same torch RNG calls), which pytorchvideo's
    ``RandomResizedCrop`` follows.  Integer bookkeeping on the host."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(num_tries):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= width and 0 < h <= height:
            top = torch.randint(0, height - h + 1, size=(1,)).item()
            left = torch.randint(0, width - w + 1, size=(1,)).item()
            return top, left, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w, h = width, int(round(width / min(ratio)))
    elif in_ratio > max(ratio):
        h, w = height, int(round(height * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def _default_rand_augment(magnitude: int):
    try:  # the reference's own class (per-clip op sampling) when its library is installed
        from pytorchvideo.transforms import RandAugment as PvRandAugment

        aug = PvRandAugment(magnitude=magnitude)
        return lambda clip_tchw: aug(clip_tchw)
    except ImportError:
        from torchvision.transforms import v2

        aug = v2.RandAugment(magnitude=magnitude)
        # one draw per clip: torchvision transforms treat the leading dims of a (T, C, H, W) uint8 tensor as a
        # batch and apply the sampled ops to every frame alike, as pytorchvideo's video RandAugment does
        return lambda clip_tchw: aug(clip_tchw)


class TrainVideoTransform:
    """The video transform of train_v2.py:143-167 split where the B200 wants it split.

    ``host_part(clip)``  uint8 (C, T, H, W) -> (uint8 (C, num_frames, H, W), params): temporal subsample,
    RandAugment on the uint8 frames, and the random draws (crop box, flip).  Runs in dataloader workers.

    ``device_part(clip_u8, params)``  -> (C, num_frames, size, size) float on the GPU: the fused
    /255 + Normalize + RandomResizedCrop (bicubic) + flip kernel.

    Calling the object does both (the clip is moved to ``device`` in between), which is the drop-in
    ``video_transform`` for ``Preprocessor`` when the dataloader runs in the training process."""

    def __init__(self, size: tuple[int, int], image_mean, image_std, num_frames: int, *, rescale: float = 1.0 / 255.0,
                 scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), flip_p: float = 0.5, magnitude: int = 5,
                 rand_augment: Callable[[torch.Tensor], torch.Tensor] | None | bool = None,
                 device: torch.device | str | None = None, dtype: torch.dtype = torch.float32) -> None:
        self.size = (int(size[0]), int(size[1]))
        self.mean, self.std = [float(m) for m in image_mean], [float(s) for s in image_std]
        self.num_frames, self.rescale = int(num_frames), float(rescale)
        self.scale, self.ratio, self.flip_p = scale, ratio, float(flip_p)
        if rand_augment is False:
            self.rand_augment = None
        elif rand_augment is None or rand_augment is True:
            self.rand_augment = _default_rand_augment(magnitude)
        else:
            self.rand_augment = rand_augment
        self.device, self.dtype = device, dtype

    def host_part(self, clip: torch.Tensor):
        assert clip.dtype == torch.uint8 and clip.dim() == 4, "decoded uint8 clip (C, T, H, W) expected"
        clip = uniform_temporal_subsample(clip, self.num_frames, temporal_dim=1)
        if self.rand_augment is not None:  # Permute((1,0,2,3)) -> RandAugment -> Permute back (train_v2.py:147-152)
            clip = self.rand_augment(clip.permute(1, 0, 2, 3)).permute(1, 0, 2, 3)
        box = resized_crop_params(clip.shape[2], clip.shape[3], self.scale, self.ratio)
        flip = bool(torch.rand(1).item() < self.flip_p)  # torchvision RandomHorizontalFlip's draw
        return clip.contiguous(), dict(box=box, flip=flip)

    def device_part(self, clip_u8: torch.Tensor, params: dict, out: torch.Tensor | None = None) -> torch.Tensor:
        from .. import ops

        return ops.crop_resize_normalize_u8(clip_u8, params["box"], self.size, self.rescale, self.mean, self.std,
                                            flip=params["flip"], out=out, dtype=self.dtype)

    def __call__(self, clip: torch.Tensor) -> torch.Tensor:
        clip_u8, params = self.host_part(clip)
        dev = self.device if self.device is not None else (clip.device if clip.is_cuda else torch.device("cuda"))
        return self.device_part(clip_u8.to(dev, non_blocking=True), params)

    @classmethod
    def from_processor(cls, processor, num_frames: int, **kw) -> "TrainVideoTransform":
        """The recipe's parameters from a ``Blip2Processor`` (train_v2.py:153-164)."""
        ip = processor.image_processor
        return cls((ip.size["height"], ip.size["width"]), ip.image_mean, ip.image_std, num_frames,
                   rescale=getattr(ip, "rescale_factor", 1.0 / 255.0), **kw)
