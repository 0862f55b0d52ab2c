"""Reader of the reference's extracted-frame directory format (eilev/data/frame.py:14-72
``FrameDataset``; written by scripts/ego4d/extract_frames.py:66-105):

    frames_dir/narrated_actions.csv          columns frame_path, video_uid, clip_index,
                                             narration_timestamp_sec, narration_text,
                                             structured_verb, structured_noun
    frames_dir/<frame_path>/*.png            the clip's frames (8 x 448 x 448 RGB), in name order

Same constructor and item contract as the reference — ``{**csv_row, "video": uint8 (C, T, H, W)}``,
integer or ``frame_path`` indexing, optional ``data_filter`` / ``transform`` — without pytorchvideo:
the PNGs are decoded with Pillow and stay uint8, which is what ``process_on_device`` /
``process(normalize_on_device=True)`` consume.  The in-context example samplers of the reference
(``FrameInterleavedDataset`` …) are dataset policy, not data format, and are not rebuilt.
"""
from __future__ import annotations

import re
from collections.abc import Callable
from csv import DictReader
from pathlib import Path
from typing import Any

import numpy as np
import torch
from torch.utils.data import Dataset

_IMAGE_SUFFIXES = {".png", ".jpg", ".jpeg"}


def _natural_key(name: str):
    """'…|2.png' before '…|10.png' (pytorchvideo's FrameVideo orders frames numerically)."""
    return [int(tok) if tok.isdigit() else tok for tok in re.split(r"(\d+)", name)]


def read_frame_dir(path: str | Path) -> torch.Tensor:
    """All images of a directory, in natural name order, as one uint8 tensor (C, T, H, W)."""
    from PIL import Image

    files = sorted((p for p in Path(path).iterdir() if p.suffix.lower() in _IMAGE_SUFFIXES),
                   key=lambda p: _natural_key(p.name))
    if not files:
        raise FileNotFoundError(f"no frames under {path}")
    frames = []
    for f in files:
        with Image.open(f) as im:
            frames.append(np.asarray(im.convert("RGB")))
    shapes = {fr.shape for fr in frames}
    if len(shapes) != 1:
        raise ValueError(f"frames under {path} differ in size: {sorted(shapes)}")
    return torch.from_numpy(np.stack(frames)).permute(3, 0, 1, 2).contiguous()  # (T, H, W, C) -> (C, T, H, W)


class FrameDataset(Dataset):
    def __init__(self, frames_dir: str, annotation_file: str | None = None,
                 transform: Callable[[dict[str, Any]], Any] | None = None,
                 data_filter: Callable[[dict[str, Any]], bool] | None = None,
                 return_frames: bool = True) -> None:
        self.frames_dir = Path(frames_dir)
        self.return_frames = return_frames
        self.annotation_file_path = (self.frames_dir / "narrated_actions.csv" if annotation_file is None
                                     else Path(annotation_file))
        assert self.annotation_file_path.exists()  # frame.py:41
        self.data: list[dict] = []
        self.dict_data: dict[str, dict] = {}
        with open(self.annotation_file_path, newline="") as csvfile:
            for row in DictReader(csvfile):
                if data_filter is not None and not data_filter(row):
                    continue
                self.data.append(row)
                self.dict_data[row["frame_path"]] = row
        self._transform = transform

    def __getitem__(self, index: int | str) -> dict[str, Any]:
        row = self.data[index] if isinstance(index, int) else self.dict_data[index]
        item = dict(row)
        if self.return_frames:
            item["video"] = read_frame_dir(self.frames_dir / row["frame_path"])
        return self._transform(item) if self._transform is not None else item

    def __len__(self) -> int:
        return len(self.data)
