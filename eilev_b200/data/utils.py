"""Interleaved video/text tokenisation and collation — the integer input contract of the
VideoBLIP path (eilev/data/utils.py:35-66 collator, :95-140 and :143-223 tokenisers,
:69-92 narration clean-up).  Pure host-side integer work, bit-exact with the reference
(the reference's tests/data/test_utils.py tables are replayed in tests/test_host_cpu.py,
seeded outputs of the real reference functions in tests/test_data_utils_golden.py).  The raw-video clip sampler of the reference is out of scope.
"""
from __future__ import annotations

import re
import string

import torch
from transformers import DataCollatorForSeq2Seq

_RE_CAMERA_WEARER = re.compile(r"^\#C\s+C", re.IGNORECASE)
_RE_EOS = re.compile(r"\<\|eos\|\>$", re.IGNORECASE)
_RE_UNSURE_TAIL = re.compile(r"#unsure\.?$", re.IGNORECASE)
_RE_UNSURE = re.compile(r"#unsure", re.IGNORECASE)


def clean_narration_text(narration_text: str) -> str:
    """Ego4D narration normalisation (eilev/data/utils.py:69-92)."""
    s = narration_text.strip()
    s = _RE_CAMERA_WEARER.sub("The camera wearer", s).strip()
    s = _RE_EOS.sub("", s).strip()
    s = _RE_UNSURE_TAIL.sub("", s).strip()
    s = _RE_UNSURE.sub("something", s)
    if s and s[-1] not in string.punctuation:
        s += "."
    return s


def generate_input_ids_and_labels(tokenizer, prompt: str, text: str, decoder_only_lm: bool):
    """Single-clip tokenisation (eilev/data/utils.py:95-140)."""
    if decoder_only_lm:
        head = tokenizer(prompt, return_attention_mask=False).input_ids
        enc = tokenizer(" " + text, return_attention_mask=False, add_special_tokens=False)
        ids = torch.tensor(list(head) + list(enc["input_ids"]) + [tokenizer.eos_token_id])
        labels = ids.clone()
        labels[: len(head)] = -100  # the prompt is not a target
        enc["input_ids"], enc["labels"] = ids, labels
        return enc
    enc = tokenizer(prompt, return_attention_mask=False)  # the tokenizer appends eos itself
    enc["input_ids"] = torch.tensor(enc["input_ids"])
    enc["labels"] = torch.tensor(tokenizer(text, return_attention_mask=False).input_ids)
    return enc


def generate_input_ids_and_labels_from_interleaved(
    tokenizer, prompts: list[tuple[str, int]], text: str | None, num_query_tokens: int,
    decoder_only_lm: bool,
) -> dict[str, torch.Tensor]:
    """Interleaved tokenisation (eilev/data/utils.py:143-223).

    Every clip contributes ``num_query_tokens`` pad-id placeholders (mask 1) followed by a
    newline token; decoder-only LMs get a leading bos and the target text (" " + text + "\\n"
    + eos) appended as labels; seq2seq LMs get eos after the last prompt and labels =
    tokenizer(text).
    """
    nl = tokenizer("\n", add_special_tokens=False).input_ids[0]
    clip_ids = [tokenizer.pad_token_id] * num_query_tokens + [nl]
    clip_mask = [1] * num_query_tokens + [0]
    ids: list[int] = []
    mask: list[int] = []
    labels: list[int] = []
    last = len(prompts) - 1
    for i, (prompt, num_videos) in enumerate(prompts):
        if decoder_only_lm and i == 0:
            ids.append(tokenizer.bos_token_id)
            mask.append(0)
        ids += clip_ids * num_videos
        mask += clip_mask * num_videos
        piece = tokenizer(prompt if i == last else prompt + "\n", add_special_tokens=False).input_ids
        piece = list(piece)
        if not decoder_only_lm and i == last:
            piece.append(tokenizer.eos_token_id)
        ids += piece
        mask += [0] * len(piece)
    if decoder_only_lm:
        labels = [-100] * len(ids)
        if text is not None:
            tgt = list(tokenizer(" " + text + "\n", add_special_tokens=False).input_ids)
            tgt.append(tokenizer.eos_token_id)
            ids += tgt
            mask += [0] * len(tgt)
            labels += tgt
    elif text is not None:
        labels = list(tokenizer(text).input_ids)
    return {"input_ids": torch.tensor(ids), "labels": torch.tensor(labels),
            "video_input_mask": torch.tensor(mask)}


class DataCollatorForVideoSeq2Seq(DataCollatorForSeq2Seq):
    """One clip per sample: stacks ``pixel_values`` (eilev/data/utils.py:19-32)."""

    def __call__(self, features, return_tensors=None):
        pixel_values = None
        if all("pixel_values" in f for f in features):
            pixel_values = torch.stack([f.pop("pixel_values") for f in features])
        batch = super().__call__(features, return_tensors=return_tensors)
        if pixel_values is not None:
            batch["pixel_values"] = pixel_values
        return batch


class DataCollatorForInterleavedVideoSeq2Seq(DataCollatorForSeq2Seq):
    """Interleaved samples (eilev/data/utils.py:35-66): clips of all samples are concatenated
    along the clip axis in batch order — the order the model's splice consumes them in —
    ids/labels are padded by ``DataCollatorForSeq2Seq`` and ``video_input_mask`` is padded with
    zeros on the tokenizer's padding side."""

    def __call__(self, features, return_tensors=None):
        if "pixel_values" not in features[0].keys():
            raise TypeError("cat() received an invalid combination of arguments - got (NoneType)")
        pixel_values = torch.cat([f.pop("pixel_values") for f in features])
        masks = None
        if "video_input_mask" in features[0].keys():
            masks = [f.pop("video_input_mask") for f in features]
        batch = super().__call__(features, return_tensors=return_tensors)
        if masks is not None:
            width = batch["input_ids"].size(1)
            left = self.tokenizer.padding_side != "right"
            rows = []
            for m in masks:
                pad = torch.zeros(width - len(m), dtype=torch.long)
                rows.append(torch.cat([pad, m] if left else [m, pad]))
            batch["video_input_mask"] = torch.stack(rows)
        batch["pixel_values"] = pixel_values
        return batch


def generate_chunks(list_to_chunk: list, chunk_size: int):
    """Consecutive slices of ``chunk_size`` items, the last one shorter (eilev/data/utils.py:229-231;
    used by the evaluation scripts to batch prompts)."""
    start = 0
    while start < len(list_to_chunk):
        yield list_to_chunk[start:start + chunk_size]
        start += chunk_size


def parse_timestamp(timestamp: str) -> float:
    """``hh:mm:ss.cc`` -> seconds (eilev/data/utils.py:234-241).  Same float expression order as the
    reference, so the results are bit-identical."""
    hours, minutes, seconds = timestamp.split(":")
    return float(hours) * 60 * 60 + float(minutes) * 60 + float(seconds)
