"""Deterministic word-hash tokenizer shared by the golden generator (make_golden_data_utils.py,
which feeds it to the REAL reference functions) and the replay test.  No hub access needed."""
import zlib

import torch
from transformers import BatchEncoding


class HashTokenizer:
    def __init__(self, kind: str, padding_side: str = "right") -> None:
        assert kind in ("opt", "t5")
        self.kind, self.padding_side = kind, padding_side
        if kind == "opt":
            self.bos_token_id, self.eos_token_id, self.pad_token_id, self.newline = 2, 2, 1, 50118
        else:
            self.bos_token_id, self.eos_token_id, self.pad_token_id, self.newline = None, 1, 0, 3
        self.model_input_names = ["input_ids", "attention_mask"]

    def _ids(self, text: str) -> list[int]:
        out = []
        for piece in text.replace("\n", " \n ").split(" "):
            if piece == "\n":
                out.append(self.newline)
            elif piece:
                out.append(4 + zlib.crc32(piece.encode()) % 30000)
        return out

    def __call__(self, text, add_special_tokens=True, return_attention_mask=True, **_):
        ids = self._ids(text)
        if add_special_tokens:
            if self.kind == "opt":
                ids = [self.bos_token_id] + ids
            else:
                ids = ids + [self.eos_token_id]
        return BatchEncoding({"input_ids": ids})

    def pad(self, features, padding=True, max_length=None, pad_to_multiple_of=None, return_tensors=None, **_):
        width = max(len(f["input_ids"]) for f in features)
        if pad_to_multiple_of:
            width = (width + pad_to_multiple_of - 1) // pad_to_multiple_of * pad_to_multiple_of
        out = {"input_ids": [], "attention_mask": []}
        extra = [k for k in features[0] if k not in ("input_ids", "attention_mask")]
        for k in extra:
            out[k] = []
        for f in features:
            ids = list(map(int, f["input_ids"]))
            n = width - len(ids)
            if self.padding_side == "right":
                out["input_ids"].append(ids + [self.pad_token_id] * n)
                out["attention_mask"].append([1] * len(ids) + [0] * n)
            else:
                out["input_ids"].append([self.pad_token_id] * n + ids)
                out["attention_mask"].append([0] * n + [1] * len(ids))
            for k in extra:
                out[k].append(list(map(int, f[k])))
        return BatchEncoding({k: torch.tensor(v) for k, v in out.items()})
