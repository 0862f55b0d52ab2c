"""Golden vectors for the host-side integer code (eilev/data/utils.py: clean_narration_text,
generate_input_ids_and_labels, generate_input_ids_and_labels_from_interleaved, the two
collators), produced by the REAL reference functions on seeded random inputs with the
deterministic HashTokenizer.  Run in the authoring container only:
    python tests/golden/make_golden_data_utils.py
Writes tests/golden/data_utils_cases.json."""
import json
import random
import sys
import types
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, "/root/reference")

from stub_tokenizer import HashTokenizer  # noqa: E402

WORDS = ["#C", "C", "picks", "up", "the", "knife", "#O", "man", "X", "cuts", "onion.", "#unsure", "opens", "a",
         "drawer", "#Unsure", "with", "left", "hand", "What", "is", "camera", "wearer", "doing?", "Question:", "Answer:"]


def sentence(rng, lo=1, hi=7):
    return " ".join(rng.choice(WORDS) for _ in range(rng.randint(lo, hi)))


def main():
    pv = types.ModuleType("pytorchvideo")
    pvd = types.ModuleType("pytorchvideo.data")
    pvc = types.ModuleType("pytorchvideo.data.clip_sampling")
    pvd.ClipSampler = type("ClipSampler", (), {"__init__": lambda self, *a, **k: None})
    pvc.ClipInfo = tuple
    pv.data = pvd
    pvd.clip_sampling = pvc
    sys.modules.update({"pytorchvideo": pv, "pytorchvideo.data": pvd, "pytorchvideo.data.clip_sampling": pvc})
    import eilev.data.utils as R

    rng = random.Random(20240611)
    cases = {"clean": [], "pair": [], "interleaved": [], "collate": []}
    for _ in range(40):
        text = sentence(rng, 2, 9) + rng.choice(["", " ", ".", " #unsure", "  "])
        cases["clean"].append({"in": text, "out": R.clean_narration_text(text)})
    for _ in range(24):
        kind = rng.choice(["opt", "t5"])
        tok = HashTokenizer(kind)
        prompt, text = sentence(rng), sentence(rng)
        out = R.generate_input_ids_and_labels(tok, prompt, text, kind == "opt")
        cases["pair"].append({"kind": kind, "prompt": prompt, "text": text,
                              "input_ids": out["input_ids"].tolist(), "labels": out["labels"].tolist()})
    for _ in range(48):
        kind = rng.choice(["opt", "t5"])
        tok = HashTokenizer(kind)
        prompts = [(sentence(rng, 0, 5), rng.randint(0, 3)) for _ in range(rng.randint(1, 4))]
        text = rng.choice([None, sentence(rng)])
        nq = rng.choice([2, 4, 32])
        out = R.generate_input_ids_and_labels_from_interleaved(tok, prompts, text, nq, kind == "opt")
        cases["interleaved"].append({"kind": kind, "prompts": prompts, "text": text, "nq": nq,
                                     "out": {k: v.tolist() for k, v in out.items()}})
    for _ in range(24):
        kind = rng.choice(["opt", "t5"])
        side = rng.choice(["left", "right"])
        multiple = rng.choice([None, 8])
        tok = HashTokenizer(kind, side)
        nq = 2
        feats, spec = [], []
        for _ in range(rng.randint(1, 4)):
            prompts = [(sentence(rng, 0, 4), rng.randint(0, 2)) for _ in range(rng.randint(1, 3))]
            text = sentence(rng)
            f = R.generate_input_ids_and_labels_from_interleaved(tok, prompts, text, nq, kind == "opt")
            nv = sum(n for _, n in prompts)
            f["pixel_values"] = torch.arange(nv * 2, dtype=torch.float32).view(nv, 2) + 100 * len(feats)
            feats.append(f)
            spec.append({"prompts": prompts, "text": text})
        out = R.DataCollatorForInterleavedVideoSeq2Seq(tok, pad_to_multiple_of=multiple)([dict(f) for f in feats])
        cases["collate"].append({"kind": kind, "side": side, "multiple": multiple, "nq": nq, "samples": spec,
                                 "out": {k: v.tolist() for k, v in out.items()}})
    (HERE / "data_utils_cases.json").write_text(json.dumps(cases))
    print({k: len(v) for k, v in cases.items()}, (HERE / "data_utils_cases.json").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
