"""Golden beam-search outputs of the REAL reference (eilev/model/v2.py generate -> HF beam search)
for the OPT fixtures.  Run in the authoring container only:
    python tests/golden/make_golden_beams.py
Writes tests/golden/beams_<name>.pt = {case name: generated ids}."""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from transformers import Blip2Config  # noqa: E402

CASES = {
    "beams3": dict(num_beams=3, max_new_tokens=6, do_sample=False),
    "beams4_lp2": dict(num_beams=4, max_new_tokens=5, do_sample=False, length_penalty=2.0),
    "beams2_min": dict(num_beams=2, max_new_tokens=6, min_new_tokens=6, do_sample=False),
    "greedy_rep": dict(num_beams=1, max_new_tokens=6, do_sample=False, repetition_penalty=1.3),
}


def main():
    sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
    from eilev.model.v2 import VideoBlipForConditionalGeneration as RefModel

    here = Path(__file__).resolve().parent
    for name in ("tiny_opt", "small_opt"):
        fx = torch.load(here / f"{name}.pt", weights_only=False)
        cfg = Blip2Config(**{k: fx["config"][k] for k in
                             ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
        model = RefModel(cfg).float().eval()
        model.load_state_dict(fx["state_dict"])
        model.tie_weights()
        out = {}
        with torch.no_grad():
            for case, kw in CASES.items():
                out[case] = model.generate(**fx["gen_inputs"], **kw)
                print(name, case, out[case].tolist())
        torch.save(out, here / f"beams_{name}.pt")


if __name__ == "__main__":
    main()
