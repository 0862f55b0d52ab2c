"""generate() with more than 16 rows (batch x beams) on the decoder-only LM: the rows are dealt to
groups of <= 16 whole beam sets (model/generation.py::_GroupedStepper), every group a paged-KV
decode state of its own on the same kernels.  The grouping logic itself is tested on CPU against
a single stepper (tests/test_host_cpu.py); here the real engine runs it.  Same acceptance rule as
tests/test_model_gpu.py::test_generate_greedy_matches_reference_golden (bf16 rounding may flip
near-ties of a random-init model): the first token of every row must equal the reference's, and
overall agreement >= 60 %.

Written after round 1's GPU budget was spent: this file first runs in the round-end
``pytest -m gpu`` (it sorts last, so nothing else depends on it)."""
import json
import os
from pathlib import Path

import pytest
import torch
from transformers import Blip2Config

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"


def _load(name):
    fx = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(fx["state_dict"])
    return fx, m.to("cuda").eval()


@pytest.mark.parametrize("copies", [10, 9])  # 20 rows -> groups of 16 + 4; 18 rows -> 16 + 2
def test_greedy_generate_with_more_than_16_rows(copies):
    fx, m = _load("small_opt")
    g = {k: v.cuda() for k, v in fx["gen_inputs"].items()}
    n_new = fx["generated"].shape[1]
    big = dict(input_ids=g["input_ids"].repeat(copies, 1), attention_mask=g["attention_mask"].repeat(copies, 1),
               video_input_mask=g["video_input_mask"].repeat(copies, 1),
               pixel_values=g["pixel_values"].repeat(copies, 1, 1, 1, 1))
    toks = m.generate(**big, max_new_tokens=n_new, min_new_tokens=n_new, do_sample=False, num_beams=1).cpu()
    m.check_splice()
    want = fx["generated"].repeat(copies, 1)
    assert toks.shape == want.shape
    agree = float((toks == want).float().mean())
    out = Path(os.environ.get("GRAFT_REPO_ROOT", ".")) / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"parity_report_decode_rows_{copies}.json").write_text(json.dumps(dict(rows=toks.shape[0], agree=agree)))
    assert torch.equal(toks[:, 0], want[:, 0])
    assert agree >= 0.6, agree
    # rows are independent: every copy of a prompt decodes like the first one inside its own group
    small = m.generate(**g, max_new_tokens=n_new, min_new_tokens=n_new, do_sample=False, num_beams=1).cpu()
    assert torch.equal(toks[:2, 0], small[:, 0])


def test_beam_search_with_more_than_16_rows():
    fx, m = _load("small_opt")
    g = {k: v.cuda() for k, v in fx["gen_inputs"].items()}
    big = dict(input_ids=g["input_ids"].repeat(3, 1), attention_mask=g["attention_mask"].repeat(3, 1),
               video_input_mask=g["video_input_mask"].repeat(3, 1), pixel_values=g["pixel_values"].repeat(3, 1, 1, 1, 1))
    beams = m.generate(**big, max_new_tokens=5, num_beams=4)  # 6 prompts x 4 beams = 24 rows -> 16 + 8
    assert beams.shape[0] == 6 and 1 <= beams.shape[1] <= 5
    assert int(beams.min()) >= 0 and int(beams.max()) < m.config.text_config.vocab_size
    # the three copies of each prompt live in different groups (16 + 8 rows) and must agree with each other
    # wherever the search is not sitting on a numerical tie: demand it for the first token of 2 of 3 copies
    first = beams[:, 0].cpu().view(3, 2)
    for col in range(2):
        vals = first[:, col].tolist()
        assert max(vals.count(v) for v in vals) >= 2, vals
