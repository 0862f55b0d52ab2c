"""Host-side logic that needs no GPU: HF-surface construction / state-dict contract,
save/load round trip, error behaviour, tokeniser + collator golden vectors from the
reference's tests/data/test_utils.py (replayed with a table-driven stub tokenizer)."""
import sys
from pathlib import Path

import pytest
import torch
from transformers import Blip2Config

GOLDEN = Path(__file__).resolve().parent / "golden"


def small_cfg():
    sys.path.insert(0, str(GOLDEN))
    import make_golden as MG
    return Blip2Config(**MG.CONFIGS["small_opt"]["config"])


def test_state_dict_keys_match_the_reference_checkpoint_layout():
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    fx = torch.load(GOLDEN / "small_opt.pt", weights_only=False)
    m = VideoBlipForConditionalGeneration(small_cfg())
    assert set(m.state_dict()) == set(fx["state_dict"])
    for k, v in m.state_dict().items():
        assert v.shape == fx["state_dict"][k].shape, k
    m.load_state_dict(fx["state_dict"], strict=True)
    lm = m.language_model
    assert lm.lm_head.weight.data_ptr() == lm.model.decoder.embed_tokens.weight.data_ptr()
    assert m.get_input_embeddings() is lm.model.decoder.embed_tokens
    assert float(VideoBlipForConditionalGeneration(small_cfg()).query_tokens.abs().sum()) == 0.0  # v2.py:115-117


def test_save_and_from_pretrained_round_trip(tmp_path):
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    fx = torch.load(GOLDEN / "tiny_opt.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    m = VideoBlipForConditionalGeneration(cfg)
    m.load_state_dict(fx["state_dict"])
    m.save_pretrained(tmp_path)
    m2 = VideoBlipForConditionalGeneration.from_pretrained(tmp_path, low_cpu_mem_usage=True)
    for k, v in m2.state_dict().items():
        assert torch.equal(v, fx["state_dict"][k]), k
    m3 = VideoBlipForConditionalGeneration.from_pretrained(tmp_path, torch_dtype=torch.bfloat16)
    assert m3.dtype == torch.bfloat16
    assert m3.language_model.lm_head.weight.data_ptr() == m3.language_model.model.decoder.embed_tokens.weight.data_ptr()
    assert m3.config.num_query_tokens == cfg.num_query_tokens
    m3.config.text_config.eos_token_id = 7  # train_v2.py:122 writes this
    assert m3.config.text_config.eos_token_id == 7


def test_recipe_freezing_and_param_census():
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import freeze_for_recipe
    m = VideoBlipForConditionalGeneration(small_cfg())
    freeze_for_recipe(m)
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert all(n.startswith(("qformer.", "query_tokens", "language_projection.")) for n in trainable)
    fx = torch.load(GOLDEN / "small_opt.pt", weights_only=False)
    assert set(trainable) == set(fx["grads"])


def test_forward_contract_errors_on_cpu():
    from eilev_b200 import _lib
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration, VideoBlipVisionModel
    cfg = small_cfg()
    m = VideoBlipForConditionalGeneration(cfg)
    with pytest.raises(AssertionError):  # pixel_values without video_input_mask (v2.py:154-157)
        m(torch.ones(1, 4).long(), pixel_values=torch.zeros(1, 3, 1, 56, 56))
    with pytest.raises(_lib.VbError):  # no silent CPU fallback
        m(torch.ones(1, 4).long())
    with pytest.raises(ValueError):  # v2.py:50-51
        VideoBlipVisionModel(cfg.vision_config)(None)
    with pytest.raises(NotImplementedError):  # T5 feed-forward variants other than relu / gated-gelu
        VideoBlipForConditionalGeneration(Blip2Config(
            vision_config=cfg.vision_config.to_dict(), qformer_config=cfg.qformer_config.to_dict(),
            text_config={"model_type": "t5", "d_model": 8, "d_kv": 4, "d_ff": 16, "num_layers": 2, "num_heads": 2,
                         "feed_forward_proj": "gated-silu"}))


def test_t5_holder_variants_follow_the_reference_checkpoint_layout():
    """The reference's own T5 test config (tests/model/test_model_v2.py:122-146: T5Config defaults =
    ReLU T5DenseActDense, head tied to the embedding) and the flan-style gated-gelu config: parameter
    names / shapes of the real reference's state_dict (golden fixtures), head tying."""
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    for name, tied in (("tiny_t5_relu", True), ("small_t5", False)):
        fx = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
        cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
        m = VideoBlipForConditionalGeneration(cfg)
        assert set(m.state_dict()) == set(fx["state_dict"]), name
        m.load_state_dict(fx["state_dict"], strict=True)
        lm = m.language_model
        assert (lm.lm_head.weight.data_ptr() == lm.shared.weight.data_ptr()) == tied, name
        assert lm.encoder.embed_tokens.weight.data_ptr() == lm.shared.weight.data_ptr()
        assert torch.equal(lm.lm_head.weight, fx["state_dict"]["language_model.lm_head.weight"])
        ff = lm.encoder.block[0].layer[1].DenseReluDense
        assert hasattr(ff, "wi") != hasattr(ff, "wi_0")


# ------------------------------------------------------------------ tokeniser / collator
class StubTokenizer:
    """Table-driven tokenizer with the ids of the reference fixtures
    (tests/data/test_utils.py:112-127, :463-474): OPT bos=eos=2 pad=1 '\\n'=50118."""

    def __init__(self, table, bos, eos, pad, add_bos, add_eos, padding_side="right"):
        self.table, self.bos_token_id, self.eos_token_id, self.pad_token_id = table, bos, eos, pad
        self.add_bos, self.add_eos, self.padding_side = add_bos, add_eos, padding_side
        self.model_input_names = ["input_ids", "attention_mask"]

    def __call__(self, text, add_special_tokens=True, return_attention_mask=True, **_):
        from transformers import BatchEncoding
        ids = list(self.table[text])
        if add_special_tokens:
            if self.add_bos:
                ids = [self.bos_token_id] + ids
            if self.add_eos:
                ids = ids + [self.eos_token_id]
        return BatchEncoding({"input_ids": ids})

    def pad(self, features, padding=True, max_length=None, pad_to_multiple_of=None, return_tensors=None, **_):
        from transformers import BatchEncoding
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


OPT_TABLE = {"\n": [50118], "A prompt": [250, 14302], " A text\n": [83, 2788, 50118],
             "Prompt 1 Text 1\n": [35396, 3320, 112, 14159, 112, 50118], "Prompt 2": [35396, 3320, 132],
             " Text 2\n": [14159, 132, 50118]}
T5_TABLE = {"\n": [3], "A prompt": [71, 9005], "A text": [71, 1499]}


def opt_tok(side="right"):
    return StubTokenizer(OPT_TABLE, 2, 2, 1, add_bos=True, add_eos=False, padding_side=side)


def t5_tok():
    return StubTokenizer(T5_TABLE, None, 1, 0, add_bos=False, add_eos=True)


def test_interleaved_tokeniser_opt_golden():
    """tests/data/test_utils.py:115-127 of the reference."""
    from eilev_b200.data.utils import generate_input_ids_and_labels_from_interleaved as gen
    r = gen(opt_tok(), [("A prompt", 1)], "A text", 2, True)
    assert r["input_ids"].tolist() == [2, 1, 1, 50118, 250, 14302, 83, 2788, 50118, 2]
    assert r["labels"].tolist() == [-100] * 6 + [83, 2788, 50118, 2]
    assert r["video_input_mask"].tolist() == [0, 1, 1, 0, 0, 0, 0, 0, 0, 0]
    r = gen(opt_tok(), [("Prompt 1 Text 1", 1), ("Prompt 2", 2)], "Text 2", 2, True)
    assert r["input_ids"].tolist() == [2, 1, 1, 50118, 35396, 3320, 112, 14159, 112, 50118,
                                       1, 1, 50118, 1, 1, 50118, 35396, 3320, 132, 14159, 132, 50118, 2]
    assert r["video_input_mask"].tolist() == [0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0]
    assert r["labels"].tolist() == [-100] * 19 + [14159, 132, 50118, 2]
    r = gen(opt_tok(), [("A prompt", 1)], None, 2, True)  # generation: no target
    assert r["input_ids"].tolist() == [2, 1, 1, 50118, 250, 14302]
    assert r["labels"].tolist() == [-100] * 6


def test_interleaved_tokeniser_t5_golden():
    """tests/data/test_utils.py:466-474 of the reference."""
    from eilev_b200.data.utils import generate_input_ids_and_labels_from_interleaved as gen
    r = gen(t5_tok(), [("A prompt", 1)], "A text", 2, False)
    assert r["input_ids"].tolist() == [0, 0, 3, 71, 9005, 1]
    assert r["labels"].tolist() == [71, 1499, 1]
    assert r["video_input_mask"].tolist() == [1, 1, 0, 0, 0, 0]


@pytest.mark.parametrize("side", ["right", "left"])
@pytest.mark.parametrize("multiple", [None, 8])
def test_interleaved_collator(side, multiple):
    """Padding side x pad_to_multiple_of matrix of tests/data/test_utils.py:674-862."""
    from eilev_b200.data.utils import DataCollatorForInterleavedVideoSeq2Seq
    tok = opt_tok(side)
    col = DataCollatorForInterleavedVideoSeq2Seq(tok, pad_to_multiple_of=multiple)
    feats = [
        {"input_ids": torch.tensor([2, 1, 1, 50118, 250]), "labels": torch.tensor([-100, -100, -100, -100, 250]),
         "video_input_mask": torch.tensor([0, 1, 1, 0, 0]), "pixel_values": torch.zeros(1, 3, 2, 4, 4)},
        {"input_ids": torch.tensor([2, 1, 1, 50118, 1, 1, 50118, 250, 14302]),
         "labels": torch.tensor([-100] * 7 + [250, 14302]),
         "video_input_mask": torch.tensor([0, 1, 1, 0, 1, 1, 0, 0, 0]), "pixel_values": torch.ones(2, 3, 2, 4, 4)},
    ]
    out = col(feats)
    width = 16 if multiple else 9
    assert out["input_ids"].shape == (2, width) and out["video_input_mask"].shape == (2, width)
    assert out["pixel_values"].shape == (3, 3, 2, 4, 4)
    assert float(out["pixel_values"][0].sum()) == 0 and float(out["pixel_values"][1:].min()) == 1  # batch order
    n0 = width - 5
    if side == "right":
        assert out["video_input_mask"][0].tolist() == [0, 1, 1, 0, 0] + [0] * n0
        assert out["input_ids"][0].tolist() == [2, 1, 1, 50118, 250] + [1] * n0
        assert out["labels"][0].tolist() == [-100, -100, -100, -100, 250] + [-100] * n0
        assert out["attention_mask"][0].tolist() == [1] * 5 + [0] * n0
    else:
        assert out["video_input_mask"][0].tolist() == [0] * n0 + [0, 1, 1, 0, 0]
        assert out["input_ids"][0].tolist() == [1] * n0 + [2, 1, 1, 50118, 250]
        assert out["labels"][0].tolist() == [-100] * n0 + [-100, -100, -100, -100, 250]
    assert int(out["video_input_mask"].sum()) == 3 * 2


def test_clean_narration_text_table():
    """tests/data/test_utils.py:19-54 of the reference."""
    from eilev_b200.data.utils import clean_narration_text as c
    assert c("#C C drops a plate") == "The camera wearer drops a plate."
    assert c("#c c looks around <|eos|>") == "The camera wearer looks around."
    assert c("#C C picks #unsure") == "The camera wearer picks."
    assert c("#C C holds #Unsure in hand") == "The camera wearer holds something in hand."
    assert c("  already punctuated!  ") == "already punctuated!"
    assert c("") == ""


def test_single_clip_tokeniser():
    from eilev_b200.data.utils import generate_input_ids_and_labels as gen
    table = dict(OPT_TABLE)
    table[" A text"] = [83, 2788]
    tok = StubTokenizer(table, 2, 2, 1, add_bos=True, add_eos=False)
    r = gen(tok, "A prompt", "A text", True)
    assert r["input_ids"].tolist() == [2, 250, 14302, 83, 2788, 2]
    assert r["labels"].tolist() == [-100, -100, -100, 83, 2788, 2]
    r = gen(t5_tok(), "A prompt", "A text", False)
    assert r["input_ids"].tolist() == [71, 9005, 1] and r["labels"].tolist() == [71, 1499, 1]


def test_process_unfolds_time_axis():
    """tests/model/test_model_utils.py of the reference (Mock processor)."""
    from unittest.mock import Mock
    from transformers import BatchEncoding
    from eilev_b200.model.utils import process
    proc = Mock(return_value=BatchEncoding({"pixel_values": torch.zeros(2 * 5, 3, 224, 224)}))
    out = process(proc, video=torch.zeros(2, 3, 5, 32, 48, dtype=torch.uint8), text="hi")
    assert out["pixel_values"].shape == (2, 3, 5, 224, 224)
    assert proc.call_args.kwargs["images"].shape == (10, 3, 32, 48)
    proc = Mock(return_value=BatchEncoding({"pixel_values": torch.zeros(5, 3, 224, 224)}))
    assert process(proc, video=torch.zeros(3, 5, 32, 48))["pixel_values"].shape == (1, 3, 5, 224, 224)


# ------------------------------------------------------------------ flan-T5 host logic
def _t5_fixture():
    fx = torch.load(GOLDEN / "small_t5.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    return fx, cfg


def test_t5_state_dict_keys_match_the_reference_checkpoint_layout(tmp_path):
    """Blip2Config with a T5 text_config builds the seq2seq model with HF T5's key names
    (shared / encoder.block.N.layer.* / decoder.block.N.layer.* / lm_head), embeddings aliased."""
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    fx, cfg = _t5_fixture()
    m = VideoBlipForConditionalGeneration(cfg)
    assert set(m.state_dict()) == set(fx["state_dict"])
    for k, v in m.state_dict().items():
        assert v.shape == fx["state_dict"][k].shape, k
    m.load_state_dict(fx["state_dict"], strict=True)
    lm = m.language_model
    assert lm.encoder.embed_tokens.weight.data_ptr() == lm.shared.weight.data_ptr()
    assert lm.decoder.embed_tokens.weight.data_ptr() == lm.shared.weight.data_ptr()
    assert m.get_input_embeddings() is lm.shared
    assert not m.config.use_decoder_only_language_model
    m.save_pretrained(tmp_path)
    m2 = VideoBlipForConditionalGeneration.from_pretrained(tmp_path)
    for k, v in m2.state_dict().items():
        assert torch.equal(v, fx["state_dict"][k]), k
    with pytest.raises(AssertionError):  # classify is decoder-only in the reference (v2.py:351)
        m.classify(torch.zeros(1, 4, dtype=torch.long), torch.zeros(2, 3, dtype=torch.long))


@pytest.mark.parametrize("sq,skv,bidir", [(104, 104, True), (9, 9, False), (1, 1, False), (300, 300, True), (40, 40, False)])
def test_t5_relative_bias_table_matches_the_oracle(sq, skv, bidir):
    """engine/t5.py::rel_bias_table (heads, sq+skv-1) vs T5Attention.compute_bias restated by the
    oracle: entry (j - i) + (sq - 1) is the bias of key j seen from query i — bit-exact."""
    from eilev_b200.engine import t5 as E
    from oracle import videoblip_ref as R
    _, cfg = _t5_fixture()
    tc = cfg.text_config
    w = torch.randn(tc.relative_attention_num_buckets, tc.num_heads)
    tab = E.rel_bias_table(w, sq, skv, bidir, tc)
    ref = R.t5_position_bias(w, sq, skv, bidir, tc)[0]  # (heads, sq, skv)
    i = torch.arange(sq)[:, None]
    j = torch.arange(skv)[None, :]
    assert tab.shape == (tc.num_heads, sq + skv - 1)
    assert torch.equal(tab[:, (j - i) + (sq - 1)], ref)


def test_t5_shift_right_matches_the_oracle():
    from eilev_b200.engine import t5 as E
    from oracle import videoblip_ref as R
    _, cfg = _t5_fixture()
    labels = torch.tensor([[5, 9, 7, -100, -100], [3, 4, 8, 2, 1]])
    assert torch.equal(E.shift_right(labels, cfg.text_config), R.t5_shift_right(labels, cfg.text_config))


def test_decode_op_record_layout_matches_the_header():
    """numpy mirror of vb_decode_op (ops.op_dtype) vs the struct in include/videoblip_b200.h."""
    import re
    from eilev_b200 import ops
    header = (Path(__file__).resolve().parent.parent / "include" / "videoblip_b200.h").read_text()
    body = re.search(r"typedef struct vb_decode_op \{(.*?)\} vb_decode_op;", header, re.S).group(1)
    fields = re.findall(r"(\w+)\s*(?:\[(\d+)\])?;", body)
    dt = ops.op_dtype()
    assert [f[0] for f in fields] == list(dt.names) == ["type", "i32", "ptr", "i64", "f32"]
    assert dt.itemsize == 192 and dt.fields["ptr"][1] == 32 and dt.fields["i64"][1] == 112 and dt.fields["f32"][1] == 176
    assert int(re.search(r"#define VB_DECODE_STEP_WS_BYTES \((\d+) \+ (\d+) \* (\d+)\)", header).group(1)) == 4096
    assert ops.DECODE_STEP_WS_BYTES == 4096 + 1008 * 512


# ------------------------------------------------------------------ generation host logic
def test_logit_processors_match_huggingface():
    """The token-bookkeeping half of generate() (repetition penalty, min-new-tokens EOS
    suppression, temperature / top-k / top-p warping) against the HuggingFace processors the
    reference reaches through language_model.generate (samples/*.py pass these kwargs)."""
    from transformers.generation.logits_process import (
        MinNewTokensLengthLogitsProcessor, RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper,
        TopKLogitsWarper, TopPLogitsWarper)
    from eilev_b200.model import generation as G
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(3, 50, generator=g) * 3
    generated = torch.randint(0, 50, (3, 6), generator=g)
    # repetition penalty
    want = RepetitionPenaltyLogitsProcessor(1.3)(generated, logits.clone())
    got = G._process_logits(logits.clone(), generated, step=6, min_new_tokens=0, eos_ids=[2], repetition_penalty=1.3)
    assert torch.allclose(got, want)
    # min_new_tokens: EOS suppressed while fewer than min new tokens exist
    mn = MinNewTokensLengthLogitsProcessor(prompt_length_to_skip=0, min_new_tokens=8, eos_token_id=[2, 7])
    want = mn(generated, logits.clone())
    got = G._process_logits(logits.clone(), generated, step=6, min_new_tokens=8, eos_ids=[2, 7], repetition_penalty=1.0)
    assert torch.equal(torch.isinf(got), torch.isinf(want)) and torch.allclose(got[~torch.isinf(got)], want[~torch.isinf(want)])
    got = G._process_logits(logits.clone(), generated, step=8, min_new_tokens=8, eos_ids=[2, 7], repetition_penalty=1.0)
    assert not torch.isinf(got).any()
    # warpers, in HF order: temperature, top-k, top-p
    want = TopPLogitsWarper(0.8)(generated, TopKLogitsWarper(10)(generated, TemperatureLogitsWarper(0.7)(generated, logits.clone())))
    got = G._warp(logits.clone(), 0.7, 10, 0.8)
    assert torch.equal(torch.isinf(got), torch.isinf(want))
    assert torch.allclose(got[~torch.isinf(got)], want[~torch.isinf(want)])


def test_decoding_loops_match_huggingface_generate():
    """generate()'s greedy and beam-search loops (host logic: hypothesis bookkeeping, length
    penalty, early-stopping heuristic, min_new_tokens / repetition penalty on log-probs, EOS
    padding) against HuggingFace generate fed with inputs_embeds — the reference's call
    (eilev/model/v2.py:318-322) — on a tiny random OPT on CPU, through an injected stepper."""
    import types
    from transformers import OPTConfig, OPTForCausalLM
    from eilev_b200.model import generation as G

    class Stepper:
        start_token = None
        status = None

        def __init__(self, lm):
            self.lm, self.table = lm, lm.get_input_embeddings().weight.detach()

        def _logits(self):
            with torch.no_grad():
                return self.lm(inputs_embeds=self.emb).logits[:, -1].float()

        def prefill(self, input_ids, attention_mask, video_mask, feats, max_new):
            self.emb = self.table[input_ids]
            return self._logits()

        def graph(self, rows, dev):
            return None

        def step(self, tokens):
            self.emb = torch.cat([self.emb, self.table[tokens.view(-1)][:, None]], 1)
            return self._logits()

        def reorder(self, src):
            self.emb = self.emb[src]

    cases = [dict(num_beams=3), dict(num_beams=4, length_penalty=2.0, repetition_penalty=1.3),
             dict(num_beams=2, early_stopping=True), dict(num_beams=3, length_penalty=0.5, min_new_tokens=3, repetition_penalty=1.2),
             dict(num_beams=4, early_stopping="never"), dict(num_beams=1, min_new_tokens=2, repetition_penalty=1.5)]
    bad = []
    for seed in range(6):
        torch.manual_seed(seed)
        cfg = OPTConfig(hidden_size=16, num_hidden_layers=2, ffn_dim=32, num_attention_heads=2, vocab_size=24,
                        max_position_embeddings=64, word_embed_proj_dim=16, pad_token_id=1, eos_token_id=2, bos_token_id=0)
        lm = OPTForCausalLM(cfg).eval()
        for p in lm.parameters():
            p.data.normal_(0, 0.6)
        ids = torch.randint(3, 24, (2, 5))
        model = types.SimpleNamespace(config=types.SimpleNamespace(text_config=cfg, use_decoder_only_language_model=True),
                                      language_model=lm)
        for kw in cases:
            kw = dict(kw, max_new_tokens=7, do_sample=False, pad_token_id=1, eos_token_id=2)
            want = lm.generate(inputs_embeds=lm.get_input_embeddings()(ids).detach(), attention_mask=torch.ones_like(ids), **kw)
            got = G.generate(model, ids, torch.ones_like(ids), None, None, _stepper=Stepper(lm), **kw)
            n = max(got.shape[1], want.shape[1])
            got = torch.cat([got, torch.full((2, n - got.shape[1]), 1)], 1)
            want = torch.cat([want, torch.full((2, n - want.shape[1]), 1)], 1)
            if not torch.equal(got, want):
                bad.append((seed, kw, got.tolist(), want.tolist()))
    assert not bad, bad[:3]


def test_seq2seq_decoding_loops_match_huggingface_generate():
    """Same as above for the encoder-decoder branch (flan-T5: [decoder_start] + new tokens, the
    start token counts for the repetition penalty).  Rows are compared up to their first EOS:
    the pinned transformers 4.33.1 pads finished beams with pad_token_id (as this code does),
    the installed 5.5.0 fills with EOS when pad_token_id == 0 (`pad_token_id or eos_token_id`)."""
    import types
    from transformers import T5Config, T5ForConditionalGeneration
    from eilev_b200.model import generation as G

    class Stepper:
        status = None

        def __init__(self, lm):
            self.lm, self.start_token = lm, lm.config.decoder_start_token_id

        def _logits(self):
            with torch.no_grad():
                return self.lm(encoder_outputs=self.enc, attention_mask=self.am,
                               decoder_input_ids=self.prefix).logits[:, -1].float()

        def prefill(self, input_ids, attention_mask, video_mask, feats, max_new):
            with torch.no_grad():
                self.enc = self.lm.encoder(inputs_embeds=self.lm.shared(input_ids), attention_mask=attention_mask)
            self.am = attention_mask
            self.prefix = torch.full((input_ids.shape[0], 1), self.start_token, dtype=torch.long)
            return self._logits()

        def graph(self, rows, dev):
            return None

        def step(self, tokens):
            self.prefix = torch.cat([self.prefix, tokens.view(-1, 1)], 1)
            return self._logits()

        def reorder(self, src):
            self.prefix = self.prefix[src]

    def trim(t, width):
        t = torch.cat([t, torch.zeros((t.shape[0], width - t.shape[1]), dtype=torch.long)], 1)
        for r in range(t.shape[0]):
            e = (t[r, 1:] == 1).nonzero()
            if e.numel():
                t[r, int(e[0]) + 2:] = 0
        return t

    cases = [dict(num_beams=1), dict(num_beams=3), dict(num_beams=4, length_penalty=2.0),
             dict(num_beams=2, early_stopping=True, min_new_tokens=2), dict(num_beams=1, repetition_penalty=1.4)]
    bad = []
    for seed in range(6):
        torch.manual_seed(seed)
        cfg = T5Config(d_model=16, d_kv=8, d_ff=32, num_layers=2, num_decoder_layers=2, num_heads=2, vocab_size=20,
                       feed_forward_proj="gated-gelu", tie_word_embeddings=False, decoder_start_token_id=0,
                       pad_token_id=0, eos_token_id=1)
        lm = T5ForConditionalGeneration(cfg).eval()
        for n_, p in lm.named_parameters():
            if "layer_norm" not in n_:
                p.data.normal_(0, 0.5)
        ids = torch.randint(2, 20, (2, 6))
        am = torch.ones_like(ids)
        am[1, 4:] = 0
        model = types.SimpleNamespace(config=types.SimpleNamespace(text_config=cfg, use_decoder_only_language_model=False),
                                      language_model=lm)
        for kw in cases:
            kw = dict(kw, max_new_tokens=6, do_sample=False)
            want = lm.generate(inputs_embeds=lm.shared(ids).detach(), attention_mask=am, **kw)
            got = G.generate(model, ids, am, None, None, _stepper=Stepper(lm), **kw)
            n = max(got.shape[1], want.shape[1])
            if not torch.equal(trim(got, n), trim(want, n)):
                bad.append((seed, kw, got.tolist(), want.tolist()))
    assert not bad, bad[:3]


# ------------------------------------------------------------------------------------- v1 wrapper
def _v1_case(name):
    base = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = Blip2Config(**{k: base["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    return torch.load(GOLDEN / f"v1_{name}.pt", weights_only=False), base["state_dict"], cfg


def test_v1_prepend_video_slots_and_compact_left():
    from eilev_b200.model.v1 import compact_left, prepend_video_slots
    ids = torch.tensor([[5, 6, 7, 1], [1, 1, 8, 9]])
    am = torch.tensor([[1, 1, 1, 0], [0, 0, 1, 1]])
    lab = torch.tensor([[5, 6, 7, -100], [-100, -100, 8, 9]])
    fi, fa, fv, fl = prepend_video_slots(ids, am, lab, 3, 1, decoder_only=True)
    assert fi.tolist() == [[1, 1, 1, 5, 6, 7, 1], [1, 1, 1, 1, 1, 8, 9]]
    assert fa.tolist() == [[1, 1, 1, 1, 1, 1, 0], [1, 1, 1, 0, 0, 1, 1]]
    assert fv.tolist() == [[1, 1, 1, 0, 0, 0, 0]] * 2
    # HF 4.33.1 shifts the labels against the LAST L logits: labels[:, 0] is never a target
    assert fl.tolist() == [[-100, -100, -100, -100, 6, 7, -100], [-100, -100, -100, -100, -100, 8, 9]]
    _, _, _, same = prepend_video_slots(ids, None, lab, 3, 1, decoder_only=False)
    assert same is lab  # seq2seq labels are decoder targets
    ci, ca, cv = compact_left(fi, fa, fv)
    assert ca.tolist() == [[0, 1, 1, 1, 1, 1, 1], [0, 0, 1, 1, 1, 1, 1]]
    assert cv.tolist() == [[0, 1, 1, 1, 0, 0, 0], [0, 0, 1, 1, 1, 0, 0]]
    assert ci.tolist() == [[1, 1, 1, 1, 5, 6, 7], [1, 1, 1, 1, 1, 8, 9]]


@pytest.mark.parametrize("name", ["tiny_opt", "small_opt", "small_t5"])
def test_v1_wrapper_over_an_oracle_backed_v2_forward_matches_the_reference(name, monkeypatch):
    """The whole v1 wrapper on CPU: v2's ``forward`` (the CUDA engine) is replaced by the fp32
    oracle, so what is checked is exactly the host logic v1 adds — slot layout, label masking,
    logits slice, output packing — against the golden of the real eilev.model.v1 class."""
    from transformers.modeling_outputs import BaseModelOutputWithPooling, CausalLMOutputWithPast
    from transformers.models.blip_2.modeling_blip_2 import Blip2ForConditionalGenerationModelOutput

    from eilev_b200.model import v1, v2
    from oracle import videoblip_ref as R

    fx, sd, cfg = _v1_case(name)
    seen = {}

    def oracle_forward(self, input_ids, attention_mask=None, pixel_values=None, video_input_mask=None,
                       decoder_input_ids=None, decoder_attention_mask=None, output_attentions=None,
                       output_hidden_states=None, labels=None, return_dict=None):
        seen.update(input_ids=input_ids, video_input_mask=video_input_mask, labels=labels)
        assert return_dict is True
        if cfg.use_decoder_only_language_model:
            o = R.videoblip_forward(sd, cfg, input_ids, attention_mask, pixel_values, video_input_mask, labels)
        else:
            o = R.videoblip_forward_t5(sd, cfg, input_ids, attention_mask, pixel_values, video_input_mask, labels,
                                       decoder_input_ids)
        return Blip2ForConditionalGenerationModelOutput(
            loss=o.get("loss"), logits=o["logits"],
            vision_outputs=BaseModelOutputWithPooling(last_hidden_state=o["image_embeds"], pooler_output=o["pooler_output"]),
            qformer_outputs=BaseModelOutputWithPooling(last_hidden_state=o["query_output"], pooler_output=o["query_output"][:, 0]),
            language_model_outputs=CausalLMOutputWithPast(loss=o.get("loss"), logits=o["logits"]))

    monkeypatch.setattr(v2.VideoBlipForConditionalGeneration, "forward", oracle_forward)
    m = v1.VideoBlipForConditionalGeneration(cfg)
    with torch.no_grad():
        out = m(**fx["inputs"])
        tup = m(**fx["inputs"], return_dict=False)
    nq = cfg.num_query_tokens
    assert seen["input_ids"].shape[1] == fx["inputs"]["input_ids"].shape[1] + nq
    assert seen["video_input_mask"][:, :nq].all() and not seen["video_input_mask"][:, nq:].any()
    assert out.logits.shape == fx["logits"].shape
    assert torch.allclose(out.logits, fx["logits"], atol=5e-5, rtol=1e-4)
    assert abs(float(out.loss) - float(fx["loss"])) < 1e-5
    assert len(tup) == 5 and torch.equal(tup[1], out.logits) and float(tup[0]) == float(out.loss)
    assert out.language_model_outputs.logits.shape[1] == (
        fx["logits_no_labels"].shape[1] if cfg.use_decoder_only_language_model else fx["logits"].shape[1])
    with pytest.raises(ValueError):
        m(pixel_values=None, input_ids=fx["inputs"]["input_ids"])
    with pytest.raises(ValueError):
        m(pixel_values=fx["inputs"]["pixel_values"][:1], input_ids=fx["inputs"]["input_ids"])


def test_v1_generate_layout_and_kwargs(monkeypatch):
    from eilev_b200.model import v1, v2

    fx, sd, cfg = _v1_case("small_opt")
    calls = []

    def fake_generate(self, input_ids, pixel_values=None, video_input_mask=None, attention_mask=None, **kw):
        calls.append(dict(ids=input_ids, vm=video_input_mask, am=attention_mask, kw=kw))
        return torch.zeros(input_ids.shape[0], 1, dtype=torch.long)

    monkeypatch.setattr(v2.VideoBlipForConditionalGeneration, "generate", fake_generate)
    m = v1.VideoBlipForConditionalGeneration(cfg)
    g = fx["gen_inputs"]
    m.generate(**g, num_beams=4, top_p=0.9)
    c = calls[-1]
    nq, n_valid = cfg.num_query_tokens, g["attention_mask"].sum(1)
    assert c["kw"] == dict(num_beams=4, top_p=0.9)
    assert bool((c["am"][:, 1:] >= c["am"][:, :-1]).all())  # padding only at the left end
    assert c["am"].sum(1).tolist() == (n_valid + nq).tolist()
    for b in range(c["ids"].shape[0]):
        n = int(n_valid[b])
        assert c["ids"][b, -n:].tolist() == g["input_ids"][b, -n:].tolist()  # text stays last, in order
        assert c["vm"][b, -n - nq:-n].all() and int(c["vm"][b].sum()) == nq  # video slots right before it
    m.generate(pixel_values=g["pixel_values"])  # no prompt: [bos] per row (HF 4.33.1 generate)
    c = calls[-1]
    assert c["ids"].shape == (g["pixel_values"].shape[0], nq + 1)
    assert c["ids"][:, -1].tolist() == [cfg.text_config.bos_token_id] * g["pixel_values"].shape[0]
    assert bool(c["am"].all())


# ------------------------------------------------------------------------------------- uint8 frame path
def _image_only_processor(size):
    """Blip2Processor stand-in (no tokenizer files offline): the real BlipImageProcessor behind the
    processor call signature ``process`` uses."""
    from transformers import BatchEncoding, BlipImageProcessor
    ip = BlipImageProcessor(size={"height": size, "width": size})

    def call(images=None, text=None, return_tensors=None, **kw):
        return BatchEncoding(dict(ip(images=images, return_tensors=return_tensors, **kw)))
    call.image_processor = ip
    return call


@pytest.mark.parametrize("hw", [(56, 56), (40, 72)])  # already at the target size / needs the bicubic resize
def test_process_normalize_on_device_keeps_uint8_and_matches_the_stock_path(hw):
    """process(normalize_on_device=True) hands over the RESIZED uint8 frames; normalising them with the
    oracle restatement of BlipImageProcessor's rescale + normalize reproduces the stock float path."""
    from eilev_b200.model.utils import process
    from oracle import videoblip_ref as R
    proc = _image_only_processor(56)
    g = torch.Generator().manual_seed(1)
    video = torch.randint(0, 256, (2, 3, 4, *hw), dtype=torch.uint8, generator=g)
    stock = process(proc, video=video)["pixel_values"]
    u8 = process(proc, video=video, normalize_on_device=True)["pixel_values"]
    assert u8.dtype == torch.uint8 and u8.shape == stock.shape == (2, 3, 4, 56, 56)
    if hw == (56, 56):
        assert torch.equal(u8, video)
    assert float((R.normalize_frames(u8) - stock).abs().max()) < 1e-6
    with pytest.raises(ValueError):
        process(proc, video=video.float(), normalize_on_device=True)


def test_vision_model_frame_normalization_defaults_and_override():
    from transformers import BlipImageProcessor
    from eilev_b200.model.v2 import VideoBlipVisionModel
    from oracle import videoblip_ref as R
    m = VideoBlipVisionModel(small_cfg().vision_config)
    assert m.image_mean == R.OPENAI_CLIP_MEAN and m.image_std == R.OPENAI_CLIP_STD and m.rescale_factor == 1 / 255
    m.set_frame_normalization(BlipImageProcessor(image_mean=[0.5, 0.5, 0.5], image_std=[0.25, 0.5, 1.0]))
    assert m.image_mean == (0.5, 0.5, 0.5) and m.image_std == (0.25, 0.5, 1.0)
    assert VideoBlipVisionModel.image_mean == R.OPENAI_CLIP_MEAN  # class default untouched


# ------------------------------------------------------------------------------------- > 16 decode rows
def test_grouped_stepper_equals_one_stepper_over_all_rows():
    """generate() deals more than 16 rows (batch x beams) to groups of whole beam sets
    (_GroupedStepper).  With an injected per-group stepper over a tiny HF OPT the grouped run must
    reproduce the single-stepper run exactly: greedy, beam search (reorder stays inside a group),
    sampling; video features are split by the rows' slot counts; statuses add up."""
    import types
    from transformers import OPTConfig, OPTForCausalLM
    from eilev_b200.model import generation as G

    torch.manual_seed(0)
    cfg = OPTConfig(hidden_size=16, num_hidden_layers=2, ffn_dim=32, num_attention_heads=2, vocab_size=24,
                    max_position_embeddings=64, word_embed_proj_dim=16, pad_token_id=1, eos_token_id=2, bos_token_id=0)
    lm = OPTForCausalLM(cfg).eval()
    for p in lm.parameters():
        p.data.normal_(0, 0.6)
    seen = []

    class Stepper:
        start_token = None

        def __init__(self):
            self.table = lm.get_input_embeddings().weight.detach()

        def _logits(self):
            with torch.no_grad():
                return lm(inputs_embeds=self.emb).logits[:, -1].float()

        def prefill(self, input_ids, attention_mask, video_mask, feats, max_new):
            self.emb = self.table[input_ids].clone()
            if feats is not None:  # the reference's splice (v2.py:316): row-major over this stepper's rows
                self.emb[video_mask.bool()] = feats
            seen.append((input_ids.shape[0], None if feats is None else feats.shape[0]))
            self.status = torch.tensor([0, 0 if feats is None else feats.shape[0]], dtype=torch.int32)
            return self._logits()

        def graph(self, rows, dev):
            return None

        def step(self, tokens):
            self.emb = torch.cat([self.emb, self.table[tokens.view(-1)][:, None]], 1)
            return self._logits()

        def reorder(self, src):
            self.emb = self.emb[src]

    model = types.SimpleNamespace(config=types.SimpleNamespace(text_config=cfg, use_decoder_only_language_model=True),
                                  language_model=lm)
    batch = 7
    ids = torch.randint(3, 24, (batch, 6))
    vm = torch.zeros_like(ids)
    for b in range(batch):  # a different number of video slots per row
        vm[b, 1:1 + (b % 3)] = 1
    feats = torch.randn(int(vm.sum()), 16)
    am = torch.ones_like(ids)
    for kw in (dict(num_beams=1, do_sample=False), dict(num_beams=3, do_sample=False, length_penalty=1.3),
               dict(num_beams=5, early_stopping=True), dict(num_beams=1, do_sample=True, top_k=5, temperature=0.8)):
        kw = dict(kw, max_new_tokens=6, pad_token_id=1, eos_token_id=2)
        nb = kw["num_beams"]
        torch.manual_seed(5)
        want = G.generate(model, ids, am, vm, feats, _stepper=Stepper(), **kw)
        seen.clear()
        torch.manual_seed(5)
        grouped = G._GroupedStepper(Stepper, max(nb, 4 // nb * nb))  # small groups: several per call
        got = G.generate(model, ids, am, vm, feats, _stepper=grouped, **kw)
        assert torch.equal(got, want), (kw, got.tolist(), want.tolist())
        assert len(seen) > 1 and sum(r for r, _ in seen) == batch * nb
        assert sum(f for _, f in seen) == int(vm.sum()) * nb
        assert grouped.status.tolist() == [0, int(vm.sum()) * nb]
    with pytest.raises(RuntimeError):  # a permutation that leaves its group is refused
        grouped.reorder(torch.arange(batch * nb - 1, -1, -1))


@pytest.mark.parametrize("timestamp,expected", [("00:00:00.560", 0.56), ("00:00:49.15", 49.15),
                                                ("00:06:50.039", 410.039), ("02:06:50.039", 7610.039)])
def test_parse_timestamp_reference_table(timestamp, expected):
    """tests/data/test_utils.py:865-875 of the reference."""
    from eilev_b200.data.utils import parse_timestamp
    assert parse_timestamp(timestamp) == expected


def test_generate_chunks():
    from eilev_b200.data.utils import generate_chunks
    assert list(generate_chunks(list(range(7)), 3)) == [[0, 1, 2], [3, 4, 5], [6]]
    assert list(generate_chunks([], 3)) == [] and list(generate_chunks([1, 2], 2)) == [[1, 2]]


# ------------------------------------------------------------------------------------- frame directory format
def test_frame_dataset_reads_the_extracted_frame_directory_format(tmp_path):
    """eilev/data/frame.py:14-72 FrameDataset on a directory laid out as scripts/ego4d/extract_frames.py
    writes it: narrated_actions.csv + one sub-directory of PNG frames per row."""
    import csv
    import numpy as np
    from PIL import Image
    from eilev_b200.data.frame import FrameDataset, read_frame_dir
    rs = np.random.RandomState(0)
    cols = ["frame_path", "video_uid", "clip_index", "narration_timestamp_sec", "narration_text",
            "structured_verb", "structured_noun"]
    rows, frames = [], {}
    for uid, clip, n in (("vidA", 0, 8), ("vidA", 3, 12), ("vidB", 1, 8)):
        fp = f"{uid}|{clip}"
        (tmp_path / fp).mkdir()
        frames[fp] = rs.randint(0, 256, (n, 20, 28, 3)).astype(np.uint8)
        for i in range(n):  # unpadded indices: '…|10.png' must come after '…|9.png'
            Image.fromarray(frames[fp][i]).save(tmp_path / fp / f"{fp}|{i}.png")
        rows.append(dict(zip(cols, [fp, uid, clip, 1.5 * clip, f"#C C does {clip}", f"verb{clip}", "noun"])))
    with open(tmp_path / "narrated_actions.csv", "w", newline="") as f:
        w = csv.DictWriter(f, cols)
        w.writeheader()
        w.writerows(rows)
    ds = FrameDataset(str(tmp_path))
    assert len(ds) == 3
    item = ds[1]
    assert item["video"].dtype == torch.uint8 and item["video"].shape == (3, 12, 20, 28)
    assert torch.equal(item["video"], torch.from_numpy(frames["vidA|3"]).permute(3, 0, 1, 2))
    assert item["narration_text"] == "#C C does 3" and item["clip_index"] == "3"  # CSV fields stay strings
    assert torch.equal(ds["vidB|1"]["video"], read_frame_dir(tmp_path / "vidB|1"))
    only_b = FrameDataset(str(tmp_path), data_filter=lambda r: r["video_uid"] == "vidB",
                          transform=lambda it: {**it, "n": it["video"].shape[1]}, return_frames=True)
    assert len(only_b) == 1 and only_b[0]["n"] == 8
    assert "video" not in FrameDataset(str(tmp_path), return_frames=False)[0]
    with pytest.raises(AssertionError):
        FrameDataset(str(tmp_path / "vidA|0"))  # no narrated_actions.csv


def test_uint8_frames_flow_from_the_frame_directory_to_the_collated_batch(tmp_path):
    """Host pipeline of the uint8 frame path, end to end on CPU: FrameDataset (PNG dirs) ->
    process(normalize_on_device=True) (resize only, stays uint8) -> interleaved tokeniser -> collator.
    The collated ``pixel_values`` are the uint8 clips in batch order — what the model's fused
    normalisation consumes — and normalising them with the oracle equals the stock float path."""
    import csv
    import numpy as np
    from PIL import Image
    from eilev_b200.data.frame import FrameDataset
    from eilev_b200.data.utils import (DataCollatorForInterleavedVideoSeq2Seq,
                                       generate_input_ids_and_labels_from_interleaved as gen)
    from eilev_b200.model.utils import process
    from oracle import videoblip_ref as R
    rs = np.random.RandomState(1)
    cols = ["frame_path", "video_uid", "clip_index", "narration_timestamp_sec", "narration_text",
            "structured_verb", "structured_noun"]
    with open(tmp_path / "narrated_actions.csv", "w", newline="") as f:
        w = csv.DictWriter(f, cols)
        w.writeheader()
        for k in range(3):
            fp = f"vid|{k}"
            (tmp_path / fp).mkdir()
            for i in range(2):
                Image.fromarray(rs.randint(0, 256, (40, 72, 3)).astype(np.uint8)).save(tmp_path / fp / f"{fp}|{i}.png")
            w.writerow(dict(zip(cols, [fp, "vid", k, 0.0, "A text", "v", "n"])))
    proc = _image_only_processor(56)
    tok = opt_tok()
    ds = FrameDataset(str(tmp_path))

    def datapoint(clips, on_device):
        video = torch.stack([ds[c]["video"] for c in clips])  # (N, C, T, H, W) uint8
        enc = gen(tok, [("A prompt", len(clips))], "A text", 2, True)
        px = process(proc, video=video, normalize_on_device=on_device)["pixel_values"]
        return {**enc, "pixel_values": px}

    col = DataCollatorForInterleavedVideoSeq2Seq(tok, pad_to_multiple_of=8)
    u8 = col([datapoint([0, 1], True), datapoint([2], True)])
    f32 = col([datapoint([0, 1], False), datapoint([2], False)])
    assert u8["pixel_values"].dtype == torch.uint8 and u8["pixel_values"].shape == (3, 3, 2, 56, 56)
    assert f32["pixel_values"].dtype == torch.float32
    for k in ("input_ids", "attention_mask", "video_input_mask", "labels"):
        assert torch.equal(u8[k], f32[k])
    assert int(u8["video_input_mask"].sum()) == 3 * 2  # 3 clips x 2 query tokens
    assert float((R.normalize_frames(u8["pixel_values"]) - f32["pixel_values"]).abs().max()) < 1e-6


def test_forward_signatures_expose_the_collator_keys_to_hf_trainer():
    """HF Trainer keeps the dataset columns named in ``inspect.signature(model.forward)``
    (Trainer._set_signature_columns_if_needed) and calls ``model(**inputs)``: the v2 signature must
    name exactly the interleaved collator's keys (eilev/model/v2.py:132-144), the v1 signature the
    HF 4.33.1 Blip2 ones (pixel_values first)."""
    import inspect
    from eilev_b200.model import v1, v2
    p2 = list(inspect.signature(v2.VideoBlipForConditionalGeneration.forward).parameters)
    assert p2 == ["self", "input_ids", "attention_mask", "pixel_values", "video_input_mask", "decoder_input_ids",
                  "decoder_attention_mask", "output_attentions", "output_hidden_states", "labels", "return_dict"]
    p1 = list(inspect.signature(v1.VideoBlipForConditionalGeneration.forward).parameters)
    assert p1 == ["self", "pixel_values", "input_ids", "attention_mask", "decoder_input_ids", "decoder_attention_mask",
                  "output_attentions", "output_hidden_states", "labels", "return_dict"]
    g2 = list(inspect.signature(v2.VideoBlipForConditionalGeneration.generate).parameters)
    assert g2[:5] == ["self", "input_ids", "pixel_values", "video_input_mask", "attention_mask"]  # v2.py:254-261
    c2 = list(inspect.signature(v2.VideoBlipForConditionalGeneration.classify).parameters)
    assert c2 == ["self", "prompt_input_ids", "class_input_ids", "prompt_attention_mask", "pixel_values",
                  "prompt_video_input_mask", "class_attention_mask", "class_batch_size"]  # v2.py:326-336


def test_flan_style_checkpoint_keeps_its_own_head_through_save_and_load(tmp_path):
    """A flan-T5-style checkpoint (separate lm_head.weight, as the published eilev-blip2-flan-t5-xl has)
    survives save_pretrained -> from_pretrained here, and — when the reference checkout is present —
    loads into the real reference class with the same separate head (transformers 5.x un-ties a head
    that the checkpoint carries)."""
    import types
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    fx = torch.load(GOLDEN / "small_t5.pt", weights_only=False)
    cfg = Blip2Config(**{k: fx["config"][k] for k in ("vision_config", "qformer_config", "text_config", "num_query_tokens")})
    m = VideoBlipForConditionalGeneration(cfg)
    sd = {k: v.clone() for k, v in fx["state_dict"].items()}
    sd["language_model.lm_head.weight"] = sd["language_model.lm_head.weight"] + 0.5  # a head of its own
    m.load_state_dict(sd)
    lm = m.language_model
    assert lm.lm_head.weight.data_ptr() != lm.shared.weight.data_ptr()
    m.save_pretrained(tmp_path)
    m2 = VideoBlipForConditionalGeneration.from_pretrained(tmp_path)
    assert torch.equal(m2.language_model.lm_head.weight, sd["language_model.lm_head.weight"])
    assert torch.equal(m2.language_model.shared.weight, sd["language_model.shared.weight"])
    if Path("/root/reference/eilev/model/v2.py").exists():
        sys.path.insert(0, "/root/reference")
        sys.modules.setdefault("pytorchvideo", types.ModuleType("pytorchvideo"))
        from eilev.model.v2 import VideoBlipForConditionalGeneration as Ref
        r = Ref.from_pretrained(tmp_path)
        assert torch.equal(r.language_model.lm_head.weight, sd["language_model.lm_head.weight"])
        assert torch.equal(r.language_model.shared.weight, sd["language_model.shared.weight"])


def test_fused_gradient_views_require_adjacent_flat_buffer_slices():
    """engine/qformer.py::_GradOut._fused_view: the q / k / v gradients of one layer are written by ONE accumulating
    GEMM only when their sink views sit back to back in the trainer's flat buffer; anything else falls back to
    per-parameter accumulation."""
    from eilev_b200.engine.qformer import _GradOut
    flat = torch.zeros(3 * 8 * 4 + 3 * 8 + 5)
    wq, wk, wv = (flat[i * 32:(i + 1) * 32].view(8, 4) for i in range(3))
    bq, bk, bv = (flat[96 + i * 8:96 + (i + 1) * 8] for i in range(3))
    g = _GradOut({"q.w": wq, "k.w": wk, "v.w": wv, "q.b": bq, "k.b": bk, "v.b": bv, "odd": flat[121:125]},
                 torch.device("cpu"))
    w = g._fused_view(["q.w", "k.w", "v.w"])
    assert w is not None and w.shape == (24, 4) and w.data_ptr() == wq.data_ptr()
    w += 1.0   # writes through to all three views and nothing else
    assert float(wq.sum() + wk.sum() + wv.sum()) == 96.0 and float(flat.sum()) == 96.0
    b = g._fused_view(["q.b", "k.b", "v.b"])
    assert b is not None and b.shape == (24,) and b.data_ptr() == bq.data_ptr()
    assert g._fused_view(["q.w", "v.w", "k.w"]) is None          # wrong order: not adjacent
    assert g._fused_view(["q.w", "k.w", "missing"]) is None      # not in the sink
    assert g._fused_view(["k.b", "v.b", "odd"]) is None          # a gap of one element before "odd"
    assert g._fused_view(["q.w", "q.b"]) is None                 # different trailing shapes


def test_flat_buffer_keeps_a_layers_qkv_gradients_adjacent():
    """The trainer's flat buffer lists parameters in model order inside its decay / no-decay segments, which puts the
    query / key / value weights (and biases) of a Q-Former layer back to back: the fused projection's gradient then
    goes in with one accumulating launch per layer (engine/qformer.py::_GradOut.weights_fused)."""
    from eilev_b200.engine import qformer as E_qf
    from eilev_b200.model.v2 import VideoBlipForConditionalGeneration
    from eilev_b200.train import FlatBuffers, freeze_for_recipe
    m = VideoBlipForConditionalGeneration(small_cfg())
    freeze_for_recipe(m)
    FlatBuffers(m.named_parameters())
    sink = {n: p.grad for n, p in E_qf.qformer_param_list(m) if p.requires_grad and p.grad is not None}
    g = E_qf._GradOut(sink, torch.device("cpu"))
    n_layers = m.qformer.config.num_hidden_layers
    for i in range(n_layers):
        pre = f"qformer.encoder.layer.{i}.attention.attention."
        w = g._fused_view([pre + f"{nm}.weight" for nm in ("query", "key", "value")])
        b = g._fused_view([pre + f"{nm}.bias" for nm in ("query", "key", "value")])
        dq = sink[pre + "query.weight"].shape[0]
        assert w is not None and w.shape == (3 * dq, sink[pre + "query.weight"].shape[1]), i
        assert b is not None and b.shape == (3 * dq,), i
