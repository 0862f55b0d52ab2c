"""Host side of the training-time preprocessing (scripts/general/train_v2.py:45-75,143-167)."""
import random

import torch

from eilev_b200.data import preprocess as P


def test_uniform_temporal_subsample_matches_the_documented_indices():
    clip = torch.arange(3 * 20 * 2 * 2, dtype=torch.uint8).view(3, 20, 2, 2)
    out = P.uniform_temporal_subsample(clip, 8)
    want = torch.linspace(0, 19, 8).long()
    assert out.shape == (3, 8, 2, 2)
    assert torch.equal(out, clip[:, want])
    assert int(want[0]) == 0 and int(want[-1]) == 19  # first and last frame are kept
    assert torch.equal(P.uniform_temporal_subsample(clip[:, :4], 8)[:, :, 0, 0], clip[:, [0, 0, 0, 1, 1, 2, 2, 3], 0, 0])


def test_crop_box_sampler_equals_torchvision_draw_for_draw():
    from torchvision.transforms import RandomResizedCrop
    for seed in range(25):
        h, w = 180 + 7 * seed, 240 + 11 * seed
        torch.manual_seed(seed)
        want = RandomResizedCrop.get_params(torch.empty(3, h, w), scale=(0.5, 1.0), ratio=(3 / 4, 4 / 3))
        torch.manual_seed(seed)
        got = P.resized_crop_params(h, w, (0.5, 1.0), (3 / 4, 4 / 3))
        assert tuple(want) == tuple(got), (seed, want, got)
        top, left, ch, cw = got
        assert 0 <= top and top + ch <= h and 0 <= left and left + cw <= w
        assert 0.45 * h * w <= ch * cw <= h * w and 0.70 <= cw / ch <= 1.40
    # extreme aspect ratio: the rejection loop can fail, the fallback is the ratio-clamped centre crop
    torch.manual_seed(0)
    assert P.resized_crop_params(10, 1000, (0.9, 1.0), (3 / 4, 4 / 3)) == RandomResizedCropFallback(10, 1000)


def RandomResizedCropFallback(height, width, ratio=(3 / 4, 4 / 3)):
    h, w = height, int(round(height * max(ratio)))
    return (height - h) // 2, (width - w) // 2, h, w


class _Tok:
    """Minimal tokenizer stand-in (the real-tokenizer cases live in tests/test_data_utils_golden.py)."""
    bos_token_id, eos_token_id, pad_token_id = 2, 2, 1

    def __call__(self, text, add_special_tokens=True, **kw):
        ids = [10 + (ord(c) % 50) for c in text]
        return type("E", (), {"input_ids": ([self.bos_token_id] if add_special_tokens else []) + ids})()


def test_preprocessor_contract():
    """train_v2.py:52-75: one video per item, the last item is the query (bare prompt), its narration the target;
    every prompt is one of PROMPTS; the transformed clips are stacked in item order."""
    calls = []

    def transform(v):
        calls.append(int(v[0, 0, 0, 0]))
        return v.float() + 0.5

    pre = P.Preprocessor(_Tok(), num_query_tokens=4, decoder_only_lm=True, video_transform=transform)
    items = [dict(narration_text=f"#C C does thing {i}", video=torch.full((3, 2, 4, 4), i, dtype=torch.uint8))
             for i in range(3)]
    random.seed(5)
    out = pre(dict(items=items))
    assert calls == [0, 1, 2]
    assert out["pixel_values"].shape == (3, 3, 2, 4, 4) and out["pixel_values"].dtype == torch.float32
    assert torch.equal(out["pixel_values"][2], torch.full((3, 2, 4, 4), 2.5))
    assert set(out) >= {"input_ids", "labels", "video_input_mask", "pixel_values"}
    assert int(out["video_input_mask"].sum()) == 3 * 4  # 3 videos x num_query_tokens slots
    # the same random stream gives the same prompts as the reference's expression order
    random.seed(5)
    want_prompts = [random.choice(P.PROMPTS) for _ in range(3)]
    from eilev_b200.data.utils import clean_narration_text, generate_input_ids_and_labels_from_interleaved
    ref = generate_input_ids_and_labels_from_interleaved(
        _Tok(), [(want_prompts[i] + " " + clean_narration_text(items[i]["narration_text"]), 1) for i in range(2)]
        + [(want_prompts[2], 1)], clean_narration_text(items[2]["narration_text"]), 4, True)
    assert torch.equal(out["input_ids"], ref["input_ids"]) and torch.equal(out["labels"], ref["labels"])
    assert len(P.PROMPTS) == 9 and all("camera wearer" in p for p in P.PROMPTS)


def test_host_part_of_the_train_transform():
    tr = P.TrainVideoTransform((224, 224), [0.48, 0.45, 0.40], [0.26, 0.26, 0.27], num_frames=8, rand_augment=False)
    clip = torch.randint(0, 256, (3, 20, 120, 160), dtype=torch.uint8)
    torch.manual_seed(3)
    out, params = tr.host_part(clip)
    assert out.dtype == torch.uint8 and out.shape == (3, 8, 120, 160) and out.is_contiguous()
    assert torch.equal(out, clip[:, torch.linspace(0, 19, 8).long()])
    top, left, h, w = params["box"]
    assert 0 <= top and top + h <= 120 and 0 <= left and left + w <= 160 and isinstance(params["flip"], bool)
    # with RandAugment: still uint8, same shape, one augmentation draw shared by the frames of the clip
    tr2 = P.TrainVideoTransform((224, 224), [0.48, 0.45, 0.40], [0.26, 0.26, 0.27], num_frames=8)
    same = torch.randint(0, 256, (3, 1, 64, 64), dtype=torch.uint8).expand(3, 8, 64, 64).contiguous()
    torch.manual_seed(1)
    aug, _ = tr2.host_part(same)
    assert aug.dtype == torch.uint8 and aug.shape == (3, 8, 64, 64)
    assert all(torch.equal(aug[:, 0], aug[:, t]) for t in range(1, 8))
