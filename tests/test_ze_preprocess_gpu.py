"""Device side of the training-time frame transform (vb_crop_resize_normalize_u8) against the oracle
(oracle/frame_transforms_ref.py: the torch functions pytorchvideo's transforms call, in the reference's order):
/255 -> Normalize -> crop -> F.interpolate(bicubic) -> hflip.  Tolerance: fp32, |d| <= 2e-4 on values of
magnitude <= ~3 (the kernel normalises after the 16-tap sum, the reference before it: same real arithmetic,
different rounding; bicubic overshoot makes the values leave [0, 1] in both)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

MEAN, STD = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]


@pytest.mark.parametrize("h,w,box,size,flip", [
    (240, 320, (10, 20, 200, 260), (224, 224), False),      # downscale
    (240, 320, (0, 0, 240, 320), (224, 224), True),         # whole frame, flipped
    (120, 160, (7, 33, 90, 101), (224, 224), True),         # upscale, odd box
    (224, 224, (0, 0, 224, 224), (224, 224), False),        # identity resize: taps hit the grid exactly
    (300, 200, (100, 50, 37, 149), (64, 96), False),        # anisotropic
])
def test_crop_resize_normalize_matches_the_torch_reference(h, w, box, size, flip):
    from eilev_b200 import ops
    from oracle.frame_transforms_ref import train_clip_transform
    g = torch.Generator().manual_seed(h * 7 + w)
    clip = torch.randint(0, 256, (3, 4, h, w), dtype=torch.uint8, generator=g)
    want = train_clip_transform(clip, box, size, MEAN, STD, flip)
    got = ops.crop_resize_normalize_u8(clip.cuda(), box, size, 1.0 / 255.0, MEAN, STD, flip=flip)
    assert got.shape == want.shape and got.dtype == torch.float32
    err = (got.cpu() - want).abs().max().item()
    assert err <= 2e-4, err
    bf = ops.crop_resize_normalize_u8(clip.cuda(), box, size, 1.0 / 255.0, MEAN, STD, flip=flip, dtype=torch.bfloat16)
    assert (bf.float().cpu() - want).abs().max().item() <= 0.02


def test_train_transform_end_to_end_and_model_accepts_it():
    """TrainVideoTransform (host draws + device pass) == the oracle with the same draws; the result feeds
    VideoBlipVisionModel like the reference's float clips do."""
    from eilev_b200.data.preprocess import TrainVideoTransform
    from oracle.frame_transforms_ref import train_clip_transform
    tr = TrainVideoTransform((224, 224), MEAN, STD, num_frames=8, rand_augment=False)
    clip = torch.randint(0, 256, (3, 19, 180, 250), dtype=torch.uint8)
    torch.manual_seed(11)
    sub, params = tr.host_part(clip)
    want = train_clip_transform(sub, params["box"], (224, 224), MEAN, STD, params["flip"])
    torch.manual_seed(11)
    got = tr(clip)
    assert got.is_cuda and got.shape == (3, 8, 224, 224)
    assert (got.cpu() - want).abs().max().item() <= 2e-4
    with pytest.raises(Exception):
        tr.device_part(sub, params)  # CPU tensor: the device pass has no CPU fallback
