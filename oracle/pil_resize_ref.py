"""ORACLE — test infrastructure only.  NOT part of the product path.

A numpy restatement of the antialiased bicubic resize the reference's frame path performs:
``eilev/model/utils.py:5-26`` ``process`` -> ``Blip2Processor`` -> ``BlipImageProcessor.resize`` of the
pinned transformers 4.33.1 -> ``image_transforms.resize`` -> ``PIL.Image.resize(size, BICUBIC)``.
The arithmetic lives in an un-vendored third-party dependency, Pillow (``src/libImaging/Resample.c``;
the 8-bit path has been unchanged since Pillow 7, the installed 12.2.0 is the executable stand-in);
its published algorithm is restated here:

  * ``precompute_coeffs``: per output pixel a window [xmin, xmin + n) and n double-precision weights
    of the Keys cubic (a = -0.5), the filter stretched by the scale when down-sampling
    (antialias), weights normalised to sum 1;
  * ``normalize_coeffs_8bpc``: weights to 22-bit fixed point, rounded half away from zero;
  * horizontal pass over the rows the vertical pass needs, then the vertical pass, each
    ``clip8((2**21 + sum(pixel * weight)) >> 22)`` in 32-bit integers with a uint8 intermediate.

Parity status: PINNED bit-exactly against ``PIL.Image.resize`` in tests/test_oracle.py (Pillow
travels with the image, so the pin also runs on the GPU box).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: PRECISION_BITS


def bicubic_filter(x: float) -> float:
    """Resample.c ``bicubic_filter`` (a = -0.5, support 2)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the full-image box.
    Returns (ksize, bounds (out, 2) int32 = [xmin, count], kk (out, ksize) int32 fixed-point weights)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)  # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_last_axis(planes: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """planes (..., in) uint8 -> (..., out) uint8 along the last axis."""
    out_size, ksize = kk.shape
    idx = bounds[:, :1] + np.arange(ksize, dtype=np.int32)[None, :]           # (out, ksize)
    idx = np.minimum(idx, planes.shape[-1] - 1)                                # taps past the count carry weight 0
    taps = planes[..., idx].astype(np.int32)                                   # (..., out, ksize)
    acc = (1 << (PRECISION_BITS - 1)) + (taps * kk[None].astype(np.int32)).sum(axis=-1, dtype=np.int32)
    return _clip8(acc)


def resize_bicubic_u8(planes: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """planes (..., H, W) uint8 -> (..., out_h, out_w) uint8, every plane resized independently as
    ``PIL.Image.fromarray(plane).resize((out_w, out_h), PIL.Image.BICUBIC)`` does (RGB bands are
    independent in Pillow, so this also is the per-channel result for an RGB image).
    ImagingResampleInner: horizontal pass first (only when the width changes), restricted to the
    source rows the vertical pass reads; vertical pass only when the height changes."""
    planes = np.ascontiguousarray(planes)
    assert planes.dtype == np.uint8
    in_h, in_w = planes.shape[-2:]
    cur = planes
    need_h, need_v = out_w != in_w, out_h != in_h
    if need_v:
        _, bounds_v, kk_v = precompute_coeffs(in_h, out_h)
        first = int(bounds_v[0, 0])
        last = int(bounds_v[-1, 0] + bounds_v[-1, 1])
    if need_h:
        _, bounds_h, kk_h = precompute_coeffs(in_w, out_w)
        if need_v:  # the horizontal pass only produces the rows the vertical pass needs
            cur = cur[..., first:last, :]
            bounds_v = bounds_v.copy()
            bounds_v[:, 0] -= first
        cur = _resample_last_axis(cur, bounds_h, kk_h)
    if need_v:
        cur = np.swapaxes(_resample_last_axis(np.swapaxes(cur, -1, -2), bounds_v, kk_v), -1, -2)
    return np.ascontiguousarray(cur)
