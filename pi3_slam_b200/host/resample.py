"""Host-built tap tables for pi3b_resample2d (per-shape constants, like the pos-embed resize).

`bilinear_taps` follows ATen's upsample_bilinear2d (align_corners=False): src = scale*(dst+0.5)-0.5 clamped at 0, two
taps (reference use: F.interpolate(..., mode='bilinear', antialias=False), moge/model/v2.py:157).
`bilinear_aa_taps` follows ATen's _upsample_bilinear2d_aa (separable triangle filter whose support grows with the
down-scale factor; reference use: F.interpolate(..., antialias=True), moge/model/modules.py:121).
Both return (start int32 [out], weights fp32 [out, taps]); taps past the valid range carry weight 0.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


def bilinear_taps(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    scale = np.float32(in_size / out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = scale * (dst + np.float32(0.5)) - np.float32(0.5)
    src = np.maximum(src, np.float32(0.0)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int32), in_size - 1)
    lam1 = (src - i0.astype(np.float32)).astype(np.float32)
    has_next = i0 < in_size - 1
    w = np.zeros((out_size, 2), dtype=np.float32)
    w[:, 0] = np.float32(1.0) - lam1
    w[:, 1] = np.where(has_next, lam1, np.float32(0.0))
    # at the last source row ATen reads index i0 twice: fold the second weight into the first
    w[:, 0] = np.where(has_next, w[:, 0], np.float32(1.0))
    return i0.astype(np.int32), w


def bilinear_aa_taps(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    scale = np.float32(in_size / out_size)
    support = np.float32(scale) if scale >= 1.0 else np.float32(1.0)
    invscale = np.float32(1.0 / scale) if scale >= 1.0 else np.float32(1.0)
    taps = int(math.ceil(float(support))) * 2 + 1
    start = np.zeros(out_size, dtype=np.int32)
    w = np.zeros((out_size, taps), dtype=np.float32)
    for i in range(out_size):
        center = np.float32(scale * np.float32(i + 0.5))
        xmin = max(int(np.float32(center - support + np.float32(0.5))), 0)
        xsize = min(int(np.float32(center + support + np.float32(0.5))), in_size) - xmin
        j = np.arange(xsize, dtype=np.float32)
        arg = (j + np.float32(xmin) - center + np.float32(0.5)) * invscale
        wj = np.maximum(np.float32(0.0), np.float32(1.0) - np.abs(arg)).astype(np.float32)
        total = np.float32(wj.sum(dtype=np.float32))
        if total != 0:
            wj = (wj / total).astype(np.float32)
        start[i] = xmin
        w[i, :xsize] = wj
    return start, w


def apply_taps_np(img: np.ndarray, tab_y, tab_x) -> np.ndarray:
    """Reference application of the tables to a (C, H, W) array (used by the CPU tests of the tables)."""
    (y0, wy), (x0, wx) = tab_y, tab_x
    C, H, W = img.shape
    xi = np.minimum(x0[:, None] + np.arange(wx.shape[1])[None], W - 1)
    tmp = (img[:, :, xi] * wx[None, None]).sum(-1, dtype=np.float32)          # (C, H, Wo)
    yi = np.minimum(y0[:, None] + np.arange(wy.shape[1])[None], H - 1)
    return (tmp[:, yi, :] * wy[None, :, :, None]).sum(2, dtype=np.float32)     # (C, Ho, Wo)


# ---------------------------------------------------------------------------------------------------------------
# Pillow's 8-bit resampling (what torchvision.transforms.Resize does to a PIL image; reference loader
# datasets/image_datasets.py:193-199, utils/image_utils.py:98-103). Pillow == 12.x src/libImaging/Resample.c:
# precompute_coeffs (double), normalize_coeffs_8bpc (fixed point, PRECISION_BITS = 32 - 8 - 2), horizontal pass to a
# uint8 image, then vertical pass. The tables below are those fixed-point coefficients; pi3b_resize_u8_pil applies
# them with the same integer arithmetic, so the device result is bit-identical to `Image.resize(size, BILINEAR)`.
# ---------------------------------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(xmin int32 [out], xcount int32 [out], coeffs int32 [out, ksize]) of Pillow's BILINEAR filter (support 1.0)."""
    scale = filterscale = float(in_size) / float(out_size)
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, dtype=np.int32)
    cnt_a = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0.0 else a
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            for x in range(xmax):       # Pillow divides element by element (not by a reciprocal)
                kk[xx, x] /= ww
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    fixed = np.where(kk < 0, (-0.5 + kk * (1 << PIL_PRECISION_BITS)), (0.5 + kk * (1 << PIL_PRECISION_BITS)))
    return xmin_a, cnt_a, np.trunc(fixed).astype(np.int32)
