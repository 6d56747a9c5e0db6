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
