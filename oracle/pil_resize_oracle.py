"""ORACLE — test infrastructure only. Never imported by the product path (pi3_slam_b200/).

numpy restatement of Pillow's 8-bit BILINEAR `Image.resize` (libImaging/Resample.c: ImagingResampleHorizontal_8bpc
then ImagingResampleVertical_8bpc on the fixed-point coefficients) followed by torchvision's ToTensor — the reference
loader of datasets/image_datasets.py:193-199. Pinned against Pillow itself (the reference's own dependency, installed
in this image) by tests/test_host.py, bit for bit.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _pass(img: np.ndarray, xmin, cnt, coeffs, axis: int) -> np.ndarray:
    """img uint8 (H, W, C); resample along axis (1 = horizontal, 0 = vertical)."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)            # (in, other, C)
    out = np.empty((len(xmin),) + img.shape[1:], dtype=np.uint8)
    for o in range(len(xmin)):
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(int(cnt[o])):
            acc += img[xmin[o] + j] * int(coeffs[o, j])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int, tables_h, tables_w) -> np.ndarray:
    """uint8 (H, W, 3) -> uint8 (out_h, out_w, 3); horizontal pass first, like Pillow; a pass is skipped when that
    dimension is unchanged."""
    if img.shape[1] != out_w:
        img = _pass(img, *tables_w, axis=1)
    if img.shape[0] != out_h:
        img = _pass(img, *tables_h, axis=0)
    return img


def to_tensor(img_u8: np.ndarray) -> np.ndarray:
    """torchvision ToTensor: HWC uint8 -> CHW float32 / 255."""
    return np.ascontiguousarray(img_u8.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
