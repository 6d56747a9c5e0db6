"""Target-size rule and image loading (reference: utils/image_utils.py:13-104)."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple, Union

import torch


def target_size_from_dims(w_orig: int, h_orig: int, pixel_limit: int = 255000) -> Tuple[int, int]:
    """(H, W) multiples of 14 with H*W <= pixel_limit, aspect preserved (image_utils.py:41-50)."""
    scale = math.sqrt(pixel_limit / (w_orig * h_orig)) if w_orig * h_orig > 0 else 1
    w_t, h_t = w_orig * scale, h_orig * scale
    k, m = round(w_t / 14), round(h_t / 14)
    while (k * 14) * (m * 14) > pixel_limit:
        if k / m > w_t / h_t:
            k -= 1
        else:
            m -= 1
    return max(1, m) * 14, max(1, k) * 14


def calculate_target_size(first_image_path, pixel_limit: int = 255000) -> Tuple[int, int]:
    """Same contract as the reference: opens the first image only to read its size."""
    if isinstance(first_image_path, tuple):
        raise NotImplementedError("video inputs (torchcodec) are outside the chunk hot path of this build")
    from PIL import Image
    with Image.open(first_image_path) as im:
        w, h = im.size
    return target_size_from_dims(w, h, pixel_limit)


def load_image(path: str, target_size: Tuple[int, int]) -> torch.Tensor:
    """PIL decode -> RGB -> Resize((H, W)) (bilinear, antialias as torchvision does for PIL) -> [0,1] CHW fp32
    (image_datasets.py:193-199)."""
    from PIL import Image
    import numpy as np
    h, w = target_size
    with Image.open(path) as im:
        im = im.convert("RGB")
        if im.size != (w, h):
            im = im.resize((w, h), Image.BILINEAR)
        arr = np.asarray(im, dtype=np.uint8)
    return torch.from_numpy(arr.copy()).permute(2, 0, 1).float().div_(255.0)


def load_images_from_paths(paths: Sequence[str], target_size: Tuple[int, int]) -> torch.Tensor:
    return torch.stack([load_image(p, target_size) for p in paths])


def decode_image_u8(path: str):
    """PIL decode -> RGB uint8 (H, W, 3) numpy at the source resolution (no resize: that runs on the device)."""
    from PIL import Image
    import numpy as np
    with Image.open(path) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


def resize_u8_on_host(arr, target_size: Tuple[int, int]) -> torch.Tensor:
    """Host fallback for frames whose source size differs from the rest of the request: the reference's own
    PIL resize + ToTensor."""
    from PIL import Image
    import numpy as np
    h, w = target_size
    im = Image.fromarray(arr)
    if im.size != (w, h):
        im = im.resize((w, h), Image.BILINEAR)
    return torch.from_numpy(np.asarray(im, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255.0)


def load_images_u8(paths: Sequence[str], target_size: Tuple[int, int]) -> torch.Tensor:
    """Decode only: uint8 (N, Hs, Ws, 3) when all frames share one source size (the device then does the Pillow-exact
    resize + ToTensor, pi3b_resize_u8_pil); otherwise the float (N, 3, H, W) result of the host path."""
    import numpy as np
    arrs = [decode_image_u8(p) for p in paths]
    if all(a.shape == arrs[0].shape for a in arrs):
        return torch.from_numpy(np.stack(arrs))
    return torch.stack([resize_u8_on_host(a, target_size) for a in arrs])
