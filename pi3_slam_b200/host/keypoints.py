"""Grid keypoints (reference: utils/keypoint_extraction.py:32-171). Host logic: a lattice plus a
per-frame `torch.randperm` subsample drawn from the *global* generator of `device`, so that with the same
seed, device type and call order the indices are bit-identical to the reference's."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch


class GridKeypointExtractor:
    def __init__(self, max_num_keypoints: int = 512, grid_spacing: int = None, device: str = "cuda"):
        self.device = device
        self.max_num_keypoints = max_num_keypoints
        self.grid_spacing = grid_spacing

    def _calculate_grid_spacing(self, H: int, W: int) -> int:
        if self.grid_spacing is not None:
            return self.grid_spacing
        margin = min(H, W) * 0.05
        eff_h, eff_w = H - 2 * margin, W - 2 * margin
        if eff_h <= 0 or eff_w <= 0:
            return max(H, W)
        spacing = int(np.sqrt((eff_h * eff_w) / self.max_num_keypoints))
        return max(8, min(spacing, min(H, W) // 4))

    def extract(self, num_frames: int, H: int, W: int) -> Dict[str, torch.Tensor]:
        """keypoints (N,K,2) fp32 (x, y); descriptors zeros (N,K,128); scores ones (N,K)."""
        kps = []
        for _ in range(num_frames):
            spacing = self._calculate_grid_spacing(H, W)
            margin = min(H, W) * 0.05
            gx = torch.arange(margin, W - margin, spacing, device=self.device)
            gy = torch.arange(margin, H - margin, spacing, device=self.device)
            if len(gx) == 0 or len(gy) == 0:
                coords = torch.tensor([[W // 2, H // 2]], device=self.device, dtype=torch.float32)
            else:
                yy, xx = torch.meshgrid(gy, gx, indexing="ij")
                coords = torch.stack([xx.flatten(), yy.flatten()], dim=-1)
            if len(coords) > self.max_num_keypoints:
                idx = torch.randperm(len(coords), device=self.device)[: self.max_num_keypoints]
                coords = coords[idx]
            kps.append(coords)
        keypoints = torch.stack(kps, dim=0)
        K = keypoints.shape[1]
        return {
            "keypoints": keypoints,
            "descriptors": torch.zeros(num_frames, K, 128, device=self.device),
            "scores": torch.ones(num_frames, K, device=self.device),
        }


def create_keypoint_extractor(keypoint_type: str = "aliked", **kwargs):
    """Factory with the reference's signature (keypoint_extraction.py:392-420): keyword arguments
    max_num_keypoints, detection_threshold (ignored for 'grid'), device."""
    kt = (keypoint_type or "none").lower()
    if kt == "none":
        return None
    if kt == "grid":
        kwargs.pop("grid_spacing", None)
        kwargs.pop("detection_threshold", None)
        return GridKeypointExtractor(**kwargs)
    if kt == "aliked":
        raise ImportError("ALIKED (lightglue) is a third-party network outside the chunk hot path; use 'grid'")
    raise ValueError(f"Unknown keypoint type: {keypoint_type}. Must be 'aliked', 'grid', or 'none'")
