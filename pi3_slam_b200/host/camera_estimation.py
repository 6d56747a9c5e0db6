"""Per-frame focal / z-shift recovery and intrinsics (reference: utils/camera_estimation.py:12-132,
utils/geometry_torch.py:39-51,114-169, utils/geometry_numpy.py:79-96).

The device gathers the 64x64 nearest-downsampled samples (N x 4096 x 3 floats instead of the dense maps) and, by
default, also runs the per-frame 1-parameter Levenberg-Marquardt solve (`pi3b_focal_shift_lm`, SURVEY.md section 8 row
f3: the algorithm scipy executes, restated — 67 ms of serial host work per 100-frame chunk removed). The host scipy
solve of the reference is kept as `solver="host"` (what the device solver is tested against).
"""
from __future__ import annotations

from functools import partial
from typing import Dict, Tuple

import numpy as np
import torch

from .. import ops


def nearest_indices(in_size: int, out_size: int) -> np.ndarray:
    """Source indices of F.interpolate(mode='nearest') (ATen: floor(dst * float(in)/out), clamped)."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1).astype(np.int32)


def normalized_view_plane_uv(width: int, height: int) -> np.ndarray:
    """(H, W, 2) fp32 UV grid, corners at +-(w, h)/diagonal (geometry_torch.py:39-51)."""
    aspect = width / height
    span_x = aspect / (1 + aspect ** 2) ** 0.5
    span_y = 1 / (1 + aspect ** 2) ** 0.5
    u = torch.linspace(-span_x * (width - 1) / width, span_x * (width - 1) / width, width, dtype=torch.float32)
    v = torch.linspace(-span_y * (height - 1) / height, span_y * (height - 1) / height, height, dtype=torch.float32)
    uu, vv = torch.meshgrid(u, v, indexing="xy")
    return torch.stack([uu, vv], dim=-1).numpy()


def solve_optimal_focal_shift(uv: np.ndarray, xyz: np.ndarray) -> Tuple[np.float32, float]:
    """min_{shift, f} | f * xy / (z + shift) - uv |, LM over shift with f eliminated (geometry_numpy.py:79-96)."""
    from scipy.optimize import least_squares
    uv, xy, z = uv.reshape(-1, 2), xyz[..., :2].reshape(-1, 2), xyz[..., 2].reshape(-1)

    def fn(uv, xy, z, shift):
        xy_proj = xy / (z + shift)[:, None]
        f = (xy_proj * uv).sum() / np.square(xy_proj).sum()
        return (f * xy_proj - uv).ravel()

    sol = least_squares(partial(fn, uv, xy, z), x0=0, ftol=1e-3, method="lm")
    shift = sol["x"].squeeze().astype(np.float32)
    xy_proj = xy / (z + shift)[:, None]
    focal = (xy_proj * uv).sum() / np.square(xy_proj).sum()
    return shift, focal


def recover_focal_shift_from_samples(pts: np.ndarray, valid: np.ndarray, uv_lr: np.ndarray):
    """pts (N, S, 3), valid (N, S) bool, uv_lr (S, 2) -> focal (N,), shift (N,) fp32."""
    focal, shift = [], []
    for i in range(pts.shape[0]):
        m = valid[i]
        if int(m.sum()) < 2:
            focal.append(1.0)
            shift.append(0.0)
            continue
        s, f = solve_optimal_focal_shift(uv_lr[m], pts[i][m])
        focal.append(float(f))
        shift.append(float(s))
    return np.asarray(focal, dtype=np.float32), np.asarray(shift, dtype=np.float32)


_SAMPLE_CACHE: Dict[tuple, tuple] = {}


def sample_for_intrinsics(local_points: torch.Tensor, conf: torch.Tensor,
                          downsample_size: Tuple[int, int] = (64, 64)):
    """Device half of the intrinsics estimate: the 64x64 nearest samples (N, S, 3) and their validity
    (sigmoid(conf) > 0.1) (N, S) uint8, both still on the device, plus the matching uv samples (S, 2) numpy."""
    N, H, W = local_points.shape[:3]
    dev = local_points.device
    key = (H, W, tuple(downsample_size), str(dev))
    c = _SAMPLE_CACHE.get(key)
    if c is None:
        # per-shape constants, uploaded once: a pageable host->device copy makes the driver wait for all the work
        # already queued on the stream, which would serialise the host behind the chunk's forward
        ys = nearest_indices(H, downsample_size[0])
        xs = nearest_indices(W, downsample_size[1])
        c = (torch.from_numpy(ys).to(dev), torch.from_numpy(xs).to(dev),
             normalized_view_plane_uv(W, H)[ys][:, xs].reshape(-1, 2))
        _SAMPLE_CACHE[key] = c
    pts, val = ops.sample_grid(local_points, conf, c[0], c[1], 0.1)
    return pts, val, c[2]


def solve_focal_shift_on_device(pts: torch.Tensor, valid: torch.Tensor, uv_lr: np.ndarray) -> torch.Tensor:
    """Device half of the solve: fp32 [N, 4] = (shift, focal, nfev, #valid) per frame, still on the device."""
    key = ("uv", id(uv_lr), str(pts.device))
    uv_dev = _SAMPLE_CACHE.get(key)
    if uv_dev is None:
        uv_dev = torch.from_numpy(np.ascontiguousarray(uv_lr, dtype=np.float32)).to(pts.device)
        _SAMPLE_CACHE[key] = uv_dev
    return ops.focal_shift_lm(pts, valid, uv_dev, 1e-3)


def camera_parameters_from_focal_shift(focal_np: np.ndarray, shift_np: np.ndarray, H: int, W: int,
                                       device="cpu") -> Dict[str, torch.Tensor]:
    """The reference's focal -> (fx, fy, cx, cy, K) formulas (camera_estimation.py:47-60). Returns the reference-shaped
    dict: intrinsics (N,3,3); the rest (1,N)."""
    N = focal_np.shape[0]
    focal = torch.from_numpy(np.ascontiguousarray(focal_np, dtype=np.float32)).to(device).reshape(1, N)
    shift = torch.from_numpy(np.ascontiguousarray(shift_np, dtype=np.float32)).to(device).reshape(1, N)
    return _intrinsics_dict(focal, shift, H, W, device)


def camera_parameters_from_samples(pts: np.ndarray, valid: np.ndarray, uv_lr: np.ndarray, H: int, W: int,
                                   device="cpu") -> Dict[str, torch.Tensor]:
    """Host solve (`solver="host"`): per-frame scipy LM, then the intrinsics formulas."""
    N = pts.shape[0]
    focal_np, shift_np = recover_focal_shift_from_samples(pts, valid.astype(bool), uv_lr)
    focal = torch.from_numpy(focal_np).to(device).reshape(1, N)
    shift = torch.from_numpy(shift_np).to(device).reshape(1, N)
    return _intrinsics_dict(focal, shift, H, W, device)


def _intrinsics_dict(focal: torch.Tensor, shift: torch.Tensor, H: int, W: int, device) -> Dict[str, torch.Tensor]:
    N = focal.shape[1]
    aspect = W / H
    fx = focal / 2 * (1 + aspect ** 2) ** 0.5 / aspect * W
    fy = focal / 2 * (1 + aspect ** 2) ** 0.5 * H
    cx = torch.full_like(fx, W // 2)
    cy = torch.full_like(fy, H // 2)
    K = torch.zeros(1, N, 3, 3, device=device, dtype=fx.dtype)
    K[..., 0, 0] = fx
    K[..., 1, 1] = fy
    K[..., 0, 2] = cx
    K[..., 1, 2] = cy
    K[..., 2, 2] = 1.0
    # the reference squeezes the batch dim only from tensors with >= 3 dims (camera_estimation.py:127-130)
    return {"intrinsics": K.squeeze(0), "focal": focal, "shift": shift, "fx": fx, "fy": fy, "cx": cx, "cy": cy}


def estimate_camera_parameters_single_chunk(local_points: torch.Tensor, conf: torch.Tensor,
                                            downsample_size: Tuple[int, int] = (64, 64),
                                            solver: str = "device") -> Dict[str, torch.Tensor]:
    """local_points (N,H,W,3), conf (N,H,W,1) on the device -> dict like the reference's
    estimate_camera_parameters_single_chunk: intrinsics (N,3,3); focal/shift/fx/fy/cx/cy (1,N), on the device."""
    H, W = local_points.shape[1:3]
    pts, val, uv_lr = sample_for_intrinsics(local_points, conf, downsample_size)
    if solver == "host":
        return camera_parameters_from_samples(pts.cpu().numpy(), val.cpu().numpy(), uv_lr, H, W, device=local_points.device)
    sol = solve_focal_shift_on_device(pts, val, uv_lr).cpu().numpy()
    return camera_parameters_from_focal_shift(sol[:, 1], sol[:, 0], H, W, device=local_points.device)
