"""ORACLE — test infrastructure only (never imported by pi3_slam_b200/).

CPU restatement of the per-chunk post-processing of the reference's `OfflineChunkCreator`
(slam/offline_chunk_creator.py:114-254) and of the chunk-to-chunk Sim3 step
(utils/reconstruction_alignment.py:16-37, 79-105), in plain torch / numpy.

Pinned: `oracle/make_golden_chunk.py` runs the real reference `_process_single_chunk` (reference model on
CPU, synthetic weights) and stores its output as tests/golden/chunk_tiny.pt; tests/test_oracle.py checks
this restatement against that file. The Sim3 arithmetic lives in un-vendored pytheia==0.2.9
(requirements.txt:10), which is absent here: `umeyama_sim3` restates the published closed form
(Umeyama 1991) and is pinned only by known-answer tests — Sim3 parity vs pytheia is UNPINNED.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ---- masks: offline_chunk_creator.py:114-119 + pi3/utils/geometry.py:347-375 ----
def depth_edge(depth: torch.Tensor, rtol: float) -> torch.Tensor:
    shape = depth.shape
    d = depth.reshape(-1, 1, *shape[-2:])
    diff = F.max_pool2d(d, 3, stride=1, padding=1) + F.max_pool2d(-d, 3, stride=1, padding=1)
    edge = (diff / d).nan_to_num_() > rtol
    return edge.reshape(*shape)


def compute_masks(local_points: torch.Tensor, conf: torch.Tensor) -> torch.Tensor:
    """local_points (...,H,W,3), conf (...,H,W,1) -> bool (...,H,W)."""
    masks = torch.sigmoid(conf[..., 0]) > 0.1
    return torch.logical_and(masks, ~depth_edge(local_points[..., 2], rtol=0.03))


# ---- metric scale: offline_chunk_creator.py:121-127 ----
def scale_factor(moge_depth: torch.Tensor, pi3_depth: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    return (moge_depth[mask] / pi3_depth[mask]).median()


# ---- grid keypoints: utils/keypoint_extraction.py:53-171 (global generator of `device`, as the reference's
# GridKeypointExtractor(device=...) draws them; 'cuda' is the reference default) ----
def grid_keypoints(num_frames: int, H: int, W: int, max_kp: int, device="cpu") -> torch.Tensor:
    margin = min(H, W) * 0.05
    eff_h, eff_w = H - 2 * margin, W - 2 * margin
    spacing = int(np.sqrt((eff_h * eff_w) / max_kp))
    spacing = max(8, min(spacing, min(H, W) // 4))
    out = []
    for _ in range(num_frames):
        gx = torch.arange(margin, W - margin, spacing, device=device)
        gy = torch.arange(margin, H - margin, spacing, device=device)
        yy, xx = torch.meshgrid(gy, gx, indexing="ij")
        coords = torch.stack([xx.flatten(), yy.flatten()], dim=-1)
        if len(coords) > max_kp:
            coords = coords[torch.randperm(len(coords), device=device)[:max_kp]]
        out.append(coords)
    return torch.stack(out, 0)


# ---- sampling at keypoints: offline_chunk_creator.py:129-159, keypoint_extraction.py:207-229 ----
def _grid(keypoints: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Normalised grid, evaluated on the keypoints' device and THEN moved to the CPU, as the reference does
    (`grid = ....cpu()`, offline_chunk_creator.py:143; `grid.cpu()`, keypoint_extraction.py:222)."""
    gx = (keypoints[:, :, 0] / (W - 1)) * 2 - 1
    gy = (keypoints[:, :, 1] / (H - 1)) * 2 - 1
    return torch.stack([gx, gy], dim=-1).unsqueeze(1).cpu()


def interpolate_at_keypoints(points, local_points, conf, masks, keypoints, H, W) -> Dict[str, torch.Tensor]:
    grid = _grid(keypoints, H, W)
    kw = dict(align_corners=False, padding_mode="border")
    p = F.grid_sample(points.permute(0, 3, 1, 2), grid, mode="bilinear", **kw)
    l = F.grid_sample(local_points.permute(0, 3, 1, 2), grid, mode="bilinear", **kw)
    c = F.grid_sample(conf.permute(0, 3, 1, 2), grid, mode="nearest", **kw)
    m = F.grid_sample(masks.unsqueeze(1).float(), grid, mode="nearest", **kw)
    return {"points": p.squeeze(2).permute(0, 2, 1), "local_points": l.squeeze(2).permute(0, 2, 1),
            "conf": c.squeeze(1).permute(0, 2, 1), "masks": m.squeeze(1).permute(0, 2, 1).bool()}


def colors_at_keypoints(images: torch.Tensor, keypoints: torch.Tensor) -> torch.Tensor:
    H, W = images.shape[-2:]
    col = F.grid_sample(images, _grid(keypoints, H, W), mode="bilinear", align_corners=False, padding_mode="border")
    return (col.squeeze(2).transpose(1, 2) * 255).to(torch.uint8)


# ---- intrinsics: utils/camera_estimation.py:12-70 + geometry_torch.py:114-169 + geometry_numpy.py:79-96 ----
def _uv(width, height):
    aspect = width / height
    sx = aspect / (1 + aspect ** 2) ** 0.5
    sy = 1 / (1 + aspect ** 2) ** 0.5
    u = torch.linspace(-sx * (width - 1) / width, sx * (width - 1) / width, width, dtype=torch.float32)
    v = torch.linspace(-sy * (height - 1) / height, sy * (height - 1) / height, height, dtype=torch.float32)
    u, v = torch.meshgrid(u, v, indexing="xy")
    return torch.stack([u, v], dim=-1)


def camera_params(local_points: torch.Tensor, conf: torch.Tensor) -> Dict[str, torch.Tensor]:
    """local_points (N,H,W,3), conf (N,H,W,1) CPU -> reference-shaped dict (intrinsics (N,3,3), others (1,N))."""
    from functools import partial
    from scipy.optimize import least_squares
    N, H, W = local_points.shape[:3]
    masks = torch.sigmoid(conf[..., 0]) > 0.1
    uv = _uv(W, H)
    pts_lr = F.interpolate(local_points.permute(0, 3, 1, 2), (64, 64), mode="nearest").permute(0, 2, 3, 1).numpy()
    uv_lr = F.interpolate(uv.unsqueeze(0).permute(0, 3, 1, 2), (64, 64), mode="nearest").squeeze(0).permute(1, 2, 0).numpy()
    m_lr = (F.interpolate(masks.float().unsqueeze(1), (64, 64), mode="nearest").squeeze(1) > 0).numpy()
    focal, shift = [], []
    for i in range(N):
        p, u = pts_lr[i][m_lr[i]], uv_lr[m_lr[i]]
        if u.shape[0] < 2:
            focal.append(1.0); shift.append(0.0)
            continue
        uvf, xy, z = u.reshape(-1, 2), p[..., :2].reshape(-1, 2), p[..., 2].reshape(-1)

        def fn(uvf, xy, z, s):
            xp = xy / (z + s)[:, None]
            f = (xp * uvf).sum() / np.square(xp).sum()
            return (f * xp - uvf).ravel()
        sol = least_squares(partial(fn, uvf, xy, z), x0=0, ftol=1e-3, method="lm")
        s = sol["x"].squeeze().astype(np.float32)
        xp = xy / (z + s)[:, None]
        focal.append(float((xp * uvf).sum() / np.square(xp).sum()))
        shift.append(float(s))
    focal = torch.tensor(focal, dtype=torch.float32).reshape(1, N)
    shift = torch.tensor(shift, dtype=torch.float32).reshape(1, N)
    aspect = W / H
    fx = focal / 2 * (1 + aspect ** 2) ** 0.5 / aspect * W
    fy = focal / 2 * (1 + aspect ** 2) ** 0.5 * H
    cx = torch.full_like(fx, W // 2)
    cy = torch.full_like(fy, H // 2)
    K = torch.zeros(N, 3, 3)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = fx[0], fy[0], cx[0], cy[0], 1.0
    return {"intrinsics": K, "focal": focal, "shift": shift, "fx": fx, "fy": fy, "cx": cx, "cy": cy}


# ---- Sim3 between consecutive chunks ----
def create_view_graph_matches(chunk_size: int, overlap_size: int) -> List[Tuple[int, int]]:
    """(ref_view, qry_view) pairs of the overlap (reconstruction_alignment.py:16-37)."""
    return [(chunk_size - overlap_size + i, i) for i in range(overlap_size)]


def umeyama_sim3(qry: np.ndarray, ref: np.ndarray) -> Tuple[float, np.ndarray, np.ndarray]:
    """Closed-form (s, R, t) minimising sum |ref - (s R qry + t)|^2 (Umeyama 1991), float64."""
    q = np.asarray(qry, dtype=np.float64)
    r = np.asarray(ref, dtype=np.float64)
    mq, mr = q.mean(0), r.mean(0)
    dq, dr = q - mq, r - mr
    S = dr.T @ dq / len(q)
    U, sig, Vt = np.linalg.svd(S)
    D = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2, 2] = -1
    R = U @ D @ Vt
    var_q = (dq ** 2).sum() / len(q)
    s = float(np.trace(np.diag(sig) @ D) / var_q)
    t = mr - s * R @ mq
    return s, R, t


def sim3_matrix(s: float, R: np.ndarray, t: np.ndarray) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = s * R
    T[:3, 3] = t
    return T


def common_points(ref_chunk: Dict, qry_chunk: Dict, chunk_length: int, overlap: int):
    """Correspondences between two consecutive chunks: for every overlap view pair (ref L-ov+i <-> qry i,
    the same source frame), keypoints with identical (x, y) in both chunks' keypoint sets (the reference
    matches tracks by feature, reconstruction_alignment.py:74) that are valid in both masks."""
    pr, pq = [], []
    for ref_v, qry_v in create_view_graph_matches(chunk_length, overlap):
        kr = ref_chunk["keypoints"][ref_v].float().numpy()
        kq = qry_chunk["keypoints"][qry_v].float().numpy()
        index = {(float(x), float(y)): i for i, (x, y) in enumerate(kq)}
        for i, (x, y) in enumerate(kr):
            j = index.get((float(x), float(y)))
            if j is None:
                continue
            if bool(ref_chunk["masks"][ref_v, i, 0]) and bool(qry_chunk["masks"][qry_v, j, 0]):
                pr.append(ref_chunk["points"][ref_v, i].float().numpy())
                pq.append(qry_chunk["points"][qry_v, j].float().numpy())
    return np.asarray(pr, dtype=np.float64).reshape(-1, 3), np.asarray(pq, dtype=np.float64).reshape(-1, 3)


def chunk_pair_alignment(ref_chunk: Dict, qry_chunk: Dict, chunk_length: int, overlap: int) -> np.ndarray:
    """Sim3 (4x4) mapping the query chunk onto the reference chunk: common points, keep those closer to the
    last reference camera than the median distance, closed-form Sim3 (reconstruction_alignment.py:74-105)."""
    pr, pq = common_points(ref_chunk, qry_chunk, chunk_length, overlap)
    last_cam = ref_chunk["camera_poses"][-1, :3, 3].double().numpy()
    dist = np.linalg.norm(pr - last_cam, axis=1)
    keep = dist < np.median(dist)
    return sim3_matrix(*umeyama_sim3(pq[keep], pr[keep]))
