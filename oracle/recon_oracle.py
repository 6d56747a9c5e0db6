"""ORACLE — test infrastructure only (never imported by pi3_slam_b200/).

numpy restatement of the track / observation construction of the reference's stage-2 front end
(`ChunkPTRecon.create_recon_from_chunk`, utils/chunk_reconstruction.py:128-185, with `_project_points_to_other_cams`
:445-479 and `_project_3d_points_to_camera` :481-509): what the Python loops hand to pytheia, as flat arrays in the order
of the reference's AddTrack / AddObservation calls. pytheia itself (bundle adjustment) is absent and out of scope.

Pinned: oracle/make_golden_recon.py imports the real `utils.chunk_reconstruction` with a recording stand-in for the
`pytheia` module, runs `create_recon_from_chunk` on the committed chunk golden and stores the recorded calls as
tests/golden/recon_tiny.pt; tests/test_oracle.py checks this restatement against that file (order and membership exact,
coordinates bit-equal)."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch


def target_frames(frame_idx: int, num_frames: int, max_observations_per_track: int) -> List[int]:
    """Frames a source frame's keypoints are projected into (:162-167): ALL earlier frames, then the next
    max_observations_per_track // 2 frames."""
    return list(range(frame_idx)) + list(range(frame_idx + 1, num_frames))[: max_observations_per_track // 2]


def project(points_3d: torch.Tensor, target_pose: np.ndarray, target_intrinsics) -> np.ndarray:
    """:494-509 (the source pose is unused by the reference: `points` are already world coordinates)."""
    world_to_target = np.linalg.inv(target_pose)
    pts = (world_to_target @ np.hstack([points_3d.cpu().numpy(), np.ones((points_3d.shape[0], 1))]).T).T
    p2 = pts[:, :3] / pts[:, 2:3]
    if target_intrinsics is not None:
        return (target_intrinsics @ p2.T).T
    return p2[:, :2]


def observation_list(chunk: Dict, max_observations_per_track: int = 5) -> Dict[str, np.ndarray]:
    """Tracks: one per (frame, keypoint) in frame-major order (point homogeneous float64, colour). Observations in
    AddObservation order: per source frame its own K keypoints, then per target frame the in-bounds reprojections."""
    N, K = chunk["keypoints"].shape[:2]
    W, H = chunk["original_width"], chunk["original_height"]
    has_intr = "intrinsics" in chunk
    trk_pt, trk_col, o_view, o_track, o_xy = [], [], [], [], []
    for f in range(N):
        p3 = chunk["points"][f].cpu().numpy()
        col = chunk["colors"][f].cpu().numpy()
        kp = chunk["keypoints"][f].cpu().numpy()
        base = f * K
        for k in range(K):
            trk_pt.append(np.hstack([p3[k], 1]))
            trk_col.append(col[k])
            o_view.append(f); o_track.append(base + k); o_xy.append(kp[k].astype(np.float64))
        for g in target_frames(f, N, max_observations_per_track):
            intr = chunk["intrinsics"][g].cpu().numpy() if has_intr else None
            pr = project(chunk["points"][f], chunk["camera_poses"][g].cpu().numpy(), intr)[:, :2]
            for k in range(K):
                if 0 <= pr[k, 0] < W and 0 <= pr[k, 1] < H:
                    o_view.append(g); o_track.append(base + k); o_xy.append(pr[k].astype(np.float64))
    return {"track_point": np.asarray(trk_pt, np.float64).reshape(-1, 4), "track_color": np.asarray(trk_col).reshape(-1, 3),
            "obs_view": np.asarray(o_view, np.int64), "obs_track": np.asarray(o_track, np.int64),
            "obs_xy": np.asarray(o_xy, np.float64).reshape(-1, 2)}
