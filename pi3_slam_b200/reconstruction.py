"""Stage-2 front end on the device: the track / observation list that
`ChunkPTRecon.create_recon_from_chunk` (reference utils/chunk_reconstruction.py:128-185) builds with per-keypoint Python
loops — one track per (frame, keypoint) with its own 2-D observation, plus the reprojection of that 3-D point into all
earlier frames and the next `max_observations_per_track // 2` frames, kept when it lands inside the image — as flat arrays
in the reference's AddTrack / AddObservation order. The O(N^2 K) projection + bounds test + ordered compaction is one
C-ABI call (`pi3b_project_observations`); pytheia (bundle adjustment, out of scope and absent here) would only receive
the list."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import _lib


def target_frames(frame_idx: int, num_frames: int, max_observations_per_track: int) -> List[int]:
    """Frames a source frame's keypoints are projected into (chunk_reconstruction.py:162-167)."""
    return list(range(frame_idx)) + list(range(frame_idx + 1, num_frames))[: max_observations_per_track // 2]


def chunk_observations(chunk: Dict, max_observations_per_track: int = 5, device="cuda") -> Dict[str, np.ndarray]:
    """chunk: a loaded chunk_*.pt dict (keypoints, colors, points (N,K,3), camera_poses (N,4,4), optional intrinsics
    (N,3,3), original_width / original_height). Returns numpy arrays track_point (N*K,4) float64, track_color (N*K,3),
    obs_view, obs_track (int64) and obs_xy (float64) in the order of the reference's AddObservation calls."""
    lib = _lib.load()
    N, K = chunk["keypoints"].shape[:2]
    W, H = float(chunk["original_width"]), float(chunk["original_height"])
    dev = torch.device(device)
    pairs = [(f, g) for f in range(N) for g in target_frames(f, N, max_observations_per_track)]
    pts16 = chunk["points"]
    track_point = np.concatenate([pts16.reshape(-1, 3).cpu().numpy().astype(np.float64), np.ones((N * K, 1))], axis=1)
    track_color = chunk["colors"].reshape(-1, 3).cpu().numpy()
    own_xy = chunk["keypoints"].cpu().numpy().astype(np.float64)            # (N, K, 2)
    if not pairs:
        return {"track_point": track_point, "track_color": track_color, "obs_view": np.repeat(np.arange(N), K),
                "obs_track": np.arange(N * K), "obs_xy": own_xy.reshape(-1, 2)}
    R = len(pairs)
    # inverse poses exactly as the reference forms them (numpy.linalg.inv of the fp32 pose, :495), then promoted to fp64
    poses = chunk["camera_poses"].cpu().numpy()
    w2t = np.stack([np.linalg.inv(poses[i]) for i in range(N)]).astype(np.float64).reshape(N, 16)
    d_pts = pts16.to(dev, torch.float32).contiguous()
    d_w2t = torch.from_numpy(w2t).to(dev)
    d_k = None
    if "intrinsics" in chunk:
        d_k = torch.from_numpy(chunk["intrinsics"].cpu().numpy().astype(np.float64).reshape(N, 9)).to(dev)
    d_pairs = torch.tensor(pairs, dtype=torch.int32, device=dev)
    cap = R * K
    row_uv = torch.empty(R, K, 2, device=dev, dtype=torch.float64)
    row_kp = torch.empty(R, K, device=dev, dtype=torch.int32)
    row_count = torch.empty(R, device=dev, dtype=torch.int32)
    row_offset = torch.empty(R + 1, device=dev, dtype=torch.int32)
    o_src = torch.empty(cap, device=dev, dtype=torch.int32)
    o_kp = torch.empty(cap, device=dev, dtype=torch.int32)
    o_view = torch.empty(cap, device=dev, dtype=torch.int32)
    o_xy = torch.empty(cap, 2, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        rc = lib.pi3b_project_observations(torch.cuda.current_stream().cuda_stream, d_pts.data_ptr(), d_w2t.data_ptr(),
                                           None if d_k is None else d_k.data_ptr(), d_pairs.data_ptr(), R, K, W, H,
                                           row_uv.data_ptr(), row_kp.data_ptr(), row_count.data_ptr(), row_offset.data_ptr(),
                                           o_src.data_ptr(), o_kp.data_ptr(), o_view.data_ptr(), o_xy.data_ptr(), cap)
    _lib.check(rc, "pi3b_project_observations")
    off = row_offset.cpu().numpy()
    total = int(off[-1])
    src = o_src[:total].cpu().numpy()
    kp = o_kp[:total].cpu().numpy().astype(np.int64)
    view = o_view[:total].cpu().numpy().astype(np.int64)
    xy = o_xy[:total].cpu().numpy()
    # interleave in the reference's order: per source frame its own K observations, then its reprojections
    first_row = np.searchsorted(np.asarray([p[0] for p in pairs]), np.arange(N + 1))     # rows of frame f: [first_row[f], first_row[f+1])
    seg_view, seg_track, seg_xy = [], [], []
    for f in range(N):
        seg_view.append(np.full(K, f, np.int64)); seg_track.append(f * K + np.arange(K)); seg_xy.append(own_xy[f])
        lo, hi = int(off[first_row[f]]), int(off[first_row[f + 1]])
        seg_view.append(view[lo:hi]); seg_track.append(f * K + kp[lo:hi]); seg_xy.append(xy[lo:hi])
    assert np.all(src == np.repeat(np.arange(N), np.diff(off[first_row])))
    return {"track_point": track_point, "track_color": track_color, "obs_view": np.concatenate(seg_view),
            "obs_track": np.concatenate(seg_track), "obs_xy": np.concatenate(seg_xy)}
