"""Chunk-to-chunk Sim3 (reference: utils/reconstruction_alignment.py:16-37, 74-105).

Host side: which overlap views correspond, which keypoints are common to both chunks (matched by their
lattice coordinate, as pytheia's FindCommonTracksByFeatureInReconstructions does by feature), and the
"closer than the median distance to the last reference camera" filter. Device side: one batched closed-form
Umeyama solve for all consecutive pairs (pairs are independent given the raw chunk outputs), then a prefix
product to express every chunk in the first chunk's frame. Bundle adjustment (pytheia/Ceres) is out of scope.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import ops


def create_view_graph_matches(chunk_size: int, overlap_size: int) -> List[Tuple[int, int]]:
    """(ref_view_id, qry_view_id) pairs of the overlap region (reconstruction_alignment.py:16-37)."""
    return [(chunk_size - overlap_size + i, i) for i in range(overlap_size)]


def common_points(ref_chunk: Dict, qry_chunk: Dict, chunk_length: int, overlap: int) -> Tuple[np.ndarray, np.ndarray]:
    """World points of keypoints present (same x, y) and valid in both chunks over the overlap views."""
    pr, pq = [], []
    n_ref = int(ref_chunk["keypoints"].shape[0])
    for ref_v, qry_v in create_view_graph_matches(min(chunk_length, n_ref), overlap):
        if ref_v < 0 or ref_v >= n_ref or qry_v >= int(qry_chunk["keypoints"].shape[0]):
            continue
        kr = ref_chunk["keypoints"][ref_v].float().numpy()
        kq = qry_chunk["keypoints"][qry_v].float().numpy()
        index = {(float(x), float(y)): j for j, (x, y) in enumerate(kq)}
        mr = ref_chunk["masks"][ref_v, :, 0].numpy()
        mq = qry_chunk["masks"][qry_v, :, 0].numpy()
        for i, (x, y) in enumerate(kr):
            j = index.get((float(x), float(y)))
            if j is not None and mr[i] and mq[j]:
                pr.append(ref_chunk["points"][ref_v, i].float().numpy())
                pq.append(qry_chunk["points"][qry_v, j].float().numpy())
    return (np.asarray(pr, dtype=np.float32).reshape(-1, 3), np.asarray(pq, dtype=np.float32).reshape(-1, 3))


def align_consecutive_chunks(chunks: Sequence[Dict], chunk_length: int, overlap: int,
                             device: str = "cuda") -> Dict[str, torch.Tensor]:
    """Sim3 for every consecutive pair (qry = chunk k+1 onto ref = chunk k) in ONE kernel launch, plus the
    cumulative transforms into chunk 0's frame. Returns {'pairwise' (P,4,4), 'scale' (P,), 'count' (P,),
    'cumulative' (P+1,4,4)} on the CPU in float64/float32."""
    pairs = []
    for k in range(len(chunks) - 1):
        pr, pq = common_points(chunks[k], chunks[k + 1], chunk_length, overlap)
        if len(pr):
            last_cam = chunks[k]["camera_poses"][-1, :3, 3].float().numpy()
            dist = np.linalg.norm(pr - last_cam, axis=1)
            keep = dist < np.median(dist)
            pr, pq = pr[keep], pq[keep]
        pairs.append((pr, pq))
    P = len(pairs)
    if P == 0:
        return {"pairwise": torch.zeros(0, 4, 4), "scale": torch.zeros(0), "count": torch.zeros(0),
                "cumulative": torch.eye(4, dtype=torch.float64)[None]}
    n_max = max(1, max(len(p[0]) for p in pairs))
    qry = torch.zeros(P, n_max, 3)
    ref = torch.zeros(P, n_max, 3)
    valid = torch.zeros(P, n_max, dtype=torch.uint8)
    for i, (pr, pq) in enumerate(pairs):
        n = len(pr)
        if n:
            qry[i, :n] = torch.from_numpy(pq)
            ref[i, :n] = torch.from_numpy(pr)
            valid[i, :n] = 1
    T, scale, count = ops.umeyama_sim3(qry.to(device), ref.to(device), valid.to(device))
    T, scale, count = T.cpu(), scale.cpu(), count.cpu()
    cum = [torch.eye(4, dtype=torch.float64)]
    for i in range(P):
        cum.append(cum[-1] @ T[i].double())
    return {"pairwise": T, "scale": scale, "count": count, "cumulative": torch.stack(cum)}
