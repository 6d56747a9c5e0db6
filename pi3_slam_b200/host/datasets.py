"""Chunk bookkeeping (reference: datasets/image_datasets.py:13-90)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
from torch.utils.data import Dataset

from .image_utils import load_images_from_paths


def chunk_index_list(num_items: int, chunk_length: int, overlap: int) -> List[Tuple[int, int]]:
    """[(start, end_exclusive)] with stride chunk_length - overlap; chunks shorter than 2 frames dropped
    (image_datasets.py:41-47)."""
    if chunk_length - overlap <= 0:
        raise ValueError("chunk_length must exceed overlap")
    out = []
    i = 0
    while i < num_items:
        end = min(i + chunk_length, num_items)
        if end - i >= 2:
            out.append((i, end))
        i += chunk_length - overlap
    return out


def shard_chunks(num_chunks: int, rank: int, world_size: int) -> List[int]:
    """Chunks are independent units (offline_chunk_creator.py:296-303): chunk i is owned by rank i % world."""
    return list(range(rank, num_chunks, world_size))


class ChunkImageDataset(Dataset):
    """Yields {'chunk' (N,3,H,W), 'start_idx', 'end_idx', 'chunk_paths'} like the reference dataset.
    `loader(paths, target_size) -> (N,3,H,W)` is injectable (synthetic frames in tests/bench)."""

    def __init__(self, image_paths: Sequence, chunk_length: int, overlap: int, target_size: Tuple[int, int],
                 device: str = "cpu", undistortion_maps=None,
                 loader: Optional[Callable[[Sequence, Tuple[int, int]], torch.Tensor]] = None):
        if undistortion_maps is not None:
            raise NotImplementedError("undistortion (cv2 / pytheia calibration) is outside the chunk hot path")
        self.image_paths = list(image_paths)
        self.chunk_length = chunk_length
        self.overlap = overlap
        self.target_size = target_size
        self.device = device
        self.chunk_indices = chunk_index_list(len(self.image_paths), chunk_length, overlap)
        self._loader = loader or load_images_from_paths

    def __len__(self) -> int:
        return len(self.chunk_indices)

    def __getitem__(self, idx: int) -> Dict:
        start, end = self.chunk_indices[idx]
        paths = self.image_paths[start:end]
        chunk = self._loader(paths, self.target_size)
        return {
            "chunk": chunk.to(self.device),
            "start_idx": torch.tensor([start]),
            "end_idx": torch.tensor([end]),
            "chunk_paths": [paths],
        }
