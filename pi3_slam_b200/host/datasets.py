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


def split_frames(num_frames: int, group_size: int) -> List[Tuple[int, int]]:
    """(first_frame, count) of every rank of a sequence-parallel group: contiguous, as even as possible, the first
    num_frames % group_size ranks hold one frame more."""
    base, extra = divmod(num_frames, group_size)
    out, start = [], 0
    for i in range(group_size):
        n = base + (1 if i < extra else 0)
        out.append((start, n))
        start += n
    return out


def plan_waves(frames_per_chunk: Sequence[int], world_size: int, allow_sp: bool = True) -> List[List[Tuple[int, Tuple[int, ...]]]]:
    """Hybrid schedule for a sequence of independent chunks on `world_size` GPUs.

    Returns waves; each wave is a list of (chunk_index, ranks): while at least `world_size` chunks remain every rank
    takes one whole chunk (no communication). When fewer remain, the ranks are split into k equal groups (k = the
    largest divisor of world_size not above the number of remaining chunks) and each group runs ONE chunk
    sequence-parallel (frames split over the group, K/V all-gathered in the chunk-global layers), so the last wave
    does not leave GPUs idle: 13 chunks on 8 GPUs = 8 singles, then 4 chunks on 4 pairs, then the 40-frame tail on all
    8 (makespan 1.53 chunk-times instead of 2). Frames are split as evenly as possible (`split_frames`: the first
    n % g ranks of a group hold one frame more); a group is only shrunk when the chunk has fewer frames than ranks.
    With allow_sp=False this is plain round-robin sharding (chunk i on rank i % world_size), identical to
    `shard_chunks`."""
    n = len(frames_per_chunk)
    waves: List[List[Tuple[int, Tuple[int, ...]]]] = []
    nxt = 0
    while nxt < n:
        remaining = n - nxt
        if remaining >= world_size or not allow_sp or world_size == 1:
            take = min(world_size, remaining)
            waves.append([(nxt + j, ((nxt + j) % world_size,) if not allow_sp else (j,)) for j in range(take)])
            nxt += take
            continue
        k = max(d for d in range(1, remaining + 1) if world_size % d == 0)
        g = world_size // k
        wave = []
        for j in range(k):
            gj = max(1, min(g, frames_per_chunk[nxt + j]))
            wave.append((nxt + j, tuple(range(j * g, j * g + gj))))
        waves.append(wave)
        nxt += k
    return waves


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

    def load_frames(self, start: int, end: int) -> torch.Tensor:
        """Frames [start, end) of the sequence at the target size, (n,3,H,W) — a slice of a chunk for the
        sequence-parallel ranks, which only need their own frames."""
        return self._loader(self.image_paths[start:end], self.target_size)

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


def nccl_high_priority_options():
    """ProcessGroupNCCL options that put NCCL's kernels on a HIGH-PRIORITY stream, or None when unavailable. The
    sequence-parallel K/V all-gather must start while the attention over the local block is running: that kernel
    occupies every SM (one CTA of 512 threads x 128 registers each) with thousands of CTAs queued behind, and with equal
    priorities the all-gather kernel is only scheduled once that grid has drained (measured: the whole all-gather
    exposed, profiles/r02_bench_cfg2_n2_first.json); with a higher priority its few CTAs take the first SMs that free up."""
    try:
        import torch.distributed as dist
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        return opts
    except Exception:
        return None


def group_all_ok(ok: bool, group=None, device="cpu") -> bool:
    """True iff every rank of `group` reports `ok` (one MIN all-reduce of a flag). A chunk that is processed by several
    ranks (sequence-parallel waves) is entered by all of them or by none: a rank whose frames failed to load must not leave
    the others waiting in the K/V all-gather — the group skips the chunk together, the reference's degrade-and-continue
    contract (`offline_chunk_creator.py:299-303`: a failed chunk is reported and skipped)."""
    import torch
    import torch.distributed as dist
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())
