"""`OfflineChunkCreator` — drop-in for reference `slam/offline_chunk_creator.py` on B200.

Same config dataclass, same public methods (`process_and_save`, `_process_single_chunk`), same files:
`<out>/chunks/chunk_%06d.pt`, `<out>/chunks_manifest.json`, `<out>/chunk_metadata.json`
(reference :312-368). What changed underneath:
  * the pi3 forward is the CUDA engine (`pi3_slam_b200.Pi3`), no torch.compile / autocast;
  * masks, metric-scale median, keypoint interpolation and colour sampling run on the device, so only the
    per-keypoint results (~1 MB) cross PCIe instead of the dense maps (350-550 MB per chunk, reference :204-213);
  * chunks are independent units (reference :296-303), so under torch.distributed rank r owns chunks
    r, r+world, ...; every rank writes its own chunk files and rank 0 writes the merged manifest.
Keypoint RNG: the reference draws every frame's `randperm` from one global stream; to make the output
independent of how chunks are sharded, the stream is re-seeded per chunk from (keypoint_seed, chunk_idx)
when `keypoint_seed` is set (default None = reference behaviour: consume the global stream in order).
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from .host.camera_estimation import (camera_parameters_from_focal_shift, camera_parameters_from_samples,
                                     sample_for_intrinsics, solve_focal_shift_on_device)
from .host.datasets import ChunkImageDataset, group_all_ok, nccl_high_priority_options, plan_waves, split_frames
from .host.image_utils import calculate_target_size, load_images_u8, resize_u8_on_host
from .host.keypoints import create_keypoint_extractor
from .pi3 import Pi3


@dataclass
class OfflineCreatorConfig:
    model_path: str
    output_dir: str
    chunk_length: int = 100
    overlap: int = 10
    device: str = "cuda"
    do_metric_depth: bool = True
    keypoint_type: str = "aliked"  # 'aliked' | 'grid' | 'none'
    max_num_keypoints: int = 512
    keypoint_detection_threshold: float = 0.005
    estimate_camera_params: bool = True
    num_loader_workers: int = 2
    pin_memory: bool = True
    cam_dist_path: Optional[str] = None
    # --- additions (defaults keep the reference behaviour) ---
    keypoint_seed: Optional[int] = None     # per-chunk re-seeding => sharding-independent keypoints
    moge_model_path: Optional[str] = None   # local MoGe checkpoint (model.pt); None => hub id like the reference
    sequence_parallel_tail: bool = True     # multi-GPU: run the last (partial) wave of chunks sequence-parallel
    device_resize: bool = True              # loader threads only decode; Pillow-exact resize + ToTensor run on the GPU
    intrinsics_solver: str = "device"       # 'device': pi3b_focal_shift_lm (scipy's 1-parameter LM restated); 'host': scipy


class _PendingChunk:
    """A chunk whose device work is enqueued: `finish()` waits for its event, runs the host half (per-frame scipy LM
    for the intrinsics, dict assembly in the reference's key order / dtypes) and returns the chunk dict."""

    def __init__(self, creator: "OfflineChunkCreator", chunk_paths: List, metrics: Dict, fwd_events=None):
        self.creator, self.chunk_paths, self.metrics, self.fwd_events = creator, chunk_paths, metrics, fwd_events
        self.staged: Dict[str, torch.Tensor] = {}      # pinned host tensors, valid after `event`
        self.host: Dict[str, torch.Tensor] = {}        # already-final host tensors
        self.intr = None
        self.event: Optional[torch.cuda.Event] = None

    def finish(self) -> Dict:
        self.event.synchronize()
        cr = self.creator
        if self.fwd_events is not None:                # deferred timing print of the pipelined path
            t0, ev0, ev1, num_frames = self.fwd_events
            dev_s = float(ev0.elapsed_time(ev1) * 1e-3)
            cr.last_device_seconds = dev_s
            self.metrics.update(infer_s=dev_s, fps=num_frames / max(dev_s, 1e-6))
            print(f"   ⏱️ Inference: {dev_s:.3f}s for {num_frames} frames  ->  {self.metrics['fps']:.2f} FPS")
        camera_params = None
        if self.intr is not None:
            try:
                if self.intr[0] == "device":      # (shift, focal, nfev, #valid) per frame, solved on the GPU
                    _, sol, H, W = self.intr
                    sol = sol.numpy()
                    camera_params = camera_parameters_from_focal_shift(sol[:, 1], sol[:, 0], H, W)
                else:
                    _, pts, val, uv_lr, H, W = self.intr
                    camera_params = camera_parameters_from_samples(pts.numpy(), val.numpy(), uv_lr, H, W)
            except Exception as e:
                print(f"⚠️  Camera parameter estimation failed: {e}")
        result: Dict = {"camera_poses": self.staged["camera_poses"], "image_paths": self.chunk_paths,
                        "_metrics": self.metrics}
        if camera_params is not None:
            result["camera_params"] = camera_params
        if cr.target_size is not None:
            result["original_width"] = cr.target_size[1]
            result["original_height"] = cr.target_size[0]
        for k in ("points", "local_points", "conf", "masks", "keypoints", "descriptors", "scores", "colors"):
            v = self.staged.get(k, self.host.get(k))
            if v is not None:
                result[k] = v.to(torch.float16) if k == "colors" else v
        if result.get("camera_params") is not None:
            result["intrinsics"] = result["camera_params"].get("intrinsics", None)
        return result


class OfflineChunkCreator:
    """Create per-chunk pi3 results (metric scale, keypoints) and save them to disk."""

    def __init__(self, config: OfflineCreatorConfig, model: Optional[Pi3] = None, moge_model=None):
        self.config = config
        if not str(config.device).startswith("cuda"):
            raise _lib.Pi3bError("pi3_slam_b200 runs the chunk path on CUDA (sm_100a) only; device must be 'cuda[:i]'")
        os.makedirs(self.config.output_dir, exist_ok=True)
        self.chunks_dir = os.path.join(self.config.output_dir, "chunks")
        os.makedirs(self.chunks_dir, exist_ok=True)

        self.model: Pi3 = (model if model is not None else Pi3.from_pretrained(config.model_path))
        self.model = self.model.to(config.device).eval()

        # MoGe metric depth (pi3_slam_b200.moge.MoGeModel, the sm_100a implementation of moge/model/v2.py; any object
        # with .infer(image (3,H,W)) -> {"depth": (H,W)} is accepted). The reference loads
        # "Ruicheng/moge-2-vits-normal" from the hub and silently continues without it on failure (:70-79).
        self.moge_model = moge_model
        if self.moge_model is None and config.do_metric_depth:
            try:
                from .moge import MoGeModel
                src = config.moge_model_path or "Ruicheng/moge-2-vits-normal"
                self.moge_model = MoGeModel.from_pretrained(src).to(config.device).eval()
                print("   MoGe loaded for metric scaling")
            except Exception as e:  # same degrade-and-continue contract as the reference
                print(f"⚠️  Failed to initialize MoGe: {e}. Continuing without metric depth.")
                self.moge_model = None

        self.keypoint_extractor = None
        if config.keypoint_type and config.keypoint_type.lower() != "none":
            try:
                self.keypoint_extractor = create_keypoint_extractor(
                    keypoint_type=config.keypoint_type, max_num_keypoints=config.max_num_keypoints,
                    detection_threshold=config.keypoint_detection_threshold, device=config.device)
                print(f"   Keypoint extractor: {config.keypoint_type}")
            except Exception as e:
                print(f"⚠️  Failed to initialize keypoint extractor: {e}. Continuing without keypoints.")
                self.keypoint_extractor = None

        self.target_size: Optional[Tuple[int, int]] = None
        self.undistortion_maps = None
        self._sp_groups: Dict[Tuple[int, ...], object] = {}     # rank tuple -> process group of a sequence-parallel wave
        self.last_run_stats: Dict = {}
        self._resize_tables: Dict[tuple, tuple] = {}            # (Hs, Ws, H, W) -> Pillow coefficient tables on the device
        if getattr(config, "cam_dist_path", None):
            print("⚠️  cam_dist_path given: undistortion maps (cv2/pytheia) are outside this build; ignoring")

    # ------------------------------------------------------------------ device-side steps
    @staticmethod
    def _compute_masks(pi3_result: Dict[str, torch.Tensor]) -> torch.Tensor:
        """sigmoid(conf) > 0.1 and not depth_edge(z, rtol=0.03) -> (B,N,H,W) bool (reference :114-119)."""
        lp = pi3_result["local_points"]
        B, N, H, W = lp.shape[:4]
        mask = torch.empty(B * N, H, W, device=lp.device, dtype=torch.uint8)
        ops.mask_depth_edge(lp.reshape(B * N, H, W, 3), pi3_result["conf"].reshape(B * N, H, W, 1), mask, 0.1, 0.03)
        return mask.view(B, N, H, W).bool()

    @staticmethod
    def _get_scale_factor_for_pi3(moge_metric_depth: torch.Tensor, pi3_metric_depth: torch.Tensor,
                                  mask: torch.Tensor) -> torch.Tensor:
        """median(moge_z[mask] / pi3_z[mask]) (torch lower median; reference :121-127), on the device."""
        a = moge_metric_depth.to(torch.float32).reshape(-1)
        b = pi3_metric_depth.reshape(-1)
        m = mask.reshape(-1).to(torch.uint8).contiguous()
        return ops.median_ratio(a, b, m)[0]

    @staticmethod
    def _interpolate_world_points_for_keypoints(dense: Dict, keypoints: torch.Tensor,
                                                grid_on_cuda: bool = True) -> Dict[str, torch.Tensor]:
        """Sample points / local_points (bilinear) and conf / masks (nearest) at keypoints (reference :129-159).
        `grid_on_cuda`: where the reference would have evaluated `(kp / (W-1)) * 2 - 1` — on the device its keypoints
        live on (torch-CUDA and torch-CPU round the division differently; the kernel reproduces either bit for bit)."""
        images = dense.get("images")
        p, l, c, m, col = ops.keypoint_gather(dense["points"], dense["local_points"], dense["conf"],
                                              dense["masks"].to(torch.uint8).contiguous(), images, keypoints,
                                              grid_on_cuda=grid_on_cuda)
        return {"points": p, "local_points": l, "conf": c, "masks": m, "colors": col}

    def _frames_to_device(self, frames_host: torch.Tensor) -> torch.Tensor:
        """Host frames -> (1, N, 3, H, W) fp32 on the device. float frames (already at the target size) are copied;
        uint8 (N, Hs, Ws, 3) decoded frames are copied as bytes and resized + converted by pi3b_resize_u8_pil,
        bit-identical to the reference's PIL Resize + ToTensor (datasets/image_datasets.py:193-199)."""
        dev = self.config.device
        if frames_host.dtype != torch.uint8:
            return frames_host[None].to(dev, non_blocking=True)
        from .host.resample import pil_bilinear_coeffs
        n, hs, ws, _ = frames_host.shape
        H, W = self.target_size
        key = (hs, ws, H, W)
        tabs = self._resize_tables.get(key)
        if tabs is None:
            tabs = tuple(tuple(torch.from_numpy(a).to(dev).contiguous() for a in pil_bilinear_coeffs(i, o))
                         for i, o in ((hs, H), (ws, W)))
            self._resize_tables[key] = tabs
        out = torch.empty(n, 3, H, W, device=dev, dtype=torch.float32)
        ops.resize_u8_pil(frames_host.to(dev, non_blocking=True), out, tabs[0], tabs[1])
        return out[None]

    # ------------------------------------------------------------------ one chunk
    @ops.device_guard(lambda self, *a: self.config.device)
    def _process_single_chunk(self, chunk_images: torch.Tensor, chunk_paths: List[str]) -> Dict:
        assert chunk_images.ndim == 5, "Expected (B=1, N, C, H, W) tensor for chunk images"
        dev = self.config.device
        num_frames = int(chunk_images.shape[1])
        with torch.no_grad():
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            imgs_dev = chunk_images.to(dev, non_blocking=True)
            ev0.record()
            pi3_result = self.model(imgs_dev)
            ev1.record()
            ev1.synchronize()   # unlike the reference (:169-173) the reported time is device-synchronised
            dt_inf = max(1e-6, time.time() - t0)
            fps = num_frames / dt_inf if num_frames > 0 else 0.0
            print(f"   ⏱️ Inference: {dt_inf:.3f}s for {num_frames} frames  ->  {fps:.2f} FPS")
            _metrics = {"infer_s": float(dt_inf), "num_frames": int(num_frames), "fps": float(fps)}
            self.last_device_seconds = float(ev0.elapsed_time(ev1) * 1e-3)
            return self._finish_chunk(pi3_result, imgs_dev, chunk_paths, _metrics)

    @ops.device_guard(lambda self, *a: self.config.device)
    def _process_chunk_sequence_parallel(self, frames_host: torch.Tensor, num_frames: int,
                                         chunk_paths: List, ranks: Tuple[int, ...], group) -> Optional[Dict]:
        """One chunk on a group of ranks (the tail waves of `plan_waves`): every rank runs the forward on its
        contiguous slice of the frames (`split_frames`) with the chunk-global K/V all-gathered over NCCL
        (`Pi3.forward(sp_group=)`: the local K/V block is attended to while the remote blocks are in flight), the
        per-frame outputs are all-gathered in frame order and the group's first rank does the post-processing and
        returns the chunk dict (the other ranks return None). Outputs are within the bf16 tolerance of the single-GPU
        forward (bit-identical with `model.sp_overlap = False` and an even split).
        `frames_host`: all `num_frames` frames on the group's first rank, this rank's slice on the others."""
        import torch.distributed as dist
        dev = self.config.device
        rank = dist.get_rank()
        g = len(ranks)
        li = ranks.index(rank)
        split = split_frames(num_frames, g)
        counts = [n for _, n in split]
        n_max = max(counts)
        leader = li == 0
        with torch.no_grad():
            t0 = time.time()
            if leader:
                imgs_dev = self._frames_to_device(frames_host)
                local = imgs_dev[:, :counts[0]]
            else:
                imgs_dev = None
                local = self._frames_to_device(frames_host)
            out = self.model(local.contiguous(), sp_group=group, sp_frame_counts=counts)
            full = {}
            for k in ("points", "local_points", "conf", "camera_poses"):
                src = out[k][0]
                if counts[li] < n_max:      # uneven split: blocks are padded to the largest rank's frame count
                    pad = torch.zeros((n_max,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
                    pad[:counts[li]] = src
                    src = pad
                dst = torch.empty((g * n_max,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
                dist.all_gather_into_tensor(dst, src.contiguous(), group=group)
                if leader and min(counts) < n_max:
                    dst = torch.cat([dst[i * n_max:i * n_max + c] for i, c in enumerate(counts)])
                full[k] = dst[None]
            torch.cuda.synchronize()
            dt_inf = max(1e-6, time.time() - t0)
            if not leader:
                return None
            fps = num_frames / dt_inf
            print(f"   ⏱️ Inference ({g}-way sequence-parallel): {dt_inf:.3f}s for {num_frames} frames  ->  {fps:.2f} FPS")
            _metrics = {"infer_s": float(dt_inf), "num_frames": int(num_frames), "fps": float(fps), "sp_ranks": list(ranks)}
            return self._finish_chunk(full, imgs_dev, chunk_paths, _metrics)

    def _finish_chunk(self, pi3_result: Dict[str, torch.Tensor], imgs_dev: torch.Tensor, chunk_paths: List,
                      _metrics: Dict) -> Dict:
        """Everything of `_process_single_chunk` after the forward (reference :175-256)."""
        return self._enqueue_finish(pi3_result, imgs_dev, chunk_paths, _metrics).finish()

    @ops.device_guard(lambda self, *a: self.config.device)
    def _enqueue_finish(self, pi3_result: Dict[str, torch.Tensor], imgs_dev: torch.Tensor, chunk_paths: List,
                        _metrics: Dict, fwd_events=None) -> "_PendingChunk":
        """Device half of the post-processing: masks, metric scale, the 64x64 samples for the intrinsics solve, the
        keypoint gather, and asynchronous copies of the (small) results into pinned host memory, closed by an event.
        Nothing here waits for the GPU (except the MoGe pass, whose focal/shift solve is a host step), so the caller can
        enqueue the next chunk's forward before `finish()` does the host half (scipy, dict assembly)."""
        dev = self.config.device
        num_frames = int(imgs_dev.shape[1])
        pend = _PendingChunk(self, chunk_paths, _metrics, fwd_events)

        def stage(t: torch.Tensor) -> torch.Tensor:
            dst = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            dst.copy_(t, non_blocking=True)
            return dst
        with torch.no_grad():
            masks = self._compute_masks(pi3_result)[0]

            if self.moge_model is not None:
                moge_metric_depth = self.moge_model.infer(imgs_dev[0, 0])["depth"]
                pi3_metric_depth = pi3_result["local_points"][0, 0][..., 2]
                scale_factor = self._get_scale_factor_for_pi3(moge_metric_depth, pi3_metric_depth, masks[0])
                pi3_result["local_points"] = pi3_result["local_points"] * scale_factor
                pi3_result["points"] = pi3_result["points"] * scale_factor
                pi3_result["camera_poses"][:, :, :3, 3] = pi3_result["camera_poses"][:, :, :3, 3] * scale_factor
                pi3_result["camera_poses_cw"] = torch.linalg.inv(pi3_result["camera_poses"])

            if self.config.estimate_camera_params:
                try:
                    pts, val, uv_lr = sample_for_intrinsics(pi3_result["local_points"][0], pi3_result["conf"][0])
                    hw = (int(imgs_dev.shape[-2]), int(imgs_dev.shape[-1]))
                    if self.config.intrinsics_solver == "host":
                        pend.intr = ("host", stage(pts), stage(val), uv_lr) + hw
                    else:
                        pend.intr = ("device", stage(solve_focal_shift_on_device(pts, val, uv_lr))) + hw
                except Exception as e:
                    print(f"⚠️  Camera parameter estimation failed: {e}")
            pend.staged["camera_poses"] = stage(pi3_result["camera_poses"][0])

            dense = {
                "points": pi3_result["points"][0].contiguous(),
                "local_points": pi3_result["local_points"][0].contiguous(),
                "conf": pi3_result["conf"][0].contiguous(),
                "masks": masks,
                "images": imgs_dev[0].to(torch.float32).contiguous(),
            }
            done = False
            if self.keypoint_extractor is not None:
                try:
                    H, W = imgs_dev.shape[-2:]
                    kp_res = self.keypoint_extractor.extract(num_frames, H, W)
                    kps = kp_res["keypoints"].to(dev, torch.float32).contiguous()
                    interp = self._interpolate_world_points_for_keypoints(dense, kps, kp_res["keypoints"].is_cuda)
                    pend.staged["points"] = stage(interp["points"].to(torch.float16))
                    pend.staged["local_points"] = stage(interp["local_points"].to(torch.float16))
                    pend.staged["conf"] = stage(interp["conf"].to(torch.float16))
                    pend.staged["masks"] = stage(interp["masks"])
                    pend.staged["keypoints"] = stage(kps.to(torch.float16))
                    pend.staged["colors"] = stage(interp["colors"])
                    pend.staged["descriptors"] = stage(kp_res["descriptors"].to(torch.float16))
                    pend.staged["scores"] = stage(kp_res["scores"].to(torch.float16))
                    done = True
                except Exception as e:
                    print(f"⚠️  Keypoint extraction failed: {e}")
            if not done:  # keypoints disabled or failed: dense maps are stored, like the reference
                for k in ("points", "local_points", "conf"):
                    pend.staged.pop(k, None)
                    pend.host[k] = dense[k].cpu()
                for k in ("keypoints", "colors", "descriptors", "scores"):
                    pend.staged.pop(k, None)
                    pend.host.pop(k, None)
                pend.staged.pop("masks", None)
                pend.host["masks"] = masks.cpu()
            pend.event = torch.cuda.Event()
            pend.event.record()
        return pend

    @staticmethod
    def _gather_manifest(manifest: List[Dict]) -> List[Dict]:
        """The only collective on this path: every rank contributes its (small) manifest entries; the merged
        list, ordered by chunk index, is what rank 0 writes to chunks_manifest.json."""
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return manifest
        world = torch.distributed.get_world_size()
        if world == 1:
            return manifest
        gathered: List = [None] * world
        torch.distributed.all_gather_object(gathered, manifest)
        return sorted((e for part in gathered for e in part), key=lambda e: e["chunk_index"])

    # ------------------------------------------------------------------ whole sequence
    @ops.device_guard(lambda self, *a: self.config.device)
    def process_and_save(self, image_paths: Sequence, loader: Optional[Callable] = None,
                         target_size: Optional[Tuple[int, int]] = None, distributed: Optional[bool] = None) -> List[str]:
        """Process the sequence in chunks and save one file per chunk; returns the files this rank wrote.
        `loader` / `target_size` let synthetic (in-memory) sequences bypass image decoding. Under an initialised
        torch.distributed (unless distributed=False) the chunks are scheduled over the ranks by `plan_waves`: whole
        chunks per rank, and the last partial wave sequence-parallel over groups of ranks."""
        if not image_paths:
            raise ValueError("image_paths is empty")
        self.target_size = target_size or calculate_target_size(image_paths[0], pixel_limit=255000 // 2)
        print(f"Target size: {self.target_size}")

        if loader is None and self.config.device_resize:
            loader = load_images_u8
        dataset = ChunkImageDataset(image_paths, self.config.chunk_length, self.config.overlap, self.target_size,
                                    device="cpu", loader=loader)
        rank, world = 0, 1
        if distributed is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        frames_per_chunk = [e - b for b, e in dataset.chunk_indices]
        waves = plan_waves(frames_per_chunk, world, allow_sp=self.config.sequence_parallel_tail and world > 1)
        # process groups of the sequence-parallel waves: created by every rank, in plan order (a collective)
        # (cached on the creator: an NCCL communicator costs ~0.5 s to build; the all-ranks group is the default one)
        groups = self._sp_groups
        for wave in waves:
            for _, ranks in wave:
                if len(ranks) > 1 and ranks not in groups:
                    if len(ranks) == world:
                        groups[ranks] = torch.distributed.group.WORLD
                    else:
                        opts = nccl_high_priority_options() if torch.distributed.get_backend() == "nccl" else None
                        groups[ranks] = torch.distributed.new_group(ranks=list(ranks), pg_options=opts)
        mine = [(ci, ranks) for wave in waves for ci, ranks in wave if rank in ranks]
        wave_of = {ci: wi for wi, wave in enumerate(waves) for ci, _ in wave}
        t_run0 = time.time()
        # run statistics (bench.py reads them): when each of this rank's chunks was completely finished (saved, or for
        # a non-leading sequence-parallel rank: its forward done), relative to the start of the run
        self.last_run_stats = {"rank": rank, "world": world, "waves": [[(ci, list(rk)) for ci, rk in w] for w in waves],
                               "chunk_done_s": {}, "chunk_wave": wave_of}

        saved_files: List[str] = []
        manifest: List[Dict] = []
        infer_times: List[float] = []
        infer_frames: List[int] = []
        per_chunk_fps: List[float] = []
        print(f"🔄 Processing {len(mine)} of {len(dataset)} chunks on rank {rank}/{world}...")
        pin = self.config.pin_memory

        def fetch(job: Tuple[int, Tuple[int, ...]]) -> torch.Tensor:
            """Host frames of one job: the whole chunk, or (sequence-parallel, not the group's first rank) this rank's
            slice. Runs on the loader threads, `num_loader_workers` chunks ahead of the GPU, like the reference's
            DataLoader(num_workers, prefetch_factor=1, pin_memory) (:281-289)."""
            ci, rk = job
            b, e = dataset.chunk_indices[ci]
            li = rk.index(rank)
            if li > 0:
                first, cnt = split_frames(e - b, len(rk))[li]
                b, e = b + first, b + first + cnt
            # frames of one chunk are decoded by several threads, then gathered into one (pinned) buffer
            step = max(1, -(-(e - b) // frame_threads))
            parts = list(frame_pool.map(lambda lo: dataset.load_frames(lo, min(lo + step, e)), range(b, e, step)))
            if any(p.dtype != parts[0].dtype or p.shape[1:] != parts[0].shape[1:] for p in parts):
                # mixed source sizes: finish those frames with the reference's host resize
                parts = [p if p.dtype == torch.float32 else
                         torch.stack([resize_u8_on_host(f.numpy(), self.target_size) for f in p]) for p in parts]
            out = torch.empty((e - b,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype, pin_memory=pin)
            torch.cat(parts, out=out)
            return out

        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        depth = max(1, int(self.config.num_loader_workers))
        frame_threads = max(2, min(8, (os.cpu_count() or 8) // max(1, world)))
        pool = ThreadPoolExecutor(max_workers=depth)
        frame_pool = ThreadPoolExecutor(max_workers=frame_threads)
        pending: deque = deque(pool.submit(fetch, job) for job in mine[:depth])
        t_wait = t_proc = t_save = 0.0

        def save(chunk_result: Dict, chunk_idx: int, chunk_paths: List) -> None:
            start_idx, end_idx = dataset.chunk_indices[chunk_idx]
            m = chunk_result.get("_metrics", {})
            if m:
                infer_times.append(float(m.get("infer_s", 0.0)))
                infer_frames.append(int(m.get("num_frames", 0)))
                per_chunk_fps.append(float(m.get("fps", 0.0)))
            out_name = f"chunk_{chunk_idx:06d}.pt"
            out_path = os.path.join(self.chunks_dir, out_name)
            chunk_result["chunk_index"] = chunk_idx
            chunk_result["start_idx"] = start_idx
            chunk_result["end_idx"] = end_idx
            try:
                torch.save(chunk_result, out_path)
                saved_files.append(out_path)
                manifest.append({"chunk_index": chunk_idx, "file": out_name, "start_idx": start_idx,
                                 "end_idx": end_idx, "num_frames": len(chunk_paths), "image_paths": chunk_paths})
                print(f"   💾 Saved: {out_path}")
            except Exception as e:
                print(f"❌ Failed to save chunk {chunk_idx}: {e}")
            self.last_run_stats["chunk_done_s"][chunk_idx] = time.time() - t_run0

        # Software pipeline over this rank's whole-chunk jobs: the forward of chunk k+1 is enqueued (asynchronously)
        # before the host half of chunk k (intrinsics LM, dict assembly, torch.save) runs, so the GPU never waits for it.
        prev: Optional[Tuple[_PendingChunk, int, List]] = None

        def drain() -> None:
            nonlocal prev, t_save
            if prev is not None:
                _t = time.time()
                pend, ci, cp = prev
                prev = None
                save(pend.finish(), ci, cp)
                t_save += time.time() - _t

        dev = self.config.device
        for pos, (chunk_idx, ranks) in enumerate(mine):
            _t = time.time()
            load_error = None
            try:
                frames_host = pending.popleft().result()
            except Exception as e:             # unreadable frame(s): report and skip the chunk, like the reference
                frames_host, load_error = None, e
            t_wait += time.time() - _t
            if pos + depth < len(mine):
                pending.append(pool.submit(fetch, mine[pos + depth]))
            if len(ranks) > 1:
                # a chunk shared by several ranks is entered by all of them or by none (nobody is left in the all-gather)
                if not group_all_ok(load_error is None, groups[ranks], self.config.device):
                    print(f"❌ Chunk {chunk_idx + 1}: frames could not be loaded on every rank of {ranks}"
                          + (f" ({load_error})" if load_error is not None else "") + " — skipped")
                    continue
            elif load_error is not None:
                print(f"❌ Chunk {chunk_idx + 1}: frames could not be loaded ({load_error}) — skipped")
                continue
            start_idx, end_idx = dataset.chunk_indices[chunk_idx]
            # default_collate turns the nested path list into 1-tuples (reference Appendix B)
            chunk_paths = [(p,) for p in dataset.image_paths[start_idx:end_idx]]
            print(f"📦 Chunk {chunk_idx + 1}/{len(dataset)}: frames {start_idx + 1}-{end_idx}")
            _t = time.time()
            if len(ranks) > 1:
                drain()
                if self.config.keypoint_seed is not None:
                    torch.manual_seed(self.config.keypoint_seed + chunk_idx)
                chunk_result = self._process_chunk_sequence_parallel(frames_host, end_idx - start_idx, chunk_paths, ranks,
                                                                     groups[ranks])
                t_proc += time.time() - _t
                if chunk_result is not None:      # only the group's first rank saves
                    _t = time.time()
                    save(chunk_result, chunk_idx, chunk_paths)
                    t_save += time.time() - _t
                else:
                    self.last_run_stats["chunk_done_s"][chunk_idx] = time.time() - t_run0
                continue
            num_frames = end_idx - start_idx
            with torch.no_grad():
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                imgs_dev = self._frames_to_device(frames_host)
                ev0.record()
                pi3_result = self.model(imgs_dev)
                ev1.record()
            t_proc += time.time() - _t
            drain()                                   # host half of the previous chunk, while this forward runs
            _t = time.time()
            if self.config.keypoint_seed is not None:
                torch.manual_seed(self.config.keypoint_seed + chunk_idx)
            metrics = {"infer_s": 0.0, "num_frames": int(num_frames), "fps": 0.0}
            prev = (self._enqueue_finish(pi3_result, imgs_dev, chunk_paths, metrics,
                                         fwd_events=(time.time(), ev0, ev1, num_frames)), chunk_idx, chunk_paths)
            t_proc += time.time() - _t
        drain()

        pool.shutdown(wait=True)
        frame_pool.shutdown(wait=True)
        self.last_run_stats.update(wait_frames_s=t_wait, enqueue_s=t_proc, host_half_s=t_save,
                                   total_s=time.time() - t_run0)
        print(f"   host pipeline on rank {rank}: waited {t_wait:.3f}s for frames, {t_proc:.3f}s enqueueing chunks, "
              f"{t_save:.3f}s in the host half (wait for the GPU, intrinsics solve, save)")
        total_time = sum(infer_times)
        total_frames = sum(infer_frames)
        if total_time > 0:
            print(f"\n⏱️ Overall inference: {total_frames} frames in {total_time:.3f}s  ->  "
                  f"{total_frames / total_time:.2f} FPS (weighted)")
        steady = sorted(f for f, n in zip(per_chunk_fps, infer_frames) if n == self.config.chunk_length)
        if steady:
            print(f"   Steady-state FPS (full {self.config.chunk_length}-frame chunks, median): "
                  f"{steady[len(steady) // 2]:.2f} FPS")

        if world > 1:
            manifest = self._gather_manifest(manifest)
        if rank == 0:
            try:
                with open(os.path.join(self.config.output_dir, "chunks_manifest.json"), "w") as f:
                    json.dump(manifest, f, indent=2)
            except Exception as e:
                print(f"⚠️  Failed to write manifest: {e}")
            try:
                metadata = {"chunk_length": int(self.config.chunk_length), "overlap": int(self.config.overlap),
                            "target_size": list(self.target_size) if self.target_size is not None else None}
                with open(os.path.join(self.config.output_dir, "chunk_metadata.json"), "w") as f:
                    json.dump(metadata, f, indent=2)
            except Exception as e:
                print(f"⚠️  Failed to write chunk metadata: {e}")
        print(f"✅ Completed. Saved {len(saved_files)} chunks to {self.chunks_dir}")
        return saved_files
