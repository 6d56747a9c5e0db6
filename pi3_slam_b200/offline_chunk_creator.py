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
from .host.camera_estimation import estimate_camera_parameters_single_chunk
from .host.datasets import ChunkImageDataset, plan_waves
from .host.image_utils import calculate_target_size
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
    def _interpolate_world_points_for_keypoints(dense: Dict, keypoints: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Sample points / local_points (bilinear) and conf / masks (nearest) at keypoints (reference :129-159)."""
        images = dense.get("images")
        p, l, c, m, col = ops.keypoint_gather(dense["points"], dense["local_points"], dense["conf"],
                                              dense["masks"].to(torch.uint8).contiguous(), images, keypoints)
        return {"points": p, "local_points": l, "conf": c, "masks": m, "colors": col}

    # ------------------------------------------------------------------ one chunk
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

    def _process_chunk_sequence_parallel(self, load_frames: Callable[[int, int], torch.Tensor], num_frames: int,
                                         chunk_paths: List, ranks: Tuple[int, ...], group) -> Optional[Dict]:
        """One chunk on a group of ranks (the tail waves of `plan_waves`): every rank runs the forward on its
        contiguous slice of the frames with the chunk-global K/V all-gathered over NCCL (`Pi3.forward(sp_group=)`), the
        per-frame outputs are all-gathered in frame order and the group's first rank does the post-processing and
        returns the chunk dict (the other ranks return None). Outputs are bit-identical to the single-GPU forward."""
        import torch.distributed as dist
        dev = self.config.device
        rank = dist.get_rank()
        g = len(ranks)
        li = ranks.index(rank)
        n_loc = num_frames // g
        leader = li == 0
        with torch.no_grad():
            t0 = time.time()
            if leader:
                imgs_dev = load_frames(0, num_frames)[None].to(dev, non_blocking=True)
                local = imgs_dev[:, :n_loc]
            else:
                imgs_dev = None
                local = load_frames(li * n_loc, (li + 1) * n_loc)[None].to(dev, non_blocking=True)
            out = self.model(local.contiguous(), sp_group=group)
            full = {}
            for k in ("points", "local_points", "conf", "camera_poses"):
                src = out[k][0].contiguous()
                dst = torch.empty((num_frames,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
                dist.all_gather_into_tensor(dst, src, group=group)
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
        """Everything of `_process_single_chunk` after the forward (reference :175-256), on the device."""
        dev = self.config.device
        num_frames = int(imgs_dev.shape[1])
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

            camera_params = None
            if self.config.estimate_camera_params:
                try:
                    camera_params = estimate_camera_parameters_single_chunk(pi3_result["local_points"][0],
                                                                            pi3_result["conf"][0])
                except Exception as e:
                    print(f"⚠️  Camera parameter estimation failed: {e}")
                    camera_params = None

            result: Dict = {
                "camera_poses": pi3_result["camera_poses"][0].cpu(),
                "image_paths": chunk_paths,
                "_metrics": _metrics,
            }
            if camera_params is not None:
                result["camera_params"] = {k: v.cpu() for k, v in camera_params.items()}
            if self.target_size is not None:
                result["original_width"] = self.target_size[1]
                result["original_height"] = self.target_size[0]

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
                    interp = self._interpolate_world_points_for_keypoints(dense, kps)
                    result["points"] = interp["points"].to(torch.float16).cpu()
                    result["local_points"] = interp["local_points"].to(torch.float16).cpu()
                    result["conf"] = interp["conf"].to(torch.float16).cpu()
                    result["masks"] = interp["masks"].cpu()
                    result["keypoints"] = kps.to(torch.float16).cpu()
                    result["descriptors"] = kp_res["descriptors"].to(torch.float16).cpu()
                    result["scores"] = kp_res["scores"].to(torch.float16).cpu()
                    result["colors"] = interp["colors"].cpu().to(torch.float16)
                    done = True
                except Exception as e:
                    print(f"⚠️  Keypoint extraction failed: {e}")
            if not done:  # keypoints disabled or failed: dense maps are stored, like the reference
                result["points"] = dense["points"].cpu()
                result["local_points"] = dense["local_points"].cpu()
                result["conf"] = dense["conf"].cpu()
                result["masks"] = masks.cpu()
            if result.get("camera_params") is not None:
                result["intrinsics"] = result["camera_params"].get("intrinsics", None)
        return result

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
                    groups[ranks] = (torch.distributed.group.WORLD if len(ranks) == world
                                     else torch.distributed.new_group(ranks=list(ranks)))
        mine = [(ci, ranks) for wave in waves for ci, ranks in wave if rank in ranks]

        saved_files: List[str] = []
        manifest: List[Dict] = []
        infer_times: List[float] = []
        infer_frames: List[int] = []
        per_chunk_fps: List[float] = []
        print(f"🔄 Processing {len(mine)} of {len(dataset)} chunks on rank {rank}/{world}...")
        pin = self.config.pin_memory
        for chunk_idx, ranks in mine:
            start_idx, end_idx = dataset.chunk_indices[chunk_idx]
            # default_collate turns the nested path list into 1-tuples (reference Appendix B)
            chunk_paths = [(p,) for p in dataset.image_paths[start_idx:end_idx]]
            print(f"📦 Chunk {chunk_idx + 1}/{len(dataset)}: frames {start_idx + 1}-{end_idx}")
            if self.config.keypoint_seed is not None:
                torch.manual_seed(self.config.keypoint_seed + chunk_idx)
            if len(ranks) > 1:
                def load_frames(a: int, b: int, _s=start_idx) -> torch.Tensor:
                    t = dataset.load_frames(_s + a, _s + b)
                    return t.pin_memory() if pin else t
                chunk_result = self._process_chunk_sequence_parallel(load_frames, end_idx - start_idx, chunk_paths, ranks,
                                                                     groups[ranks])
                if chunk_result is None:      # not the group's first rank: nothing to save
                    continue
            else:
                chunk_images = dataset[chunk_idx]["chunk"][None]
                if pin:
                    chunk_images = chunk_images.pin_memory()
                chunk_result = self._process_single_chunk(chunk_images, chunk_paths)
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
