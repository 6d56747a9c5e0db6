"""BASELINE configs[3] end to end on N GPUs (torchrun, one rank per GPU): pi3 + MoGe metric scale (native MoGe-2, synthetic
weights) + grid keypoints (max-kp 200) + intrinsics + chunk files on a EuRoC-shaped 752x480 synthetic sequence (target
280x448), chunk 100 / overlap 20, hybrid chunk / sequence-parallel schedule; then the Sim3 chain between consecutive
chunks on rank 0 (one batched pi3b_umeyama_sim3_batched launch). Prints one JSON line on rank 0.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/cfg4_run.py [frames]
"""
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
from pi3_slam_b200.alignment import align_consecutive_chunks  # noqa: E402
from pi3_slam_b200.host.datasets import nccl_high_priority_options  # noqa: E402
from pi3_slam_b200.moge import MoGeModel  # noqa: E402
from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig  # noqa: E402
from pi3_slam_b200.pi3 import Pi3  # noqa: E402
from pi3_slam_b200.synthetic import make_frames, make_moge_state_dict, make_state_dict, moge_synth_config  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, pg_options=nccl_high_priority_options())
    pi3 = Pi3()
    pi3.load_state_dict(make_state_dict(0), strict=True)
    pi3 = pi3.to(dev).eval()
    mcfg = moge_synth_config()
    moge = MoGeModel(**mcfg)
    moge.load_state_dict(make_moge_state_dict(mcfg, 0), strict=True)
    moge = moge.to(dev).eval()
    H, W, L, ov = 280, 448, 100, 20
    paths = [f"/synthetic/euroc/{i:05d}.png" for i in range(frames)]
    index = {p: i for i, p in enumerate(paths)}
    loader = lambda ps, ts: torch.cat([make_frames(1, *ts, start_index=index[p]) for p in ps])  # noqa: E731
    box = [tempfile.mkdtemp(prefix="pi3b_cfg4_")]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    cfg = OfflineCreatorConfig(model_path="unused", output_dir=box[0], device=str(dev), do_metric_depth=True,
                               keypoint_type="grid", max_num_keypoints=200, chunk_length=L, overlap=ov, keypoint_seed=7)
    cr = OfflineChunkCreator(cfg, model=pi3, moge_model=moge)
    times = []
    saved = sys.stdout
    for rep in range(2):                    # first pass warms allocations and NCCL groups
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sys.stdout = open(os.devnull, "w")
        t0 = time.perf_counter()
        cr.process_and_save(paths, loader=loader, target_size=(H, W))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        sys.stdout.close()
        sys.stdout = saved
    if rank == 0:
        files = sorted(os.path.join(box[0], "chunks", f) for f in os.listdir(os.path.join(box[0], "chunks")))
        chunks = [torch.load(f, weights_only=False) for f in files]
        t0 = time.perf_counter()
        full = [c for c in chunks if c["points"].shape[0] == L]
        res = align_consecutive_chunks(full, L, ov)
        torch.cuda.synchronize()
        t_align = time.perf_counter() - t0
        scales = [float(abs(torch.linalg.det(p[:3, :3].double())) ** (1 / 3)) for p in res["pairwise"]]
        finite = all(bool(torch.isfinite(c["points"].float()).all()) for c in chunks)
        print(json.dumps({"workload": f"cfg4: {frames} frames 752x480-shaped (280x448), chunk {L}/{ov}, MoGe metric scale, "
                                      f"grid keypoints 200, intrinsics, chunk files, Sim3 chain", "n_gpus": world,
                          "chunks": len(chunks), "makespan_s": times[-1], "first_pass_s": times[0],
                          "sequence_frames_per_s": frames / times[-1], "sim3_chain_s": t_align, "sim3_pairs": len(scales),
                          "sim3_scale_min_max": [min(scales), max(scales)] if scales else None, "outputs_finite": finite,
                          "waves": [[len(r) for _, r in w] for w in cr.last_run_stats["waves"]]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
