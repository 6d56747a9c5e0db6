"""2+ GPUs under torchrun: the hybrid chunk schedule (whole chunks per rank, last wave sequence-parallel) writes the
same chunk files, bit for bit, as a single-GPU run of the same sequence; prints both makespans.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/hybrid_check.py [frames H W L ov]
"""
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
from pi3_slam_b200.host.datasets import chunk_index_list, plan_waves  # noqa: E402
from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig  # noqa: E402
from pi3_slam_b200.pi3 import Pi3  # noqa: E402
from pi3_slam_b200.synthetic import make_frames, make_state_dict  # noqa: E402


def main():
    skip_single = "nosingle" in sys.argv
    a = [int(v) for v in sys.argv[1:6] if v != "nosingle"]
    frames, H, W, L, ov = a if len(a) == 5 else (20, 84, 112, 8, 2)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    model = Pi3()
    model.load_state_dict(make_state_dict(0), strict=True)
    model = model.cuda().eval()
    paths = [f"/synthetic/{i:05d}.png" for i in range(frames)]
    index = {p: i for i, p in enumerate(paths)}
    loader = lambda ps, ts: torch.cat([make_frames(1, *ts, start_index=index[p]) for p in ps])  # noqa: E731
    box = [os.environ.get("HYBRID_OUT") or tempfile.mkdtemp()]
    dist.broadcast_object_list(box, src=0)          # every rank writes under rank 0's directory
    root = box[0]
    base = dict(model_path="unused", device="cuda", do_metric_depth=False, keypoint_type="grid", max_num_keypoints=200,
                chunk_length=L, overlap=ov, keypoint_seed=5)
    out = {"world": world, "frames": frames, "size": [H, W], "chunk": [L, ov]}
    chunks = chunk_index_list(frames, L, ov)
    out["waves"] = [[(c, list(r)) for c, r in w] for w in plan_waves([e - b for b, e in chunks], world)]
    times = {}
    for mode, sp in (("hybrid", True), ("round_robin", False)):
        cr = OfflineChunkCreator(OfflineCreatorConfig(output_dir=os.path.join(root, mode), sequence_parallel_tail=sp, **base),
                                 model=model)
        for rep in range(2 if mode == "hybrid" else 1):     # first pass warms allocations / NCCL communicators
            dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
            cr.process_and_save(paths, loader=loader, target_size=(H, W))
            dist.barrier(); torch.cuda.synchronize()
            times[mode] = time.time() - t0
    out["makespan_s"] = times
    if rank == 0 and skip_single:
        out["total_frames_processed"] = sum(e - b for b, e in chunks)
        out["frames_per_s"] = {k: out["total_frames_processed"] / v for k, v in times.items()}
        print(json.dumps(out))
    elif rank == 0:
        single = OfflineChunkCreator(OfflineCreatorConfig(output_dir=os.path.join(root, "single"), **base), model=model)
        t0 = time.time()
        files = single.process_and_save(paths, loader=loader, target_size=(H, W), distributed=False)
        times["single_gpu"] = time.time() - t0
        same = True
        for f in files:
            a_ = torch.load(f, weights_only=False)
            for mode in ("hybrid", "round_robin"):
                b_ = torch.load(os.path.join(root, mode, "chunks", os.path.basename(f)), weights_only=False)
                for k in ("points", "local_points", "conf", "masks", "camera_poses", "keypoints", "colors", "intrinsics"):
                    if not torch.equal(a_[k], b_[k]):
                        same = False
                        print("MISMATCH", mode, os.path.basename(f), k, float((a_[k].float() - b_[k].float()).abs().max()))
        man = json.load(open(os.path.join(root, "hybrid", "chunks_manifest.json")))
        out["manifest_chunks"] = [m["chunk_index"] for m in man]
        out["bit_identical_to_single_gpu"] = same
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
