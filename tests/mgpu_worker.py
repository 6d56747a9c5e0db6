"""Worker of tests/test_multigpu_gpu.py (and bench.py's multi-GPU extras): run under torchrun, one rank per GPU.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_worker.py sp|hybrid OUT.json

sp:     (1) kernel: local queries against gathered K/V == rows of the full attention; (2) model: the sequence-parallel
        forward (`Pi3.forward(sp_group=)`) of an UNEVEN frame split equals the single-GPU forward of the whole chunk —
        bit-identical with the blocking all-gather on an even split, within the bf16 tolerance with the overlapped
        (local block first, LSE merge) variant that is the default.
hybrid: `process_and_save` under the hybrid schedule (whole chunks per rank, last wave sequence-parallel) writes chunk
        files whose keypoints / bookkeeping are bit-identical to a single-GPU run and whose maps are within tolerance
        (bit-identical for the chunks that were not sequence-parallel).
Rank 0 writes the result dict to OUT.json; every assertion failure exits non-zero."""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
from pi3_slam_b200 import ops  # noqa: E402
from pi3_slam_b200.host.datasets import chunk_index_list, nccl_high_priority_options, plan_waves, split_frames  # noqa: E402
from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig  # noqa: E402
from pi3_slam_b200.pi3 import Pi3  # noqa: E402
from pi3_slam_b200.synthetic import make_frames, make_state_dict  # noqa: E402

TOL = dict(points=1e-2, local_points=1e-2, conf=2e-2, camera_poses=2e-3)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def check_sp(model, dev, rank, world) -> dict:
    res = {}
    torch.manual_seed(0)
    S, H = 4 * 300 * world, 16
    qkv = torch.randn(S, 3 * H * 64, device=dev).bfloat16()
    full = torch.empty(S, H * 64, device=dev, dtype=torch.bfloat16)
    ops.attention_qkv(qkv, full, 1, S, H)
    loc = S // world
    part = torch.empty(loc, H * 64, device=dev, dtype=torch.bfloat16)
    ops.attention_q_kv(qkv[rank * loc:(rank + 1) * loc], qkv[:, H * 64:].contiguous(), part, 1, loc, S, H)
    res["kernel_max_abs_diff"] = (part.float() - full[rank * loc:(rank + 1) * loc].float()).abs().max().item()
    assert res["kernel_max_abs_diff"] < 2e-2
    Hh, Ww = 56, 84
    for label, total, overlap in (("even_blocking", 2 * world, False), ("even_overlap", 2 * world, True),
                                  ("uneven_overlap", 2 * world + 1, True)):
        split = split_frames(total, world)
        first, cnt = split[rank]
        frames = make_frames(total, Hh, Ww)[None].to(dev)
        ref = {k: v[:, first:first + cnt].clone() for k, v in model(frames).items()}
        model.sp_overlap = overlap
        out = model(frames[:, first:first + cnt].contiguous(), sp_group=dist.group.WORLD,
                    sp_frame_counts=[n for _, n in split])
        model.sp_overlap = True
        errs = {k: rel(out[k], ref[k]) for k in out}
        same = all(torch.equal(out[k], ref[k]) for k in out)
        flags = torch.tensor([1.0 if same else 0.0] + [errs[k] for k in sorted(errs)], device=dev, dtype=torch.float64)
        dist.all_reduce(flags[:1], op=dist.ReduceOp.MIN)
        dist.all_reduce(flags[1:], op=dist.ReduceOp.MAX)
        res[label] = {"bit_identical": bool(flags[0].item() == 1.0), "max_rel_l2": dict(zip(sorted(errs), flags[1:].tolist()))}
        if not overlap:
            assert res[label]["bit_identical"], res[label]
        for k, v in res[label]["max_rel_l2"].items():
            assert v < TOL[k], (label, k, v)
    return res


def check_hybrid(model, dev, rank, world, frames=28, H=84, W=112, L=8, ov=2) -> dict:
    paths = [f"/synthetic/{i:05d}.png" for i in range(frames)]
    index = {p: i for i, p in enumerate(paths)}
    loader = lambda ps, ts: torch.cat([make_frames(1, *ts, start_index=index[p]) for p in ps])  # noqa: E731
    box = [tempfile.mkdtemp()]
    dist.broadcast_object_list(box, src=0)          # every rank writes under rank 0's directory
    root = box[0]
    base = dict(model_path="unused", device=str(dev), do_metric_depth=False, keypoint_type="grid", max_num_keypoints=60,
                chunk_length=L, overlap=ov, keypoint_seed=5)
    chunks = chunk_index_list(frames, L, ov)
    waves = plan_waves([e - b for b, e in chunks], world)
    sp_chunks = {c for w in waves for c, r in w if len(r) > 1}
    cr = OfflineChunkCreator(OfflineCreatorConfig(output_dir=os.path.join(root, "hybrid"), sequence_parallel_tail=True,
                                                  **base), model=model)
    cr.process_and_save(paths, loader=loader, target_size=(H, W))
    dist.barrier()
    res = {"waves": [[(c, list(r)) for c, r in w] for w in waves], "sp_chunks": sorted(sp_chunks)}
    if rank == 0:
        single = OfflineChunkCreator(OfflineCreatorConfig(output_dir=os.path.join(root, "single"), **base), model=model)
        files = single.process_and_save(paths, loader=loader, target_size=(H, W), distributed=False)
        worst = 0.0
        for f in files:
            a = torch.load(f, weights_only=False)
            b = torch.load(os.path.join(root, "hybrid", "chunks", os.path.basename(f)), weights_only=False)
            ci = a["chunk_index"]
            for k in ("keypoints", "colors", "descriptors", "scores"):
                assert torch.equal(a[k], b[k]), (ci, k)
            assert a["image_paths"] == b["image_paths"] and (a["start_idx"], a["end_idx"]) == (b["start_idx"], b["end_idx"])
            for k in ("points", "local_points", "conf", "camera_poses"):
                if ci in sp_chunks:
                    e = rel(b[k].float(), a[k].float())
                    worst = max(worst, e)
                    assert e < 2e-2, (ci, k, e)          # fp16-stored samples of maps that agree to bf16 tolerance
                else:
                    assert torch.equal(a[k], b[k]), (ci, k)
        man = json.load(open(os.path.join(root, "hybrid", "chunks_manifest.json")))
        assert [m["chunk_index"] for m in man] == list(range(len(chunks)))
        res["worst_sp_rel_l2"] = worst
        res["files"] = len(files)
    dist.barrier()
    return res


def main():
    what, out_path = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, pg_options=nccl_high_priority_options())
    model = Pi3()
    model.load_state_dict(make_state_dict(0), strict=True)
    model = model.to(dev).eval()
    res = {"world": world}
    res[what] = check_sp(model, dev, rank, world) if what == "sp" else check_hybrid(model, dev, rank, world)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
