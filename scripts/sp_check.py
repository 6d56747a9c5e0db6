"""Sequence-parallel check (run under torchrun on >= 2 GPUs of one box):
   (1) kernel: local queries vs gathered K/V equals the rows of the full attention;
   (2) model: Pi3.forward(local frames, sp_group=WORLD) equals the matching frames of a single-GPU forward of the
       whole chunk; (3) optional throughput of an over-long chunk (--frames N --size H W)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pi3_slam_b200 import ops  # noqa: E402
from pi3_slam_b200.pi3 import Pi3  # noqa: E402
from pi3_slam_b200.synthetic import make_frames, make_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--size", type=int, nargs=2, default=[308, 406])
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}

    # (1) kernel-level
    torch.manual_seed(0)
    S, H = 4 * 300 * world, 16
    qkv = torch.randn(S, 3 * H * 64, device=dev).bfloat16()
    full = torch.empty(S, H * 64, device=dev, dtype=torch.bfloat16)
    ops.attention_qkv(qkv, full, 1, S, H)
    loc = S // world
    part = torch.empty(loc, H * 64, device=dev, dtype=torch.bfloat16)
    ops.attention_q_kv(qkv[rank * loc:(rank + 1) * loc], qkv[:, H * 64:].contiguous(), part, 1, loc, S, H)
    d = (part.float() - full[rank * loc:(rank + 1) * loc].float()).abs().max().item()
    res["kernel_max_abs_diff"] = d
    assert d < 2e-2, d

    # (2) model-level
    sd = make_state_dict(0)
    model = Pi3()
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    n_loc, Hh, Ww = 2, 56, 84
    frames = make_frames(n_loc * world, Hh, Ww)[None].to(dev)
    ref = model(frames)
    ref = {k: v[:, rank * n_loc:(rank + 1) * n_loc].clone() for k, v in ref.items()}
    out = model(frames[:, rank * n_loc:(rank + 1) * n_loc].contiguous(), sp_group=dist.group.WORLD)
    errs = {k: ((out[k] - ref[k]).norm() / ref[k].norm()).item() for k in out}
    res["model_rel_l2_vs_single_gpu"] = errs
    assert max(errs.values()) < 2e-3, errs

    # (3) throughput of one over-long chunk, frames split evenly over the group
    if args.frames:
        assert args.frames % world == 0, "frames must divide evenly over the ranks"
        n_loc = args.frames // world
        Hh, Ww = args.size
        fr = make_frames(n_loc, Hh, Ww, start_index=rank * n_loc)[None].to(dev)
        for _ in range(2):
            model(fr, sp_group=dist.group.WORLD)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            model(fr, sp_group=dist.group.WORLD)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item() / args.steps
        fl = Pi3.flops(args.frames, Hh, Ww)["total"]
        res["sp_chunk"] = {"frames": args.frames, "size": [Hh, Ww], "ms_per_chunk": ms,
                           "frames_per_s": args.frames / (ms * 1e-3), "tflops_total": fl / (ms * 1e-3) / 1e12,
                           "tflops_per_gpu": fl / (ms * 1e-3) / 1e12 / world}
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
