#!/usr/bin/env python
"""Headline benchmark: pi3 chunk frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg1]

A "step" is one pi3 forward over one synthetic chunk (default workload cfg2 = BASELINE.json configs[1]:
100 frames at 378x518, bf16 tensor cores). `value` times steps with the chunk already resident in HBM;
`e2e` times the same steps through the public `Pi3.forward` call from pinned HOST frames to pinned HOST
results (H2D of the frames and D2H of all four output maps of EVERY step inside the timed region; like the product's
own chunk loop, the uploads and downloads run on a copy stream next to the neighbouring steps' compute). For N > 1 (launched
by torch.distributed.run, one rank per GPU) every rank runs its own chunk — chunks are independent, there is
no data-path collective — and `value` is total frames over the max-over-ranks device time ("weak" scaling).

`--impl reference` times the reference's CPU algorithm (the oracle port: /root/reference is a Python package
that does not exist on the GPU box) on the host cores, on a bounded sample of the same workload (8 frames of the
workload's resolution).

Outside the timed regions the line also carries
  parity        the timed step's outputs against the oracle evaluated on the GPU in strict fp32 (N = 1 only);
  cpu_baseline  the 8-frame CPU sample and one SAME-WORK pair (cfg1, 8 x 224 x 224, in full on both sides);
  extra.cfg3_sequence   `process_and_save` on the 1000-frame 308x406 synthetic sequence under the hybrid schedule
                        (whole chunks per rank, last waves sequence-parallel): makespan, frames/s, per-wave split;
  extra.cfg5_sp         (N >= 2) one 300-frame chunk, sequence-parallel over all N ranks: ms/chunk, all-gather
                        ms/layer and how much of it the local-block attention did not hide.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (frames, H, W, description)
    "cfg1": (8, 224, 224, "BASELINE configs[0]: one 8-frame 224x224 chunk"),
    "cfg2": (100, 378, 518, "BASELINE configs[1]: one 100-frame chunk at 378x518 (P=999, T=1004, 100400 tokens)"),
    "cfg3": (100, 308, 406, "BASELINE configs[2] chunk shape: 100 frames at 308x406 (640x480 source), 64300 tokens"),
}
CPU_SAMPLE_FRAMES = 8


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured")
    return dict(burst=1590.0, sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(frames: int, H: int, W: int, steps: int, warmup: int, sd=None) -> dict:
    """Oracle (CPU port of the reference forward) on `frames` frames of the workload resolution."""
    import torch
    from oracle import pi3_oracle
    from pi3_slam_b200.synthetic import make_frames, make_state_dict
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently turn the CPU arm into a single-core run
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    if torch.get_num_threads() < ncores:
        torch.set_num_threads(ncores)
    sd = sd if sd is not None else make_state_dict(0)
    imgs = make_frames(frames, H, W)[None]
    for _ in range(warmup):
        pi3_oracle.pi3_forward(imgs, sd)
    t0 = time.perf_counter()
    for _ in range(steps):
        pi3_oracle.pi3_forward(imgs, sd)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return {"value": frames / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{frames} frames of the workload at {H}x{W} ({frames * ((H // 14) * (W // 14) + 5)} tokens), "
                      f"fp32 eager torch on {torch.get_num_threads()} threads, {dt:.2f} s/step; global attention is "
                      f"quadratic in frames, so full-chunk CPU throughput is lower than this sample's",
            "s_per_step": dt}


def parity_block(model, dev_frames, out, sd) -> dict:
    """Outputs of the timed step against the oracle evaluated on the same GPU in strict fp32 (TF32 off), same inputs,
    same weights, the whole chunk (chunk-global attention couples every frame). Outside the timed region."""
    import torch
    from oracle import pi3_oracle
    t0 = time.perf_counter()
    mine = {k: v.clone() for k, v in out.items()}
    model._invalidate()                       # release the engine's arena and packed weights while the oracle runs
    torch.cuda.empty_cache()
    ref = pi3_oracle.pi3_forward_gpu_fp32(dev_frames, sd, device=dev_frames.device)
    res = {k: pi3_oracle.rel_l2(mine[k], ref[k]) for k in ("points", "local_points", "conf")}
    res["rot_rad"], res["rel_t"] = pi3_oracle.pose_errors(mine["camera_poses"].cpu(), ref["camera_poses"].cpu())
    tol = dict(points=1e-2, local_points=1e-2, conf=2e-2, rot_rad=1e-3, rel_t=1e-3)
    del ref, mine
    torch.cuda.empty_cache()
    return {"against": "oracle.pi3_oracle.pi3_forward on the same GPU, fp32, TF32 off, all frames of the timed chunk",
            "errors": res, "tolerance": tol, "within_tolerance": all(res[k] < tol[k] for k in tol),
            "seconds": time.perf_counter() - t0}


def same_work_pair(model, sd, dev) -> dict:
    """cfg1 (BASELINE configs[0], 8 x 224 x 224) in full on both sides: the one GPU/CPU ratio that compares equal work."""
    import torch
    from pi3_slam_b200.synthetic import make_frames
    frames, H, W, _ = WORKLOADS["cfg1"]
    host = make_frames(frames, H, W)[None].pin_memory()
    host_out = None
    for _ in range(3):
        o = model(host.to(dev, non_blocking=True))
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in o.items()}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        o = model(host.to(dev, non_blocking=True))
        for k, v in o.items():
            host_out[k].copy_(v, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / n
    cpu = cpu_reference_run(frames, H, W, 1, 0, sd)
    return {"workload": "cfg1: 8 frames at 224x224 (2088 tokens), whole chunk on both sides, host frames to host results",
            "gpu_frames_per_s": frames / (gpu_ms * 1e-3), "gpu_ms": gpu_ms, "cpu_frames_per_s": cpu["value"],
            "cpu_s": cpu["s_per_step"], "cpu_cores": cpu["cores"], "ratio": frames / (gpu_ms * 1e-3) / cpu["value"]}


def gemm_roofline(dev, peaks) -> dict:
    """The four trunk GEMM shapes of the workload (qkv, proj + LayerScale + residual, fc1 + GELU, fc2 + LayerScale +
    residual at M = 100 400), each timed alone with CUDA events (burst conditions), next to cuBLAS on the same box."""
    import torch
    from pi3_slam_b200 import ops
    M = 100400
    shapes = [("qkv", 3072, 1024, 0, False), ("proj_ls_resid", 1024, 1024, 0, True), ("fc1_gelu", 4096, 1024, 1, False),
              ("fc2_ls_resid", 1024, 4096, 0, True)]
    rows = []
    for name, N, K, act, resid in shapes:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        b = torch.randn(N, device=dev)
        if resid:
            g = torch.rand(N, device=dev)
            x = torch.randn(M, N, device=dev)
            fn = lambda: ops.gemm(a, w, x, bias=b, gamma=g, resid=x)                      # noqa: E731
        else:
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w, o, bias=b, act=act)                               # noqa: E731
        bb = b.bfloat16()
        ref = lambda: torch.nn.functional.linear(a, w, bb)                                # noqa: E731
        ts = []
        for f in (fn, ref):
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                f()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 10)
        fl = 2.0 * M * N * K
        rows.append({"shape": name, "M": M, "N": N, "K": K, "ms": ts[0], "tflops": fl / ts[0] / 1e9,
                     "cublas_plain_ms": ts[1], "cublas_plain_tflops": fl / ts[1] / 1e9,
                     "frac_of_burst_peak": fl / ts[0] / 1e9 / peaks["burst"]})
        del a, w
    tot_fl = sum(2.0 * r["M"] * r["N"] * r["K"] for r in rows)
    tot_ms = sum(r["ms"] for r in rows)
    prof = None
    ppath = os.path.join(ROOT, "profiles", "gemm_tensor_pipe.json")
    if os.path.exists(ppath):
        with open(ppath) as f:
            prof = json.load(f)
    return {"kernel": "gemm_kernel (tcgen05, fused epilogues), the four trunk shapes of one block", "bound": "tensor",
            "achieved": tot_fl / tot_ms / 1e9, "peak": peaks["burst"], "unit": "TFLOP/s",
            "frac": tot_fl / tot_ms / 1e9 / peaks["burst"],
            "peak_source": f"{peaks['source']} (burst figure: kernels timed alone); cuBLAS column = torch linear "
                           f"without the fused GELU / LayerScale / residual",
            "shapes": rows, "ncu_tensor_pipe": prof}


def cfg3_sequence_extra(model, dev, rank, world) -> dict:
    """BASELINE configs[2]: the whole offline chunk creation (frames -> forward -> masks, keypoints, intrinsics ->
    chunk files) on a 1000-frame 640x480-shaped synthetic sequence (target 308x406), chunk 100 / overlap 20 = 13
    chunks, scheduled over the ranks by plan_waves (hybrid: the last waves run sequence-parallel)."""
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig
    from pi3_slam_b200.synthetic import make_frames
    frames, H, W = 1000, 308, 406
    paths = [f"/synthetic/{i:05d}.png" for i in range(frames)]
    index = {p: i for i, p in enumerate(paths)}
    # The sequence as a decoder would hand it over: uint8 frames at the SOURCE resolution (640x480, 7-Scenes-shaped), held
    # in host memory and materialised before the timed region (synthesising 1000 random frames costs more host time than
    # decoding them would and is not part of the workload). The creator copies the bytes to the device and runs the
    # Pillow-exact resize to the 308x406 target + ToTensor there (pi3b_resize_u8_pil), its real input path.
    g = torch.Generator().manual_seed(4321)
    source = torch.randint(0, 256, (frames, 480, 640, 3), dtype=torch.uint8, generator=g)

    def loader(ps, ts):
        i0 = index[ps[0]]
        assert [index[p] for p in ps] == list(range(i0, i0 + len(ps)))
        return source[i0:i0 + len(ps)]
    box = [tempfile.mkdtemp(prefix="pi3b_cfg3_")]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    root = box[0]
    cfg = OfflineCreatorConfig(model_path="unused", output_dir=root, device=str(dev), do_metric_depth=False,
                               keypoint_type="grid", max_num_keypoints=200, chunk_length=100, overlap=20, keypoint_seed=5,
                               sequence_parallel_tail=True)
    cr = OfflineChunkCreator(cfg, model=model)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    devnull = open(os.devnull, "w")
    saved = sys.stdout
    sys.stdout = devnull                          # the creator prints per-chunk progress like the reference
    try:
        warm = paths[:180] if world == 1 else paths   # warm-up: allocations, NCCL groups of the sequence-parallel waves
        cr.process_and_save(warm, loader=loader, target_size=(H, W))
        sync()
        t0 = time.perf_counter()
        cr.process_and_save(paths, loader=loader, target_size=(H, W))
        sync()
        makespan = time.perf_counter() - t0
    finally:
        sys.stdout = saved
        devnull.close()
    st = cr.last_run_stats
    stats = [st]
    if world > 1:
        stats = [None] * world
        dist.all_gather_object(stats, st)
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
    t = torch.tensor([makespan], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    makespan = float(t.item())
    waves = st["waves"]
    wave_done = []
    for wi, wave in enumerate(waves):
        done = [s["chunk_done_s"].get(ci) for s in stats for ci, _ in wave if s["chunk_done_s"].get(ci) is not None]
        wave_done.append(max(done) if done else None)
    processed = sum(min(100, frames - 80 * i) for i in range(13))
    return {"workload": "cfg3: 1000 uint8 640x480 frames in host memory -> device resize to 308x406, chunk 100 / overlap 20 "
                        "(13 chunks, 1240 processed frames), grid keypoints 200, intrinsics solve, chunk files written",
            "n_gpus": world,
            "makespan_s": makespan, "sequence_frames_per_s": frames / makespan,
            "processed_frames_per_s": processed / makespan,
            "waves": [[len(r) for _, r in w] for w in waves], "wave_done_s": wave_done,
            "rank0_host_split_s": {k: st[k] for k in ("wait_frames_s", "enqueue_s", "host_half_s", "total_s")},
            "slowest_rank_host_half_s": max(s["host_half_s"] for s in stats)}


def cfg5_sp_extra(model, dev, rank, world) -> dict:
    """BASELINE configs[4]: one 300-frame 308x406 chunk (192 900 tokens), sequence-parallel over all ranks: per-rank
    frames as even as possible, K|V all-gathered over NVLink in each of the 18 chunk-global layers, overlapped with the
    attention over the local block."""
    import torch
    import torch.distributed as dist
    from pi3_slam_b200.host.datasets import split_frames
    from pi3_slam_b200.pi3 import Pi3
    from pi3_slam_b200.synthetic import make_frames
    frames, H, W = 300, 308, 406
    T = (H // 14) * (W // 14) + 5
    split = split_frames(frames, world)
    first, cnt = split[rank]
    counts = [n for _, n in split]
    fr = make_frames(cnt, H, W, start_index=first)[None].to(dev)
    res = {"workload": "cfg5: one 300-frame chunk at 308x406 (192 900 tokens)", "n_gpus": world, "frames_per_rank": counts}

    def run(steps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model(fr, sp_group=dist.group.WORLD, sp_frame_counts=counts)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    run(1)
    model.sp_events = []
    ms = run(2)
    exposed = [a.elapsed_time(b) for a, b in model.sp_events]
    model.sp_events = None
    model.sp_overlap = False
    ms_blocking = run(1) if len(set(counts)) == 1 else None       # the blocking variant needs an even split
    model.sp_overlap = True
    # the all-gather alone
    m_max = max(counts) * T
    send = torch.empty(m_max, 2048, device=dev, dtype=torch.bfloat16)
    recv = torch.empty(world * m_max, 2048, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        dist.all_gather_into_tensor(recv, send)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dist.all_gather_into_tensor(recv, send)
    e1.record()
    torch.cuda.synchronize()
    ag_ms = e0.elapsed_time(e1) / 5
    fl = Pi3.flops(frames, H, W)
    ex = torch.tensor([sum(exposed) / max(1, len(exposed))], device=dev, dtype=torch.float64)
    ex_min = ex.clone()
    dist.all_reduce(ex, op=dist.ReduceOp.MAX)
    dist.all_reduce(ex_min, op=dist.ReduceOp.MIN)
    layer_ms = (fl["global_attn"] / 18 / world) / 1e9 / 900.0     # a global layer's attention at ~900 TFLOP/s per GPU
    res.update(ms_per_chunk=ms, frames_per_s=frames / (ms * 1e-3),
               tflops_per_gpu=fl["total"] / (ms * 1e-3) / 1e12 / world, ms_per_chunk_blocking_allgather=ms_blocking,
               allgather_ms_per_layer=ag_ms,
               allgather_recv_bytes_per_rank_per_layer=(world - 1) * m_max * 4096,
               allgather_gbs_per_rank=(world - 1) * m_max * 4096 / (ag_ms * 1e-3) / 1e9,
               # time the stream waits for the gathered K/V after the local-block attention. MAX over ranks includes the
               # wait of the ranks that hold one frame less for the others (load imbalance of an uneven split); MIN over
               # ranks (the most loaded rank) is the transfer time the local block really failed to hide
               exposed_allgather_ms_per_layer=float(ex.item()),
               exposed_allgather_ms_per_layer_most_loaded_rank=float(ex_min.item()),
               exposed_share_of_global_layer=float(ex.item()) / (layer_ms + float(ex.item())))
    # correctness on a small shape: blocking variant bit-identical to the single-GPU forward, overlapped within tolerance
    n_small = 2 * world
    small = make_frames(n_small, 56, 84)[None].to(dev)
    ref = {k: v[:, 2 * rank:2 * rank + 2].clone() for k, v in model(small).items()}
    model.sp_overlap = False
    o1 = model(small[:, 2 * rank:2 * rank + 2].contiguous(), sp_group=dist.group.WORLD)
    same = all(torch.equal(o1[k], ref[k]) for k in ref)
    model.sp_overlap = True
    o2 = model(small[:, 2 * rank:2 * rank + 2].contiguous(), sp_group=dist.group.WORLD)
    err = max(float((o2[k].double() - ref[k].double()).norm() / ref[k].double().norm()) for k in ref)
    flag = torch.tensor([1.0 if same else 0.0, -err], device=dev, dtype=torch.float64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["small_shape_blocking_bit_identical_to_1gpu"] = bool(flag[0].item() == 1.0)
    res["small_shape_overlapped_max_rel_l2"] = -float(flag[1].item())
    # ... and at the FULL cfg5 shape: every rank also runs the whole 300-frame chunk alone (it fits one GPU: ~10 GB) and
    # compares its own frames of the sequence-parallel result with that single-GPU forward
    del small, ref, o1, o2
    sp_out = {k: v.clone() for k, v in model(fr, sp_group=dist.group.WORLD, sp_frame_counts=counts).items()}
    whole = make_frames(frames, H, W)[None].to(dev)
    one = model(whole)
    errs = {k: float((sp_out[k].double() - one[k][:, first:first + cnt].double()).norm() /
                     one[k][:, first:first + cnt].double().norm()) for k in sp_out}
    e = torch.tensor([errs[k] for k in sorted(errs)], device=dev, dtype=torch.float64)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    res["full_shape_sp_vs_1gpu_max_rel_l2"] = dict(zip(sorted(errs), e.tolist()))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the GPU-fp32 oracle comparison of the timed step")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.cfg3_sequence / extra.cfg5_sp / roofline.gemm")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: until it is emitted fd 1 points at stderr, so native writers (the NCCL
    # version banner is a printf to stdout, whatever NCCL_DEBUG_FILE says) cannot interleave with it
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj) -> None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(obj), flush=True)

    frames, H, W, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: {desc}", "frames_per_chunk": frames, "height": H, "width": W,
              "weights": "synthetic seed 0 (pi3_slam_b200.synthetic)", "chunks_per_step_per_gpu": 1,
              "sharding": "chunk-wise, one chunk per rank per step, no data-path collective",
              "l2": "per-step working set (235 MB frames, >5 GB activations) exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        import torch
        steps = max(1, min(args.steps, 2))        # ~12 s per 8-frame step: the whole arm stays within a minute
        r = cpu_reference_run(CPU_SAMPLE_FRAMES, H, W, steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": "pi3_chunk_frames_per_sec", "value": r["value"], "unit": "frames/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["s_per_step"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from pi3_slam_b200 import _lib
    from pi3_slam_b200.pi3 import Pi3
    from pi3_slam_b200.synthetic import make_frames, make_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL writes its version banner / debug output to stdout by default
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        from pi3_slam_b200.host.datasets import nccl_high_priority_options
        dist.init_process_group("nccl", device_id=dev, pg_options=nccl_high_priority_options())

    sd = make_state_dict(0)
    model = Pi3()
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    # each rank owns a different chunk of the synthetic sequence (stride 80 = chunk 100 / overlap 20)
    host_frames = make_frames(frames, H, W, start_index=rank * max(1, frames - 20))[None].pin_memory()
    dev_frames = host_frames.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value) ----
    for _ in range(args.warmup):
        model(dev_frames)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    model.global_attn_events = []
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = model(dev_frames)
    e1.record()
    barrier()
    launches = _lib.launch_count - launches0
    ms_total = e0.elapsed_time(e1)
    attn_ms = [a.elapsed_time(b) for a, b in model.global_attn_events]
    model.global_attn_events = None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # ---- end to end through the public API: pinned host frames -> Pi3.forward -> pinned host results ----
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
    h2d = host_frames.numel() * host_frames.element_size()
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())

    # The loop a streaming caller writes (and `OfflineChunkCreator.process_and_save` runs): a copy stream uploads the next
    # step's frames while the current forward computes, and the results leave through a device staging copy (0.2 ms) whose
    # download overlaps the next forward. Every step's H2D and D2H happens inside the timed region; the region ends only
    # after the last step's results have landed in host memory.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [torch.empty_like(dev_frames) for _ in range(2)]
    staged = {k: torch.empty_like(v) for k, v in out.items()}
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    staged_full, staged_free = torch.cuda.Event(), torch.cuda.Event()

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(in_free[i % 2])          # the forward that read this buffer two steps ago is done
            dev_in[i % 2].copy_(host_frames, non_blocking=True)
            in_ready[i % 2].record(copy_stream)

    def e2e_run(steps):
        cur = torch.cuda.current_stream()
        for ev in in_free:
            ev.record(cur)
        staged_free.record(cur)
        upload(0)
        for i in range(steps):
            if i + 1 < steps:
                upload(i + 1)
            cur.wait_event(in_ready[i % 2])
            o = model(dev_in[i % 2])
            in_free[i % 2].record(cur)
            cur.wait_event(staged_free)                     # the previous step's download has left the staging buffers
            for k, v in o.items():
                staged[k].copy_(v, non_blocking=True)
            staged_full.record(cur)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(staged_full)
                for k in staged:
                    host_out[k].copy_(staged[k], non_blocking=True)
                staged_free.record(copy_stream)
        cur.wait_stream(copy_stream)                        # the last download is inside the timed region

    e2e_run(1)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_run(args.steps)
    s1.record()
    barrier()
    t2 = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t2.item()) / args.steps
    finite = all(bool(torch.isfinite(v).all()) for v in host_out.values())

    # ---- outside the timed regions -------------------------------------------------------------------------------
    def guarded(fn, *a):
        try:
            return fn(*a)
        except Exception as e:          # an extra must never cost the headline line
            import traceback
            traceback.print_exc()
            return {"error": f"{type(e).__name__}: {e}"}

    extra = {}
    gemm_entry = None
    if not args.no_extras and args.workload == "cfg2":
        extra["cfg3_sequence"] = guarded(cfg3_sequence_extra, model, dev, rank, world)
        if world > 1:
            extra["cfg5_sp"] = guarded(cfg5_sp_extra, model, dev, rank, world)
        if rank == 0:
            gemm_entry = guarded(gemm_roofline, dev, _peaks())
    parity = None
    if rank == 0 and world == 1 and not args.no_parity:
        parity = guarded(parity_block, model, dev_frames, out, sd)

    if rank == 0:
        fl = Pi3.flops(frames, H, W)
        peaks = _peaks()
        T = (H // 14) * (W // 14) + 5
        attn_flops = 16 * 4.0 * (frames * T) ** 2 * 64          # one chunk-global attention launch
        attn_avg_ms = sum(attn_ms) / max(1, len(attn_ms))
        achieved = attn_flops / (attn_avg_ms * 1e-3) / 1e12 if attn_ms else None
        traffic, traffic_detail = None, None
        tpath = os.path.join(ROOT, "profiles", "attn_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic_detail = json.load(f).get(args.workload)
            if traffic_detail:
                traffic = traffic_detail["bytes_per_launch"]   # dram read + write of one ncu --set full capture
        line = {
            "metric": "pi3_chunk_frames_per_sec", "value": world * frames / (ms_step * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config,
            "e2e": {"value": world * frames / (e2e_ms_step * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_step, "outputs_finite": finite,
                    "pipelined": "H2D of step k+1 and D2H of step k-1 overlap the forward of step k (copy stream)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "model_tflops_per_step": fl["total"] / 1e12,
            "achieved_tflops": world * fl["total"] / (ms_step * 1e-3) / 1e12,
            "frac_of_bf16_peak": {"burst": fl["total"] / (ms_step * 1e-3) / 1e12 / peaks["burst"],
                                  "sustained": fl["total"] / (ms_step * 1e-3) / 1e12 / peaks["sustained"],
                                  "note": "per-GPU achieved / per-GPU measured peak"},
            "roofline": {"kernel": "attn_fwd_kernel_t3 (chunk-global attention, 18 launches/step)", "bound": "tensor",
                         "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["sustained"]) if achieved else None,
                         "frac_of_burst_peak": (achieved / peaks["burst"]) if achieved else None,
                         "peak_source": f"{peaks['source']} (sustained figure: kernel timed inside a long step)",
                         "flops_per_launch": attn_flops, "avg_launch_ms": attn_avg_ms, "launches_timed": len(attn_ms),
                         "share_of_step": (sum(attn_ms) / args.steps) / ms_step if attn_ms else None,
                         "traffic": traffic, "traffic_detail": traffic_detail},
        }
        line["parity"] = parity
        if gemm_entry is not None:
            line["roofline"]["gemm"] = gemm_entry
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(CPU_SAMPLE_FRAMES, H, W, 1, 0, sd)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["same_work"] = guarded(same_work_pair, model, sd, dev)
            if "ratio" in line["cpu_baseline"]["same_work"]:
                line["cpu_baseline"]["same_work_ratio"] = line["cpu_baseline"]["same_work"]["ratio"]
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
