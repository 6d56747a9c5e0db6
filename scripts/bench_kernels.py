"""Micro-benchmarks of individual C-ABI kernels against their roofline (run on the GPU box)."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pi3_slam_b200 import ops  # noqa: E402


def time_cuda(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def bench_gemm(results):
    dev = "cuda"
    shapes = [(100400, 3072, 1024, 0), (100400, 1024, 1024, 0), (100400, 4096, 1024, 1), (100400, 1024, 4096, 0),
              (64300, 3072, 1024, 0), (64300, 4096, 1024, 1), (2088, 4096, 1024, 1)]
    for M, N, K, act in shapes:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t = time_cuda(lambda: ops.gemm(a, w, out, bias=b, act=act))
        t_ref = time_cuda(lambda: torch.nn.functional.linear(a, w, b.bfloat16()))
        fl = 2.0 * M * N * K
        results.append(dict(kernel="gemm_bf16", M=M, N=N, K=K, act=act, ms=t * 1e3, tflops=fl / t / 1e12,
                            cublas_ms=t_ref * 1e3, cublas_tflops=fl / t_ref / 1e12))
        print(results[-1], flush=True)
    # residual epilogue
    M, N, K = 100400, 1024, 4096
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.randn(N, device=dev)
    g = torch.rand(N, device=dev)
    x = torch.randn(M, N, device=dev)
    t = time_cuda(lambda: ops.gemm(a, w, x, bias=b, gamma=g, resid=x))
    results.append(dict(kernel="gemm_bf16_ls_resid", M=M, N=N, K=K, ms=t * 1e3, tflops=2.0 * M * N * K / t / 1e12))
    print(results[-1], flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="gemm")
    ap.add_argument("--out", default="gpurun_out/bench_kernels.json")
    args = ap.parse_args()
    res = []
    which = args.which.split(",")
    if "gemm" in which:
        bench_gemm(res)
    if "attn" in which:
        from bench_attn import bench_attn  # type: ignore
        bench_attn(res)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
