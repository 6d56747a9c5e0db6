"""Where does the proj + LayerScale + residual GEMM (M = 100 400, N = K = 1024) spend its time? Same shape with the
residual path off / fp32 out / bf16 out, and cuBLAS beside it."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pi3_slam_b200 import ops  # noqa: E402
from bench_kernels import time_cuda  # noqa: E402

dev = "cuda"
M = 100400
for N, K in [(1024, 1024), (1024, 4096)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.randn(N, device=dev)
    g = torch.rand(N, device=dev)
    x = torch.randn(M, N, device=dev)
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    of = torch.empty(M, N, device=dev)
    cases = {
        "bf16 out, bias": lambda: ops.gemm(a, w, ob, bias=b),
        "fp32 out, bias+gamma": lambda: ops.gemm(a, w, of, bias=b, gamma=g),
        "fp32 resid RMW in place": lambda: ops.gemm(a, w, x, bias=b, gamma=g, resid=x),
        "fp32 resid -> separate fp32 out": lambda: ops.gemm(a, w, of, bias=b, gamma=g, resid=x),
        "cuBLAS bf16": lambda: torch.nn.functional.linear(a, w, b.bfloat16()),
    }
    for name, f in cases.items():
        t = time_cuda(f, iters=20, warmup=3)
        print(f"N={N} K={K} {name:34s} {t * 1e3:7.4f} ms {2.0 * M * N * K / t / 1e12:7.1f} TF/s", flush=True)
