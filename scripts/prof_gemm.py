"""Driver for ncu: one launch of each trunk GEMM shape of the cfg2 chunk (qkv, proj + LayerScale + residual, fc1 + GELU,
fc2 + LayerScale + residual at M = 100 400) through pi3b_gemm, followed by the cuBLAS (torch linear) kernels of the same
shapes for comparison. `python scripts/prof_gemm.py [reps]`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pi3_slam_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M = 100400
dev = "cuda"
ours, ref = [], []
for name, N, K, act, resid in [("qkv", 3072, 1024, 0, False), ("proj", 1024, 1024, 0, True), ("fc1", 4096, 1024, 1, False),
                               ("fc2", 1024, 4096, 0, True)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.randn(N, device=dev)
    bb = b.bfloat16()
    if resid:
        g = torch.rand(N, device=dev)
        x = torch.randn(M, N, device=dev)
        ours.append(lambda a=a, w=w, b=b, g=g, x=x: ops.gemm(a, w, x, bias=b, gamma=g, resid=x))
    else:
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ours.append(lambda a=a, w=w, b=b, o=o, act=act: ops.gemm(a, w, o, bias=b, act=act))
    ref.append(lambda a=a, w=w, bb=bb: torch.nn.functional.linear(a, w, bb))
# the decoder's fused QKV projection (q/k LayerNorm + RoPE in the epilogue)
T, pw = 1004, 37
qkn = torch.cat([torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1, torch.rand(64, device=dev) + 0.5,
                 torch.randn(64, device=dev) * 0.1]).contiguous()
inv_freq = 1.0 / (100.0 ** (torch.arange(0, 32, 2).float() / 32))
ang = torch.einsum("i,j->ij", torch.arange(pw + 2, dtype=torch.float32), inv_freq)
cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().to(dev)
a_q = torch.randn(M, 1024, device=dev).bfloat16()
w_q = (torch.randn(3072, 1024, device=dev) * 0.02).bfloat16()
b_q = torch.randn(3072, device=dev)
o_q = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
ours.append(lambda: ops.gemm_qkv(a_q, w_q, o_q, b_q, 16, qkn, 1e-5, cs, T, 5, pw))
for _ in range(reps):
    for f in ours:
        f()
    for f in ref:
        f()
torch.cuda.synchronize()
print("ok")
