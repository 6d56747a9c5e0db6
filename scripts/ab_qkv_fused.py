"""A/B of the QKV projection's q/k post-processing (same box, CUDA events): pi3b_gemm + pi3b_qknorm_rope (two launches,
822 MB read-modify-write at M = 100 400) vs pi3b_gemm_qkv (fused epilogue), next to the reference's own native kernel
rope_2d_cuda_kernel (pi3/models/curope/kernels.cu:17-108, built for sm_100a by oracle/build_ref_curope.py) doing ONLY the
RoPE part on the same q and k. Writes gpurun_out/ab_qkv_fused.json."""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_kernels import time_cuda  # noqa: E402
from pi3_slam_b200 import ops  # noqa: E402

H, K = 16, 1024
res = []
ref = None
so = os.path.join(ROOT, "oracle", "_ref", "ref_curope.so")
if os.path.exists(so):
    spec = importlib.util.spec_from_file_location("ref_curope", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
for frames, T, pw in ((100, 1004, 37), (100, 643, 29)):
    M = frames * T
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(3 * 64 * H, K, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(3 * 64 * H, device="cuda") * 0.1
    qn = (torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1)
    kn = (torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1)
    qkn = torch.cat([qn[0], qn[1], kn[0], kn[1]]).contiguous()
    inv_freq = 1.0 / (100.0 ** (torch.arange(0, 32, 2).float() / 32))
    ang = torch.einsum("i,j->ij", torch.arange(pw + 2, dtype=torch.float32), inv_freq)
    cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().cuda()
    out = torch.empty(M, 3 * 64 * H, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * 3072 * K
    t_gemm = time_cuda(lambda: ops.gemm(a, w, out, bias=bias))
    t_pass = time_cuda(lambda: ops.qknorm_rope(out, H, T, 5, pw, cs, qn, kn, 1e-5))
    t_rope_only = time_cuda(lambda: ops.qknorm_rope(out, H, T, 5, pw, cs, None, None, 1e-5))
    t_fused = time_cuda(lambda: ops.gemm_qkv(a, w, out, bias, H, qkn, 1e-5, cs, T, 5, pw))
    t_fused_rope = time_cuda(lambda: ops.gemm_qkv(a, w, out, bias, H, None, 1e-5, cs, T, 5, pw))
    r = dict(M=M, gemm_ms=t_gemm * 1e3, gemm_tflops=fl / t_gemm / 1e12, qknorm_rope_pass_ms=t_pass * 1e3,
             pass_gbs=2 * (2 / 3) * M * 3072 * 2 / t_pass / 1e9, rope_only_pass_ms=t_rope_only * 1e3,
             two_launch_total_ms=(t_gemm + t_pass) * 1e3, fused_ms=t_fused * 1e3, fused_tflops=fl / t_fused / 1e12,
             fused_rope_only_ms=t_fused_rope * 1e3, extra_over_plain_gemm_ms=(t_fused - t_gemm) * 1e3)
    if ref is not None:
        # the reference applies rope_2d to q and to k, each a contiguous (B, N, H, D) tensor
        tok = torch.randn(1, M, H, 64, device="cuda").bfloat16()
        t_idx = torch.arange(M, device="cuda") % T
        p = (t_idx - 5).clamp(min=0)
        pos = torch.stack([torch.where(t_idx >= 5, p // pw + 1, 0), torch.where(t_idx >= 5, p % pw + 1, 0)], -1)[None].contiguous()
        t_ref = time_cuda(lambda: ref.rope_2d(tok, pos, 100.0, 1.0))
        r["reference_curope_rope2d_q_and_k_ms"] = 2 * t_ref * 1e3
        r["reference_curope_gbs"] = 2 * M * 1024 * 2 / t_ref / 1e9
    res.append(r)
    print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_qkv_fused.json"), "w"), indent=1)
