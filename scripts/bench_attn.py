import torch
from pi3_slam_b200 import ops
from bench_kernels import time_cuda


def bench_attn(results):
    for (B, S, H, iters) in [(100, 1004, 16, 5), (100, 643, 16, 5), (1, 20080, 16, 3), (1, 64300, 16, 4), (1, 100400, 16, 4)]:
        qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
        out = torch.empty(B * S, H * 64, device="cuda", dtype=torch.bfloat16)
        t = time_cuda(lambda: ops.attention_qkv(qkv, out, B, S, H), iters=iters, warmup=1)
        fl = 4.0 * B * H * S * S * 64
        r = dict(kernel="attn_hd64", B=B, S=S, H=H, ms=t * 1e3, tflops=fl / t / 1e12)
        if S <= 20080:
            q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
            with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.FLASH_ATTENTION):
                t2 = time_cuda(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters=iters, warmup=1)
            r["torch_fa2_ms"] = t2 * 1e3
            r["torch_fa2_tflops"] = fl / t2 / 1e12
        results.append(r)
        print(r, flush=True)
