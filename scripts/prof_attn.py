"""Tiny driver for ncu: a few launches of the attention kernel at one shape."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pi3_slam_b200 import ops
B, S, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty(B * S, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_qkv(qkv, out, B, S, H)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
