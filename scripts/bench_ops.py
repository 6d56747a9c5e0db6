import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pi3_slam_b200 import ops
from bench_kernels import time_cuda
M, H = 100400, 16
qkv = torch.randn(M, 3 * H * 64, device="cuda").bfloat16()
qn = (torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"))
ang = torch.rand(40, 16, device="cuda")
cs = torch.stack((ang.cos(), ang.sin()), -1).contiguous()
t = time_cuda(lambda: ops.qknorm_rope(qkv, H, 1004, 5, 37, cs, qn, qn), iters=20)
print(f"qknorm_rope M={M}: {t*1e3:.3f} ms  {2*2*M*1024*2/t/1e12:.2f} TB/s (RMW of q,k)")
x = torch.randn(M, 1024, device="cuda"); g = torch.ones(1024, device="cuda"); b = torch.zeros(1024, device="cuda")
o = torch.empty(M, 1024, device="cuda", dtype=torch.bfloat16)
t = time_cuda(lambda: ops.layernorm(x, g, b, 1e-6, o), iters=20)
print(f"layernorm M={M}: {t*1e3:.3f} ms  {M*1024*6/t/1e12:.2f} TB/s")
