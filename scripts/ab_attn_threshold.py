"""Where should pi3b_attn_fwd_hd64 switch from the four-tile kernel (online softmax, P in place) to the three-tile kernel
(bounded-reference softmax)? Times both kernels (PI3B_ATTN_TILES forces one; read once per process, hence subprocesses)
at constant work (B * S^2 fixed) over a range of sequence lengths. Writes gpurun_out/ab_attn_threshold.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/scripts')
from pi3_slam_b200 import ops
from bench_kernels import time_cuda
res = {}
for S in (643, 1004, 2048, 3571, 4096, 6144, 8192, 12288, 16384, 32768, 64300):
    B = max(1, int(round(1004 * 1004 * 100 / (S * S))))
    H = 16
    qkv = torch.randn(B * S, 3 * H * 64, device='cuda').bfloat16(); out = torch.empty(B * S, H * 64, device='cuda', dtype=torch.bfloat16)
    t = time_cuda(lambda: ops.attention_qkv(qkv, out, B, S, H), iters=8, warmup=2)
    res[S] = (B, round(t * 1e3, 4), round(4.0 * B * H * S * S * 64 / t / 1e12, 1))
print('RES', json.dumps(res))
"""
out = {}
for tiles in (3, 4):
    r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT)], capture_output=True, text=True,
                       env=dict(os.environ, PI3B_ATTN_TILES=str(tiles)), timeout=300)
    out[tiles] = json.loads(r.stdout.split("RES", 1)[1]) if "RES" in r.stdout else {"error": r.stderr[-500:]}
    print(tiles, out[tiles], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_attn_threshold.json"), "w"), indent=1)
