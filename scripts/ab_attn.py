#!/usr/bin/env python
"""A/B of the attention-kernel variants under pi3_slam_b200/build/variants (scripts/build_attn_variants.sh), same box,
one subprocess per variant (PI3B_LIB selects the library): time at the frame-local and the two chunk-global shapes (and
one sequence-parallel shape), and check the S = 100 400 result against the fp64 sampled-row reference of
tests/test_scale_parity_gpu.py. Writes gpurun_out/ab_attn.log."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests'); sys.path.insert(0, %(root)r + '/scripts')
from pi3_slam_b200 import ops
from bench_kernels import time_cuda
res = {}
for (B, S, H, it) in [(100, 1004, 16, 10), (1, 64300, 16, 4), (1, 100400, 16, 5)]:
    qkv = torch.randn(B * S, 3 * H * 64, device='cuda').bfloat16(); out = torch.empty(B * S, H * 64, device='cuda', dtype=torch.bfloat16)
    t = time_cuda(lambda: ops.attention_qkv(qkv, out, B, S, H), iters=it, warmup=2)
    res[f'{B}x{S}'] = (round(t * 1e3, 3), round(4.0 * B * H * S * S * 64 / t / 1e12, 1))
# queries pre-multiplied by scale * log2(e) (what pi3b_gemm_qkv's q_scale emits), scale = ln 2: the zero-reference softmax
import math
C = H * 64
qs = qkv.clone(); qs[:, :C] = (qs[:, :C].float() * (0.125 * 1.4426950408889634)).bfloat16()
t = time_cuda(lambda: ops.attention_qkv(qs, out, B, S, H, scale=math.log(2.0)), iters=it, warmup=2)
res['prescaled_1x100400'] = (round(t * 1e3, 3), round(4.0 * B * H * S * S * 64 / t / 1e12, 1))
rows = torch.arange(0, S, 997, device='cuda')
errs = []
for h in (0, 7, 15):
    qh = qs[rows, h * 64:(h + 1) * 64].double(); kh = qs[:, C + h * 64:C + (h + 1) * 64].double(); vh = qs[:, 2 * C + h * 64:2 * C + (h + 1) * 64].double()
    ref = torch.softmax(qh @ kh.t() * math.log(2.0), dim=-1) @ vh
    errs.append(float((out[rows, h * 64:(h + 1) * 64].double() - ref).norm() / ref.norm()))
res['prescaled_err'] = float('%%.2e' %% max(errs))
del qs
# sequence-parallel shape: 1/8 of the cfg2 queries against all keys
S, Q, H = 100400, 12550, 16
q = qkv[:Q]; kv = qkv[:, H * 64:].contiguous(); out = torch.empty(Q, H * 64, device='cuda', dtype=torch.bfloat16)
t = time_cuda(lambda: ops.attention_q_kv(q, kv, out, 1, Q, S, H), iters=5, warmup=2)
res['sp_12550x100400'] = (round(t * 1e3, 3), round(4.0 * H * Q * S * 64 / t / 1e12, 1))
del qkv, kv, out, q
if %(check)d:
    from test_scale_parity_gpu import check_attention
    for mode in ('spread', 'rising'):
        r = check_attention(100400, mode)
        res['err_' + mode] = (float('%%.2e' %% r['rel_l2']), float('%%.2e' %% r['worst_row']))
print('RES', json.dumps(res))
"""


def main():
    reps = int(os.environ.get("REPS", "2"))
    only = os.environ.get("ONLY")
    libs = sorted(glob.glob(os.path.join(ROOT, "pi3_slam_b200/build/variants/*.so")))
    if only:
        libs = [l for l in libs if os.path.basename(l)[:-3] in only.split(",")]
    # ENVS="name:K=V,K2=V2;name2:..." runs each library once more per named environment (e.g. PI3B_ATTN_TILES=4)
    envsets = [("", {})]
    for spec in filter(None, os.environ.get("ENVS", "").split(";")):
        nm, kv = spec.split(":", 1)
        envsets.append(("@" + nm, dict(x.split("=", 1) for x in kv.split(","))))
    libs = [(l, sfx, ev) for l in libs for sfx, ev in envsets]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_attn.log"), "a") as log:
        for rep in range(reps):
            for lib, sfx, extra_env in libs:
                env = dict(os.environ, PI3B_LIB=lib, **extra_env)
                code = CHILD % dict(root=ROOT, check=1 if rep == 0 else 0)
                try:
                    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=180)
                    line = r.stdout.split("RES", 1)[1].strip() if "RES" in r.stdout else "FAILED " + r.stderr[-400:]
                except subprocess.TimeoutExpired:
                    line = "TIMEOUT"
                msg = f"rep {rep} {os.path.basename(lib)[:-3] + sfx:18s} {line}"
                print(msg, flush=True)
                log.write(msg + "\n"); log.flush()


if __name__ == "__main__":
    main()
