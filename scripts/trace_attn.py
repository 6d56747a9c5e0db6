"""Debug: per-iteration timeline of CTA 0 of the attention kernel (needs a -DPI3B_TRACE build via PI3B_LIB)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pi3_slam_b200 import _lib, ops
lib = _lib.load()
lib.pi3b_debug_set_trace.argtypes = [ctypes.c_void_p]
S, H = int(sys.argv[1]), 16
qkv = torch.randn(S, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty(S, H * 64, device="cuda", dtype=torch.bfloat16)
tr = torch.zeros(3 * 64 * 8, device="cuda", dtype=torch.int64)
ops.attention_qkv(qkv, out, 1, S, H)
torch.cuda.synchronize()
assert lib.pi3b_debug_set_trace(tr.data_ptr()) == 0
ops.attention_qkv(qkv, out, 1, S, H)
torch.cuda.synchronize()
t = tr.cpu().view(3, 64, 8)
base = int(t[t > 0].min())
def rel(x): return int(x) - base if int(x) > 0 else -1
print("MMA warp: it | s_free0 s_free1 QK0iss QK1iss p_full0 p_full1 iter_end")
for j in range(8, 24): print(j, [rel(x) for x in t[0, j, :7]])
for role in (1, 2):
    print(f"softmax tile {role-1}: it | start S_ready S_in_regs exp_done pv_done P_stored ; step length")
    for j in range(8, 24):
        r = [rel(x) for x in t[role, j, :6]]
        print(j, r, "| wait_S", r[1]-r[0], "ld", r[2]-r[1], "math", r[3]-r[2], "wait_pv", r[4]-r[3], "st", r[5]-r[4], "| step", int(t[role, j+1, 0]) - int(t[role, j, 0]))
