"""Time the native MoGe metric-depth pass (row a15) at the chunk path's size: one 378x518 frame, 3600 base tokens."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from pi3_slam_b200 import _lib  # noqa: E402
from pi3_slam_b200.moge import MoGeModel  # noqa: E402
from pi3_slam_b200.synthetic import make_frames, make_moge_state_dict, moge_synth_config  # noqa: E402

cfg = moge_synth_config()
model = MoGeModel(**cfg)
model.load_state_dict(make_moge_state_dict(cfg, 0))
model = model.cuda().eval()
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (378, 518)
img = make_frames(1, H, W)[0].cuda()
for _ in range(2):
    model.infer(img)
torch.cuda.synchronize()
n0 = _lib.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
out = model.forward(img[None], 3600)
e1.record()
torch.cuda.synchronize()
fwd_ms = e0.elapsed_time(e1)
launches = _lib.launch_count - n0
t0 = time.perf_counter()
for _ in range(3):
    d = model.infer(img)["depth"]
torch.cuda.synchronize()
infer_ms = (time.perf_counter() - t0) / 3 * 1e3
print(json.dumps({"frame": [H, W], "forward_device_ms": round(fwd_ms, 3), "infer_wall_ms": round(infer_ms, 3),
                  "kernel_launches_forward": launches, "finite_frac": float(torch.isfinite(d).float().mean()),
                  "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))
