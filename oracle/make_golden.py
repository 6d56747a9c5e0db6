"""Generate tests/golden/pi3_forward_*.pt from the REAL reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference

Loads the synthetic state_dict (pi3_slam_b200.synthetic.make_state_dict(0)) into the unmodified
reference `Pi3` (strict=True — proves the key/shape layout), runs the reference forward in fp32 on CPU,
checks the oracle restatement against it, and stores the reference's outputs as golden vectors.
"""
import os
import sys
import time

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

from oracle import pi3_oracle  # noqa: E402
from pi3_slam_b200.synthetic import make_frames, make_state_dict, pi3_param_shapes  # noqa: E402

CASES = {
    # name: (N, H, W, spatial stride used when storing the dense maps)
    "tiny": (3, 42, 56, 1),
    "cfg1": (8, 224, 224, 8),
}


def main():
    from pi3.models.pi3 import Pi3  # the reference

    torch.manual_seed(0)
    t0 = time.time()
    ref = Pi3().eval()
    ref_sd = ref.state_dict()
    shapes = pi3_param_shapes()
    assert set(ref_sd.keys()) == set(shapes.keys()), "state_dict keys differ from the reference"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, v.shape, shapes[k])
    sd = make_state_dict(0)
    ref.load_state_dict(sd, strict=True)
    print(f"reference built + synthetic weights loaded in {time.time() - t0:.1f}s")

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (N, H, W, stride) in CASES.items():
        imgs = make_frames(N, H, W)[None]
        t0 = time.time()
        with torch.no_grad():
            r = ref(imgs)
        t_ref = time.time() - t0
        t0 = time.time()
        o = pi3_oracle.pi3_forward(imgs, sd)
        t_or = time.time() - t0
        for k in ("points", "local_points", "conf", "camera_poses"):
            d = (r[k] - o[k]).abs().max().item()
            print(f"[{name}] {k}: reference vs oracle max abs diff {d:.3e} (ref max {r[k].abs().max().item():.3f})")
            assert d < 1e-4 * max(1.0, r[k].abs().max().item()), f"oracle deviates from the reference on {k}"
        gold = {
            "N": N, "H": H, "W": W, "stride": stride, "seed": 0,
            "points": r["points"][:, :, ::stride, ::stride].contiguous(),
            "local_points": r["local_points"][:, :, ::stride, ::stride].contiguous(),
            "conf": r["conf"][:, :, ::stride, ::stride].contiguous(),
            "camera_poses": r["camera_poses"].contiguous(),
            "ref_seconds": t_ref, "oracle_seconds": t_or, "threads": torch.get_num_threads(),
        }
        torch.save(gold, os.path.join(out_dir, f"pi3_forward_{name}.pt"))
        print(f"[{name}] reference {t_ref:.1f}s, oracle {t_or:.1f}s -> tests/golden/pi3_forward_{name}.pt")


if __name__ == "__main__":
    main()
