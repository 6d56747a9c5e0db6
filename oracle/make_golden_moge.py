"""Generate tests/golden/moge_*.pt from the REAL reference `moge.model.v2.MoGeModel` (run in the build container
only; needs /root/reference) and check oracle/moge_oracle.py against it.

`utils3d` (un-vendored, unpinned) gets a two-function stand-in: `intrinsics_from_focal_center` only assembles K, and
`depth_to_points` only affects the `points` output, which the chunk path never reads (it uses ["depth"],
slam/offline_chunk_creator.py:184). The model_config is the recorded synthetic one (pi3_slam_b200.synthetic
.moge_synth_config) because the reference reads it from a hub checkpoint that is not in the repo (v2.py:89-94).

Cases: "up" = 84x112 frame at 120 base tokens (9x12 -> 126x168 ViT input: the resize up-samples, as in the real
configs where 378x518 goes to 714x980); "down" = 140x182 frame at 48 tokens (6x7 -> 84x98: antialias active).
"""
import os
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

u3 = types.ModuleType("utils3d")
u3.torch = types.ModuleType("utils3d.torch")


def _intrinsics_from_focal_center(fx, fy, cx, cy):
    fx = torch.as_tensor(fx)
    K = torch.zeros(*fx.shape, 3, 3, dtype=fx.dtype)
    K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = fx, fy, cx, cy, 1.0
    return K


u3.torch.intrinsics_from_focal_center = _intrinsics_from_focal_center
u3.torch.depth_to_points = lambda depth, intrinsics=None: torch.zeros(*depth.shape, 3)
sys.modules["utils3d"] = u3
sys.modules["utils3d.torch"] = u3.torch

from oracle import moge_oracle  # noqa: E402
from pi3_slam_b200.synthetic import make_frames, make_moge_state_dict, moge_synth_config  # noqa: E402

CASES = {"up": dict(H=84, W=112, num_tokens=120, frame=0), "down": dict(H=140, W=182, num_tokens=48, frame=3)}


def main():
    from moge.model.v2 import MoGeModel
    cfg = moge_synth_config()
    sd = make_moge_state_dict(cfg, 0)
    model = MoGeModel(**cfg).eval()
    ref_keys = model.state_dict()
    assert set(ref_keys) == set(sd) and all(tuple(ref_keys[k].shape) == tuple(sd[k].shape) for k in sd)
    model.load_state_dict(sd, strict=True)
    for name, c in CASES.items():
        img = make_frames(1, c["H"], c["W"], start_index=c["frame"])[0]
        with torch.no_grad():
            fwd = model.forward(img[None], num_tokens=c["num_tokens"])
        inf = model.infer(img, num_tokens=c["num_tokens"], use_fp16=False)
        mine = moge_oracle.moge_infer(sd, cfg, img, num_tokens=c["num_tokens"])
        errs = {
            "points": float((mine["fwd_points"] - fwd["points"][0]).abs().max()),
            "mask": float((mine["fwd_mask"] - fwd["mask"][0]).abs().max()),
            "metric_scale": float((mine["metric_scale"] - fwd["metric_scale"][0]).abs().max()),
            "mask_binary_mismatch": int((mine["mask"] != inf["mask"]).sum()),
            "intrinsics": float((mine["intrinsics"] - inf["intrinsics"]).abs().max()),
        }
        both = torch.isfinite(inf["depth"]) & torch.isfinite(mine["depth"])
        errs["depth"] = float((mine["depth"][both] - inf["depth"][both]).abs().max())
        errs["depth_inf_mismatch"] = int((torch.isfinite(inf["depth"]) != torch.isfinite(mine["depth"])).sum())
        print(name, errs, "valid frac", float(inf["mask"].float().mean()))
        assert errs["points"] < 2e-4 and errs["mask"] < 1e-5 and errs["metric_scale"] < 1e-5, errs
        assert errs["depth"] < 2e-4 and errs["depth_inf_mismatch"] == 0 and errs["mask_binary_mismatch"] == 0, errs
        out = {
            "case": dict(c), "model_config": cfg, "seed": 0,
            "fwd_points": fwd["points"][0].clone(), "fwd_mask": fwd["mask"][0].clone(),
            "metric_scale": fwd["metric_scale"][0].clone(),
            "depth": inf["depth"].clone(), "mask": inf["mask"].clone(), "intrinsics": inf["intrinsics"].clone(),
        }
        path = os.path.join(ROOT, "tests", "golden", f"moge_{name}.pt")
        torch.save(out, path)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
