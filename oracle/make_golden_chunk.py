"""Generate tests/golden/chunk_tiny.pt from the REAL reference `OfflineChunkCreator._process_single_chunk`
(run in the build container only; needs /root/reference).

Third-party modules the reference imports but that are absent here (pytheia, torchcodec, plyfile, natsort,
matplotlib) are stubbed with MagicMock — none of them is touched by `_process_single_chunk`; `utils3d` gets a
two-line stand-in for `intrinsics_from_focal_center` (it only assembles K from fx, fy, cx, cy). The creator
object is built with object.__new__ because its __init__ downloads weights (offline_chunk_creator.py:65).
MoGe is disabled (its checkpoint config is not in the repo), so the metric-scale branch is not golden-pinned.
"""
import os
import sys
import types
from unittest.mock import MagicMock

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

for name in ("torchcodec", "torchcodec.decoders", "plyfile", "natsort", "pytheia", "matplotlib", "matplotlib.pyplot",
             "matplotlib.animation"):
    sys.modules.setdefault(name, MagicMock())
u3 = types.ModuleType("utils3d")
u3.torch = types.ModuleType("utils3d.torch")


def _intrinsics_from_focal_center(fx, fy, cx, cy):
    K = torch.zeros(*fx.shape, 3, 3, dtype=fx.dtype, device=fx.device)
    K[..., 0, 0], K[..., 1, 1], K[..., 0, 2], K[..., 1, 2], K[..., 2, 2] = fx, fy, cx, cy, 1.0
    return K


u3.torch.intrinsics_from_focal_center = _intrinsics_from_focal_center
sys.modules["utils3d"] = u3
sys.modules["utils3d.torch"] = u3.torch

from pi3_slam_b200.synthetic import make_frames, make_state_dict  # noqa: E402

N, H, W, MAX_KP, SEED = 4, 84, 112, 30, 4321


def main():
    from pi3.models.pi3 import Pi3
    from slam.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig
    from utils.keypoint_extraction import create_keypoint_extractor

    model = Pi3().eval()
    model.load_state_dict(make_state_dict(0), strict=True)
    cfg = OfflineCreatorConfig(model_path="unused", output_dir="/tmp/unused", device="cpu", do_metric_depth=False,
                               keypoint_type="grid", max_num_keypoints=MAX_KP)
    creator = object.__new__(OfflineChunkCreator)
    creator.config = cfg
    creator.model = model
    creator.moge_model = None
    creator.keypoint_extractor = create_keypoint_extractor(keypoint_type="grid", max_num_keypoints=MAX_KP, detection_threshold=0.005, device="cpu")
    creator.target_size = (H, W)
    creator.undistortion_maps = None

    imgs = make_frames(N, H, W)[None]
    paths = [(f"frame_{i:04d}.png",) for i in range(N)]
    torch.manual_seed(SEED)
    with torch.no_grad():
        dense = model(imgs)                      # dense reference outputs for the restatement check
    torch.manual_seed(SEED)
    out = creator._process_single_chunk(imgs, paths)
    assert "camera_params" in out and "keypoints" in out, list(out.keys())
    gold = {"N": N, "H": H, "W": W, "max_kp": MAX_KP, "seed": SEED, "chunk": out,
            "dense": {k: v[0].clone() for k, v in dense.items()}}
    path = os.path.join(ROOT, "tests", "golden", "chunk_tiny.pt")
    torch.save(gold, path)
    for k, v in out.items():
        print(k, (tuple(v.shape), v.dtype) if torch.is_tensor(v) else type(v).__name__)
    print("masks true fraction:", out["masks"].float().mean().item(), "-> saved", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
