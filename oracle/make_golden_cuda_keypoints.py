"""Generate tests/golden/keypoints_cuda.pt ON A GPU BOX: what the reference's keypoint path produces when it runs with
its default `device='cuda'` — the CUDA-generator `randperm` order (utils/keypoint_extraction.py:140-143), the CUDA
`arange` lattice (:125-126), and the colours / sampled maps that follow from a grid evaluated on the CUDA device and
a CPU `grid_sample` (:214-229, slam/offline_chunk_creator.py:139-152).

/root/reference does not exist on the GPU box, so this runs oracle.chunk_oracle — the restatement that
oracle/make_golden_chunk.py pins bit-exactly to the real reference on the CPU — with device='cuda': the same torch call
sequence, the other device. Usage (from the repo root, under gpurun):
    python oracle/make_golden_cuda_keypoints.py gpurun_out/keypoints_cuda.pt
then copy the file to tests/golden/. tests/test_creator_gpu.py compares the product path with it bit for bit."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import chunk_oracle as co  # noqa: E402
from pi3_slam_b200.synthetic import make_frames  # noqa: E402

CASES = [dict(name="tiny", N=3, H=56, W=84, max_kp=25, seed=11),
         dict(name="cfg3", N=4, H=308, W=406, max_kp=200, seed=3),
         dict(name="cfg4", N=2, H=280, W=448, max_kp=200, seed=5)]


def frames_u8_blocks(N, H, W):
    """Frames as an 8-bit image decodes to (k / 255), flat over 8x8 blocks: inside a block all four bilinear corners are
    equal, the interpolated value is k/255 times a weight sum that is 1 only up to rounding, and the x255 -> uint8
    truncation lands on k or k-1 depending on the last ulp — the case where only the exact op order is bit-exact."""
    small = make_frames(N, (H + 7) // 8, (W + 7) // 8)
    big = small.repeat_interleave(8, dim=2).repeat_interleave(8, dim=3)[:, :, :H, :W]
    return ((big * 255).floor() / 255).contiguous()


def dense_maps(N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(N, H, W, 3, generator=g)
    lp = torch.randn(N, H, W, 3, generator=g)
    conf = torch.randn(N, H, W, 1, generator=g) * 3
    masks = torch.rand(N, H, W, generator=g) > 0.3
    return pts, lp, conf, masks


def main(out_path):
    assert torch.cuda.is_available()
    out = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "cases": {}}
    for c in CASES:
        torch.manual_seed(c["seed"])
        kps = co.grid_keypoints(c["N"], c["H"], c["W"], c["max_kp"], device="cuda")
        imgs = frames_u8_blocks(c["N"], c["H"], c["W"])
        pts, lp, conf, masks = dense_maps(c["N"], c["H"], c["W"], c["seed"])
        samp = co.interpolate_at_keypoints(pts, lp, conf, masks, kps, c["H"], c["W"])
        out["cases"][c["name"]] = dict(c, keypoints=kps.cpu(), colors=co.colors_at_keypoints(imgs, kps),
                                       points=samp["points"], local_points=samp["local_points"], conf=samp["conf"],
                                       masks=samp["masks"])
        print(c["name"], "keypoints", tuple(kps.shape), "first", kps[0, :3].tolist())
    torch.save(out, out_path)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "keypoints_cuda.pt"))
