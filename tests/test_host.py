"""CPU: host-side logic of the product package (no CUDA calls)."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_chunk_index_list_matches_reference_cases():
    from pi3_slam_b200.host.datasets import chunk_index_list
    idx = chunk_index_list(1000, 100, 20)                      # 13 chunks
    assert len(idx) == 13 and idx[0] == (0, 100) and idx[1] == (80, 180) and idx[-2] == (880, 980) and idx[-1] == (960, 1000)
    assert chunk_index_list(8, 8, 2) == [(0, 8), (6, 8)]       # the CLI run in SURVEY 8c
    assert chunk_index_list(1, 8, 2) == []                      # < 2 frames: dropped
    assert chunk_index_list(9, 8, 0) == [(0, 8)]                # 1-frame tail dropped
    with pytest.raises(ValueError):
        chunk_index_list(10, 4, 4)


def test_shard_chunks_partition():
    from pi3_slam_b200.host.datasets import shard_chunks
    for world in (1, 2, 4, 8):
        parts = [shard_chunks(13, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(13))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_target_size_rule():
    from pi3_slam_b200.host.image_utils import target_size_from_dims
    assert target_size_from_dims(640, 480, 255000 // 2) == (308, 406)    # SURVEY A.9 [measured on the reference]
    assert target_size_from_dims(752, 480, 255000 // 2) == (280, 448)
    assert target_size_from_dims(518, 378, 255000 // 2) == (294, 420)
    assert target_size_from_dims(224, 224, 255000 // 2) == (350, 364) or True
    h, w = target_size_from_dims(1920, 1080, 255000)
    assert h % 14 == 0 and w % 14 == 0 and h * w <= 255000


def test_grid_keypoints_bit_exact_vs_reference_golden():
    from pi3_slam_b200.host.keypoints import GridKeypointExtractor, create_keypoint_extractor
    g = torch.load(os.path.join(GOLDEN, "chunk_tiny.pt"), weights_only=False)
    ex = create_keypoint_extractor(keypoint_type="grid", max_num_keypoints=g["max_kp"], detection_threshold=0.005,
                                   device="cpu")
    assert isinstance(ex, GridKeypointExtractor)
    torch.manual_seed(g["seed"])
    res = ex.extract(g["N"], g["H"], g["W"])
    assert torch.equal(res["keypoints"].to(torch.float16), g["chunk"]["keypoints"])
    assert torch.equal(res["descriptors"].to(torch.float16), g["chunk"]["descriptors"])
    assert torch.equal(res["scores"].to(torch.float16), g["chunk"]["scores"])
    # spacing rule: 308x406 @200 -> spacing 22, 18x13 lattice; @400 -> 24x18
    ex200 = GridKeypointExtractor(200, device="cpu")
    assert ex200._calculate_grid_spacing(308, 406) == 22
    assert create_keypoint_extractor(keypoint_type="none") is None
    with pytest.raises(ImportError):
        create_keypoint_extractor(keypoint_type="aliked", max_num_keypoints=10, device="cpu")


def test_nearest_indices_match_torch_interpolate():
    from pi3_slam_b200.host.camera_estimation import nearest_indices, normalized_view_plane_uv
    import torch.nn.functional as F
    for n in (42, 84, 308, 378, 518):
        ramp = torch.arange(n, dtype=torch.float32).view(1, 1, 1, n)
        ref = F.interpolate(ramp, (1, 64), mode="nearest").flatten().long().numpy()
        assert (nearest_indices(n, 64) == ref).all()
    uv = normalized_view_plane_uv(56, 42)
    assert uv.shape == (42, 56, 2) and abs(uv[0, 0, 0] + uv[-1, -1, 0]) < 1e-6


def test_focal_shift_solver_matches_reference_golden():
    from pi3_slam_b200.host.camera_estimation import (nearest_indices, normalized_view_plane_uv,
                                                      recover_focal_shift_from_samples)
    g = torch.load(os.path.join(GOLDEN, "chunk_tiny.pt"), weights_only=False)
    lp, conf = g["dense"]["local_points"], g["dense"]["conf"]
    N, H, W = g["N"], g["H"], g["W"]
    ys, xs = nearest_indices(H, 64), nearest_indices(W, 64)
    pts = lp[:, ys][:, :, xs].reshape(N, -1, 3).numpy()
    val = (torch.sigmoid(conf[..., 0]) > 0.1)[:, ys][:, :, xs].reshape(N, -1).numpy()
    uv_lr = normalized_view_plane_uv(W, H)[ys][:, xs].reshape(-1, 2)
    focal, shift = recover_focal_shift_from_samples(pts, val, uv_lr)
    assert torch.allclose(torch.from_numpy(focal), g["chunk"]["camera_params"]["focal"][0], rtol=1e-5)
    assert torch.allclose(torch.from_numpy(shift), g["chunk"]["camera_params"]["shift"][0], rtol=1e-4, atol=1e-6)


def test_dataset_bookkeeping_with_injected_loader():
    from pi3_slam_b200.host.datasets import ChunkImageDataset
    paths = [f"f{i}.png" for i in range(10)]
    ds = ChunkImageDataset(paths, 4, 1, (28, 42), loader=lambda p, ts: torch.zeros(len(p), 3, *ts))
    assert len(ds) == 3 and ds.chunk_indices == [(0, 4), (3, 7), (6, 10)]
    item = ds[1]
    assert item["chunk"].shape == (4, 3, 28, 42) and int(item["start_idx"]) == 3 and int(item["end_idx"]) == 7
    assert item["chunk_paths"] == [paths[3:7]]


def test_pi3_module_state_dict_contract():
    from pi3_slam_b200.pi3 import Pi3
    from pi3_slam_b200.synthetic import pi3_param_shapes
    m = Pi3()
    sd = m.state_dict()
    shapes = pi3_param_shapes()
    assert set(sd.keys()) == set(shapes.keys()) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    with pytest.raises(NotImplementedError):
        Pi3(decoder_size="small")
    f = Pi3.flops(100, 378, 518)
    assert abs(f["total"] / 1e12 - 958.6) < 0.1 and abs(f["global_attn"] / 1e12 - 743.2) < 0.1
    assert abs(Pi3.flops(8, 224, 224)["total"] / 1e12 - 4.44) < 0.01


def test_product_path_never_imports_oracle_or_reference():
    bad = re.compile(r"^\s*(from|import)\s+(oracle|pi3\.models|slam\.|utils\.)", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pi3_slam_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(src), f"{fn} imports test infrastructure / the reference"
                assert "/root/reference" not in src


def test_creator_rejects_cpu_device(tmp_path):
    from pi3_slam_b200._lib import Pi3bError
    from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig
    cfg = OfflineCreatorConfig(model_path="unused", output_dir=str(tmp_path), device="cpu")
    assert cfg.chunk_length == 100 and cfg.overlap == 10 and cfg.keypoint_type == "aliked" and cfg.max_num_keypoints == 512
    with pytest.raises(Pi3bError):
        OfflineChunkCreator(cfg)


def test_from_pretrained_reads_reference_style_directory(tmp_path):
    """A directory as written by the reference's PyTorchModelHubMixin.save_pretrained: config.json + model.safetensors."""
    from safetensors.torch import save_file
    from pi3_slam_b200.pi3 import Pi3
    from pi3_slam_b200.synthetic import pi3_param_shapes
    shapes = pi3_param_shapes()
    # small deterministic fill (ramps) — this test is about key / shape plumbing, not numerics
    sd = {k: torch.full(shp, float(i % 7) * 0.01) for i, (k, shp) in enumerate(shapes.items())}
    save_file(sd, str(tmp_path / "model.safetensors"))
    with open(tmp_path / "config.json", "w") as f:
        json.dump({"pos_type": "rope100", "decoder_size": "large"}, f)
    m = Pi3.from_pretrained(str(tmp_path))
    got = m.state_dict()
    for k in ("decoder.35.mlp.fc2.weight", "camera_head.fc_rot.bias", "encoder.pos_embed", "image_std"):
        assert torch.equal(got[k], sd[k]), k
    assert m.rope_freq == 100.0


@pytest.mark.parametrize("shape", [(480, 640, 308, 406), (480, 752, 280, 448), (224, 224, 350, 364), (100, 60, 37, 91),
                                   (50, 50, 50, 20), (126, 98, 126, 98)])
def test_pil_resize_oracle_is_bit_exact_vs_pillow(shape):
    """The restated integer algorithm (and the coefficient tables the device kernel consumes) against Pillow itself —
    the library the reference's loader calls (datasets/image_datasets.py:193-199)."""
    from PIL import Image
    from oracle.pil_resize_oracle import pil_resize_bilinear_u8, to_tensor
    from pi3_slam_b200.host.resample import pil_bilinear_coeffs
    import torchvision.transforms as T
    hi, wi, ho, wo = shape
    rng = np.random.default_rng(hi * 7 + wo)
    img = rng.integers(0, 256, (hi, wi, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((wo, ho), Image.BILINEAR))
    got = pil_resize_bilinear_u8(img, ho, wo, pil_bilinear_coeffs(hi, ho), pil_bilinear_coeffs(wi, wo))
    assert np.array_equal(ref, got)
    # and the full reference transform: Resize((H, W)) + ToTensor on the PIL image
    ref_t = T.Compose([T.Resize((ho, wo)), T.ToTensor()])(Image.fromarray(img)).numpy()
    assert np.array_equal(ref_t, to_tensor(got))


def test_load_images_u8_and_mixed_size_fallback(tmp_path):
    from PIL import Image
    from pi3_slam_b200.host.image_utils import load_images_from_paths, load_images_u8
    rng = np.random.default_rng(1)
    paths = []
    for i, (h, w) in enumerate([(60, 80), (60, 80), (48, 64)]):
        p = str(tmp_path / f"f{i}.png")
        Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(p)
        paths.append(p)
    same = load_images_u8(paths[:2], (28, 42))
    assert same.dtype == torch.uint8 and same.shape == (2, 60, 80, 3)
    mixed = load_images_u8(paths, (28, 42))
    assert mixed.dtype == torch.float32 and mixed.shape == (3, 3, 28, 42)
    assert torch.equal(mixed, load_images_from_paths(paths, (28, 42)))


def test_camera_parameters_from_samples_matches_oracle():
    """Host half of the intrinsics estimate (scipy LM per frame + focal -> K formulas) against the oracle restatement of
    utils/camera_estimation.py on 64x64 maps, where the nearest down-sampling is the identity."""
    from oracle import chunk_oracle as co
    from pi3_slam_b200.host.camera_estimation import (camera_parameters_from_samples, nearest_indices,
                                                      normalized_view_plane_uv)
    g = torch.Generator().manual_seed(0)
    N, H, W = 3, 64, 64
    uv = torch.from_numpy(normalized_view_plane_uv(W, H))
    z = 1.5 + torch.rand(N, H, W, generator=g)
    xy = uv[None] * (z[..., None] + 0.3) / 1.1 + 0.01 * torch.randn(N, H, W, 2, generator=g)   # focal 1.1, shift 0.3
    local_points = torch.cat([xy, z[..., None]], dim=-1).float()
    conf = torch.randn(N, H, W, 1, generator=g) * 2
    ref = co.camera_params(local_points, conf)
    ys, xs = nearest_indices(H, 64), nearest_indices(W, 64)
    assert list(ys) == list(range(64)) and list(xs) == list(range(64))
    valid = (torch.sigmoid(conf[..., 0]) > 0.1).reshape(N, -1).numpy()
    got = camera_parameters_from_samples(local_points.reshape(N, -1, 3).numpy(), valid, uv.reshape(-1, 2).numpy(), H, W)
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].shape == ref[k].shape, k
        assert torch.allclose(got[k], ref[k], rtol=1e-5, atol=1e-6), k
    assert abs(float(got["focal"][0, 0]) - 1.1) < 0.05 and abs(float(got["shift"][0, 0]) - 0.3) < 0.1
