"""GPU parity of the MoGe metric-depth pass through the C ABI: the dense-prediction
kernels against torch fp32, the full MoGeModel forward / infer against the oracle and against the goldens produced by
the real reference, and the creator's metric-scale step with the native MoGe model.

Tolerances (bf16 tensor-core operands, fp32 accumulation, vs the fp32 reference): relative L2 1e-2 on the point map
and on the metric depth (the pointmap tolerance of BASELINE.json's north_star; measured 3.5-3.9e-3), 3e-2 absolute on
the mask probability (measured 1.5-1.8e-2; < 1 % of the pixels flip side of 0.5), 2e-2 on the chunk's scale factor
(measured 3.5e-3).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False          # the torch references below are meant to be fp32
torch.backends.cuda.matmul.allow_tf32 = False
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_layernorm_384_768():
    from pi3_slam_b200 import ops
    torch.manual_seed(0)
    for Dm in (384, 768):
        x = torch.randn(77, Dm, device="cuda") * 2 + 0.3
        g, b = torch.randn(Dm, device="cuda"), torch.randn(Dm, device="cuda")
        out = torch.empty(77, Dm, device="cuda")
        ops.layernorm(x, g, b, 1e-6, out)
        assert torch.allclose(out, F.layer_norm(x, (Dm,), g, b, 1e-6), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("C,groups,H,W", [(32, 1, 37, 53), (64, 2, 16, 24), (384, 1, 9, 12), (384, 12, 9, 12), (256, 8, 18, 24)])
def test_groupnorm_stats(C, groups, H, W):
    from pi3_slam_b200 import ops
    torch.manual_seed(C + groups)
    x = torch.randn(H * W, C, device="cuda") * 1.7 + 0.4
    st = ops.groupnorm_stats(x, groups, 1e-5)
    xg = x.double().view(H * W, groups, C // groups)
    mean = xg.mean(dim=(0, 2))
    var = xg.var(dim=(0, 2), unbiased=False)
    assert torch.allclose(st[:, 0].double(), mean, atol=1e-6, rtol=1e-6)
    assert torch.allclose(st[:, 1].double(), 1.0 / torch.sqrt(var + 1e-5), atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("C,H,W,norm", [(32, 10, 14, True), (64, 7, 5, False), (128, 6, 8, True)])
def test_im2col3x3_conv_matches_torch(C, H, W, norm):
    """im2col (+ GroupNorm + ReLU) followed by the GEMM == F.conv2d(replicate pad) of the normalised map."""
    from pi3_slam_b200 import ops
    torch.manual_seed(C)
    Cout = 48
    x = torch.randn(H * W, C, device="cuda")
    w = torch.randn(Cout, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    bias = torch.randn(Cout, device="cuda")
    gamma, beta = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    groups = C // 32
    col = torch.empty(H * W, 9 * C, device="cuda", dtype=torch.bfloat16)
    xn = x.t().reshape(1, C, H, W)
    if norm:
        st = ops.groupnorm_stats(x, groups)
        ops.im2col3x3(x, H, W, C, col, mean_rstd=st, gamma=gamma, beta=beta, relu=True)
        xn = F.relu(F.group_norm(xn, groups, gamma, beta, 1e-5))
    else:
        ops.im2col3x3(x, H, W, C, col)
    # the gather itself is exact up to the bf16 rounding of each element
    ref_col = F.unfold(F.pad(xn, (1, 1, 1, 1), mode="replicate"), 3).view(C, 9, H * W).permute(2, 1, 0).reshape(H * W, 9 * C)
    assert (col.float() - ref_col).abs().max() <= 0.01 * ref_col.abs().max()
    out = torch.empty(H * W, Cout, device="cuda")
    ops.gemm(col, w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).to(torch.bfloat16).contiguous(), out, bias=bias)
    ref = F.conv2d(F.pad(xn, (1, 1, 1, 1), mode="replicate"), w, bias)[0].reshape(Cout, H * W).t()
    assert rel_l2(out, ref) < 6e-3


def test_conv_transpose_blocked_then_conv3x3():
    """ConvTranspose2d(2, stride 2) as a GEMM into the blocked layout, read in place by the next 3x3 conv's im2col."""
    from pi3_slam_b200 import ops
    torch.manual_seed(3)
    Cin, C, H, W = 64, 32, 5, 7
    x = torch.randn(H * W, Cin, device="cuda")
    wt = torch.randn(Cin, C, 2, 2, device="cuda") / Cin ** 0.5
    bt = torch.randn(C, device="cuda")
    w3 = torch.randn(C, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    b3 = torch.randn(C, device="cuda")
    t = torch.empty(H * W, 4 * C, device="cuda")
    ops.gemm(x.to(torch.bfloat16), wt.permute(2, 3, 1, 0).reshape(4 * C, Cin).to(torch.bfloat16).contiguous(), t,
             bias=bt.repeat(4).contiguous())
    up_ref = F.conv_transpose2d(x.t().reshape(1, Cin, H, W), wt, bt, stride=2)           # (1, C, 2H, 2W)
    got_up = t.view(H, W, 2, 2, C).permute(0, 2, 1, 3, 4).reshape(2 * H * 2 * W, C)
    assert rel_l2(got_up, up_ref[0].reshape(C, -1).t()) < 6e-3
    col = torch.empty(4 * H * W, 9 * C, device="cuda", dtype=torch.bfloat16)
    ops.im2col3x3(t, 2 * H, 2 * W, C, col, blocked=True)
    out = torch.empty(4 * H * W, C, device="cuda")
    ops.gemm(col, w3.permute(0, 2, 3, 1).reshape(C, 9 * C).to(torch.bfloat16).contiguous(), out, bias=b3)
    ref = F.conv2d(F.pad(up_ref, (1, 1, 1, 1), mode="replicate"), w3, b3)[0].reshape(C, -1).t()
    assert rel_l2(out, ref) < 8e-3


def test_uv_affine_matches_1x1_conv_on_uv():
    from oracle.moge_oracle import view_plane_uv
    from pi3_slam_b200 import ops
    torch.manual_seed(5)
    H, W, C = 18, 24, 64
    aspect = 112 / 84
    uv = view_plane_uv(W, H, aspect).cuda()                                  # (H, W, 2)
    w = torch.randn(C, 2, device="cuda")
    b = torch.randn(C, device="cuda")
    x0 = torch.randn(H * W, C, device="cuda")
    ref = x0 + F.conv2d(uv.permute(2, 0, 1)[None], w.view(C, 2, 1, 1), b)[0].reshape(C, -1).t()
    x = x0.clone()
    ops.uv_affine(x, H, W, w, b, uv[0, :, 0].contiguous(), uv[:, 0, 1].contiguous(), accumulate=True)
    assert torch.allclose(x, ref, atol=1e-5, rtol=1e-5)
    ops.uv_affine(x, H, W, w, None, uv[0, :, 0].contiguous(), uv[:, 0, 1].contiguous(), accumulate=False)
    assert torch.allclose(x, ref - x0 - b, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("hi,wi,ho,wo", [(84, 112, 126, 168), (140, 182, 84, 98), (144, 192, 84, 112)])
def test_resample2d_matches_interpolate(hi, wi, ho, wo):
    from pi3_slam_b200 import ops
    from pi3_slam_b200.host.resample import bilinear_aa_taps, bilinear_taps
    torch.manual_seed(7)
    img = torch.rand(3, hi, wi, device="cuda")

    def dev(t):
        return torch.from_numpy(t[0]).cuda(), torch.from_numpy(t[1]).cuda()
    out = ops.resample2d(img, "chw", torch.empty(3, ho, wo, device="cuda"), "chw", dev(bilinear_aa_taps(hi, ho)),
                         dev(bilinear_aa_taps(wi, wo)))
    ref = F.interpolate(img[None], (ho, wo), mode="bilinear", align_corners=False, antialias=True)[0]
    assert (out - ref).abs().max() < 1e-5       # ATen CUDA builds the same triangle weights with different fp32 rounding
    hwc = img.permute(1, 2, 0).contiguous()
    out = ops.resample2d(hwc, "hwc", torch.empty(ho, wo, 3, device="cuda"), "hwc", dev(bilinear_taps(hi, ho)),
                         dev(bilinear_taps(wi, wo)))
    ref = F.interpolate(img[None], (ho, wo), mode="bilinear", align_corners=False, antialias=False)[0].permute(1, 2, 0)
    assert (out - ref).abs().max() < 5e-5


@pytest.fixture(scope="module")
def moge():
    from pi3_slam_b200.moge import MoGeModel
    from pi3_slam_b200.synthetic import make_moge_state_dict, moge_synth_config
    cfg = moge_synth_config()
    sd = make_moge_state_dict(cfg, 0)
    model = MoGeModel(**cfg)
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval(), sd, cfg


@pytest.mark.parametrize("case", ["up", "down"])
def test_moge_forward_and_infer_match_reference_golden(moge, case):
    from pi3_slam_b200.synthetic import make_frames
    model, _, _ = moge
    gold = torch.load(os.path.join(GOLDEN, f"moge_{case}.pt"))
    c = gold["case"]
    img = make_frames(1, c["H"], c["W"], start_index=c["frame"])[0].cuda()
    fwd = model.forward(img[None], num_tokens=c["num_tokens"])
    assert fwd["points"].shape == (1, c["H"], c["W"], 3) and fwd["mask"].shape == (1, c["H"], c["W"])
    e_pts = rel_l2(fwd["points"][0].cpu(), gold["fwd_points"])
    e_mask = float((fwd["mask"][0].cpu() - gold["fwd_mask"]).abs().max())
    e_scale = abs(float(fwd["metric_scale"][0].cpu() / gold["metric_scale"]) - 1)
    print(f"moge {case}: points relL2 {e_pts:.2e}  mask prob max|d| {e_mask:.2e}  metric_scale rel {e_scale:.2e}")
    assert e_pts < 1e-2 and e_mask < 3e-2 and e_scale < 2e-3
    inf = model.infer(img, num_tokens=c["num_tokens"])
    depth, mask = inf["depth"].cpu(), inf["mask"].cpu()
    assert depth.shape == (c["H"], c["W"]) and mask.dtype == torch.bool
    assert torch.equal(torch.isfinite(depth), mask)
    flips = float((mask != gold["mask"]).float().mean())
    both = mask & gold["mask"]
    e_depth = rel_l2(depth[both], gold["depth"][both])
    ratio = float((depth[both] / gold["depth"][both]).median())
    print(f"moge {case}: mask flips {flips:.2e}  depth relL2 {e_depth:.2e}  median ratio {ratio:.5f}")
    assert flips < 1e-2 and abs(ratio - 1) < 5e-3 and e_depth < 1e-2
    assert (inf["intrinsics"].cpu() - gold["intrinsics"]).abs().max() < 0.05 * gold["intrinsics"][1, 1]


def test_moge_infer_default_tokens_matches_oracle(moge):
    """Default infer(): 3600 base tokens (v2.py:223-225). 196x266 frame -> 51x70 tokens -> 816x1120 head maps."""
    from oracle import moge_oracle
    from pi3_slam_b200.synthetic import make_frames
    model, sd, cfg = moge
    img = make_frames(1, 196, 266, start_index=11)[0]
    ref = moge_oracle.moge_infer(sd, cfg, img)
    inf = model.infer(img.cuda())
    mask = inf["mask"].cpu()
    both = mask & ref["mask"]
    flips = float((mask != ref["mask"]).float().mean())
    ratio = float((inf["depth"].cpu()[both] / ref["depth"][both]).median())
    e_depth = rel_l2(inf["depth"].cpu()[both], ref["depth"][both])
    print(f"moge 3600 tokens: mask flips {flips:.2e}  depth relL2 {e_depth:.2e}  median ratio {ratio:.5f}")
    assert flips < 1e-2 and abs(ratio - 1) < 5e-3 and e_depth < 1e-2


def test_creator_metric_scale_with_native_moge(moge, tmp_path):
    """_process_single_chunk with do_metric_depth and the native MoGeModel: scale = lower median of
    moge_depth / pi3_z over frame 0's mask (offline_chunk_creator.py:182-192), all on the device; checked against the
    oracle's MoGe depth + torch.median on the same pi3 maps."""
    from oracle import moge_oracle
    from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig
    from pi3_slam_b200.pi3 import Pi3
    from pi3_slam_b200.synthetic import make_frames, make_state_dict
    model, sd, cfg = moge
    pi3 = Pi3()
    pi3.load_state_dict(make_state_dict(0), strict=True)
    pi3 = pi3.cuda().eval()
    N, H, W = 2, 84, 112
    imgs = make_frames(N, H, W)[None]
    base = dict(model_path="unused", output_dir=str(tmp_path), device="cuda", keypoint_type="none",
                estimate_camera_params=False)
    plain = OfflineChunkCreator(OfflineCreatorConfig(do_metric_depth=False, **base), model=pi3)
    plain.target_size = (H, W)
    ref = plain._process_single_chunk(imgs, [("a",), ("b",)])
    scaled = OfflineChunkCreator(OfflineCreatorConfig(do_metric_depth=True, **base), model=pi3, moge_model=model)
    scaled.target_size = (H, W)
    out = scaled._process_single_chunk(imgs, [("a",), ("b",)])
    ref_depth = moge_oracle.moge_infer(sd, cfg, imgs[0, 0])["depth"]
    m0 = ref["masks"][0]
    ref_scale = float((ref_depth[m0] / ref["local_points"][0][..., 2][m0]).median())
    got_scale = float((out["local_points"][0][..., 2] / ref["local_points"][0][..., 2]).median())
    print(f"metric scale: native {got_scale:.5f}  oracle {ref_scale:.5f}")
    assert np.isfinite(got_scale) and abs(got_scale / ref_scale - 1) < 2e-2
    assert torch.allclose(out["camera_poses"][:, :3, 3], ref["camera_poses"][:, :3, 3] * got_scale, rtol=1e-4)


def test_cfg4_shaped_end_to_end_with_moge_keypoints_and_sim3(moge, tmp_path):
    """BASELINE configs[3] in miniature: EuRoC-shaped 752x480 frames (target 280x448), pi3 + native MoGe metric scale +
    grid keypoints (max-kp 200) + Sim3 between consecutive chunks, through process_and_save and the chunk files."""
    from oracle import chunk_oracle as co
    from pi3_slam_b200.alignment import align_consecutive_chunks
    from pi3_slam_b200.offline_chunk_creator import OfflineChunkCreator, OfflineCreatorConfig
    from pi3_slam_b200.pi3 import Pi3
    from pi3_slam_b200.synthetic import make_frames, make_state_dict
    model, _, _ = moge
    pi3 = Pi3()
    pi3.load_state_dict(make_state_dict(0), strict=True)
    pi3 = pi3.cuda().eval()
    H, W, L, ov = 280, 448, 4, 2
    cfg = OfflineCreatorConfig(model_path="unused", output_dir=str(tmp_path), device="cuda", do_metric_depth=True,
                               keypoint_type="grid", max_num_keypoints=200, chunk_length=L, overlap=ov, keypoint_seed=7)
    creator = OfflineChunkCreator(cfg, model=pi3, moge_model=model)
    paths = [f"/data/euroc/{i:05d}.png" for i in range(8)]
    index = {p: i for i, p in enumerate(paths)}
    files = creator.process_and_save(paths, loader=lambda ps, ts: torch.cat([make_frames(1, *ts, start_index=index[p]) for p in ps]),
                                     target_size=(H, W))
    chunks = [torch.load(f, weights_only=False) for f in files]
    # reference index math (datasets/image_datasets.py:41-47): (0,4) (2,6) (4,8) and a 2-frame tail (6,8)
    assert len(chunks) == 4 and chunks[3]["points"].shape == (2, 200, 3)
    chunks = chunks[:3]
    assert all(c["points"].shape == (L, 200, 3) for c in chunks)
    # metric scale was applied: same chunk without MoGe differs by one positive scalar on points and translations
    plain = OfflineChunkCreator(OfflineCreatorConfig(model_path="unused", output_dir=str(tmp_path / "plain"), device="cuda",
                                                     do_metric_depth=False, keypoint_type="grid", max_num_keypoints=200,
                                                     chunk_length=L, overlap=ov, keypoint_seed=7), model=pi3)
    files0 = plain.process_and_save(paths, loader=lambda ps, ts: torch.cat([make_frames(1, *ts, start_index=index[p]) for p in ps]),
                                    target_size=(H, W))
    c0 = torch.load(files0[0], weights_only=False)
    assert torch.equal(c0["keypoints"], chunks[0]["keypoints"])
    ratio = (chunks[0]["points"].float().norm(dim=-1) / c0["points"].float().norm(dim=-1)).flatten()
    scale = float(ratio.median())
    assert 0.05 < scale < 50 and float((ratio / scale - 1).abs().max()) < 5e-3      # fp16 storage
    t_ratio = chunks[0]["camera_poses"][1:, :3, 3].norm(dim=-1) / c0["camera_poses"][1:, :3, 3].norm(dim=-1)
    assert torch.allclose(t_ratio, torch.full_like(t_ratio, scale), rtol=5e-3)
    # Sim3 chain on the saved chunks: device kernel == float64 oracle on the same files
    res = align_consecutive_chunks(chunks, L, ov)
    for k in range(2):
        ref = co.chunk_pair_alignment(chunks[k], chunks[k + 1], L, ov)
        assert np.abs(res["pairwise"][k].numpy() - ref).max() < 5e-3 * max(1.0, np.abs(ref).max())
    print(f"cfg4-shaped: metric scale {scale:.4f}; pairwise Sim3 scales {[float(np.cbrt(np.linalg.det(p.numpy()[:3, :3]))) for p in res['pairwise']]}")
