"""CPU: the oracle restatements against the committed golden vectors produced by the REAL reference
(oracle/make_golden.py, oracle/make_golden_chunk.py), plus known-answer tests for the unpinned Sim3."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sd():
    from pi3_slam_b200.synthetic import make_state_dict
    return make_state_dict(0)


def test_synthetic_state_dict_layout(sd):
    from pi3_slam_b200.synthetic import pi3_param_shapes
    shapes = pi3_param_shapes()
    assert len(sd) == 1210 and list(sd.keys()) == list(shapes.keys())
    assert sum(v.numel() for k, v in sd.items() if k not in ("image_mean", "image_std")) == 958_696_732
    # deterministic: per-key generators
    from pi3_slam_b200.synthetic import _fill
    k = "decoder.7.attn.qkv.weight"
    assert torch.equal(sd[k], _fill(k, shapes[k], 0))


def test_oracle_forward_matches_reference_golden_tiny(sd):
    from oracle import pi3_oracle
    from pi3_slam_b200.synthetic import make_frames
    gold = torch.load(os.path.join(GOLDEN, "pi3_forward_tiny.pt"))
    out = pi3_oracle.pi3_forward(make_frames(gold["N"], gold["H"], gold["W"])[None], sd)
    for k in ("points", "local_points", "conf", "camera_poses"):
        assert out[k].shape == gold[k].shape
        assert (out[k] - gold[k]).abs().max() < 1e-4 * max(1.0, gold[k].abs().max().item()), k


def test_chunk_oracle_matches_reference_chunk_golden():
    from oracle import chunk_oracle as co
    g = torch.load(os.path.join(GOLDEN, "chunk_tiny.pt"), weights_only=False)
    chunk, dense = g["chunk"], g["dense"]
    N, H, W = g["N"], g["H"], g["W"]
    torch.manual_seed(g["seed"])
    kps = co.grid_keypoints(N, H, W, g["max_kp"])
    assert torch.equal(kps.to(torch.float16), chunk["keypoints"])            # bit-exact lattice + randperm order
    masks = co.compute_masks(dense["local_points"], dense["conf"])
    interp = co.interpolate_at_keypoints(dense["points"], dense["local_points"], dense["conf"], masks, kps, H, W)
    assert torch.equal(interp["points"].to(torch.float16), chunk["points"])
    assert torch.equal(interp["local_points"].to(torch.float16), chunk["local_points"])
    assert torch.equal(interp["conf"].to(torch.float16), chunk["conf"])
    assert torch.equal(interp["masks"], chunk["masks"])
    assert 0.05 < chunk["masks"].float().mean() < 0.95                       # the mask is a non-trivial mix
    from pi3_slam_b200.synthetic import make_frames
    col = co.colors_at_keypoints(make_frames(N, H, W), kps)
    assert torch.equal(col.to(torch.float16), chunk["colors"])
    cp = co.camera_params(dense["local_points"], dense["conf"])
    for k in ("intrinsics", "focal", "shift", "fx", "fy", "cx", "cy"):
        assert cp[k].shape == chunk["camera_params"][k].shape, k
        assert torch.allclose(cp[k], chunk["camera_params"][k], rtol=1e-5, atol=1e-6), k
    assert torch.equal(chunk["intrinsics"], chunk["camera_params"]["intrinsics"])
    assert torch.equal(chunk["camera_poses"], dense["camera_poses"])


def test_chunk_golden_contract_appendix_b():
    g = torch.load(os.path.join(GOLDEN, "chunk_tiny.pt"), weights_only=False)
    c, N, K = g["chunk"], g["N"], g["max_kp"]
    expect = {"points": ((N, K, 3), torch.float16), "local_points": ((N, K, 3), torch.float16),
              "conf": ((N, K, 1), torch.float16), "masks": ((N, K, 1), torch.bool),
              "camera_poses": ((N, 4, 4), torch.float32), "keypoints": ((N, K, 2), torch.float16),
              "descriptors": ((N, K, 128), torch.float16), "scores": ((N, K), torch.float16),
              "colors": ((N, K, 3), torch.float16), "intrinsics": ((N, 3, 3), torch.float32)}
    for k, (shape, dt) in expect.items():
        assert tuple(c[k].shape) == shape and c[k].dtype == dt, k
    assert c["original_width"] == g["W"] and c["original_height"] == g["H"]
    assert set(c["_metrics"]) == {"infer_s", "num_frames", "fps"}


def test_umeyama_known_answer():
    from oracle.chunk_oracle import sim3_matrix, umeyama_sim3
    rng = np.random.default_rng(0)
    q = rng.normal(size=(200, 3))
    A = rng.normal(size=(3, 3))
    U, _, Vt = np.linalg.svd(A)
    R = U @ np.diag([1, 1, np.linalg.det(U @ Vt)]) @ Vt
    s, t = 1.7, np.array([0.3, -2.0, 5.0])
    r = s * q @ R.T + t
    s2, R2, t2 = umeyama_sim3(q, r)
    assert abs(s2 - s) < 1e-10 and np.abs(R2 - R).max() < 1e-10 and np.abs(t2 - t).max() < 1e-9
    T = sim3_matrix(s2, R2, t2)
    assert np.allclose((T @ np.c_[q, np.ones(len(q))].T).T[:, :3], r)
    # reflection-prone (planar) configuration still yields a proper rotation
    qp = q.copy(); qp[:, 2] = 0
    rp = s * qp @ R.T + t
    _, R3, _ = umeyama_sim3(qp, rp)
    assert abs(np.linalg.det(R3) - 1) < 1e-9


def test_view_graph_matches():
    from oracle.chunk_oracle import create_view_graph_matches
    assert create_view_graph_matches(100, 20)[0] == (80, 0) and create_view_graph_matches(100, 20)[-1] == (99, 19)
    assert create_view_graph_matches(8, 2) == [(6, 0), (7, 1)]


def test_rope2d_restatement_matches_the_reference_native_kernel_source():
    """oracle.rope2d against the reference's own curope extension (pi3/models/curope/curope.cpp rope_2d, CPU path), built
    from the reference sources by oracle/build_ref_curope.py into oracle/_ref/ (skipped where that library is absent)."""
    import importlib.util
    import os

    import pytest
    import torch
    from oracle.pi3_oracle import rope2d

    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_curope.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/ref_curope.so not built")
    spec = importlib.util.spec_from_file_location("ref_curope", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(3)
    tok = torch.randn(2, 50, 4, 64, generator=g)
    pos = torch.randint(0, 40, (2, 50, 2), generator=g)
    want = rope2d(tok.transpose(1, 2), pos).transpose(1, 2)
    got = tok.clone()
    ref.rope_2d(got, pos, 100.0, 1.0)
    assert (got - want).abs().max().item() < 2e-5     # fp32 cosf/sinf/powf at angles up to 40 rad


def test_lm_restatement_matches_scipy():
    """oracle.lm_oracle (MINPACK lmder for one parameter, driven as scipy >= 1.16 drives it) against scipy itself: same
    evaluation count, same shift / focal — on the degenerate geometry of the chunk golden (scipy stops after 2
    evaluations) and on perspective-consistent clouds with noise (several LM iterations)."""
    from functools import partial

    import numpy as np
    from scipy.optimize import least_squares

    from oracle.lm_oracle import lm_focal_shift, residual

    def scipy_solve(uv, xyz):
        uv, xy, z = uv.reshape(-1, 2), xyz[..., :2].reshape(-1, 2), xyz[..., 2].reshape(-1)
        sol = least_squares(partial(residual, uv, xy, z), x0=0, ftol=1e-3, method="lm")
        s = sol["x"].squeeze().astype(np.float32)
        xp = xy / (z + s)[:, None]
        return s, (xp * uv).sum() / np.square(xp).sum(), sol.nfev

    rng = np.random.default_rng(0)
    for case in range(24):
        S = int(rng.integers(50, 4096))
        uv = (rng.random((S, 2)).astype(np.float32) - 0.5) * 1.2
        zt = rng.random(S).astype(np.float32) * 3 + 1.0
        ft = rng.random() * 1.5 + 0.4
        true_shift = rng.random() * 2 - 0.5
        xyz = np.concatenate([uv * zt[:, None] / ft, (zt - true_shift)[:, None]], 1).astype(np.float32)
        if case % 3 == 2:
            xyz = rng.normal(0, 1, xyz.shape).astype(np.float32)          # not a perspective cloud at all: degenerate
        else:
            xyz += rng.normal(0, 0.02 * (case % 4), xyz.shape).astype(np.float32)
        a, b = scipy_solve(uv, xyz), lm_focal_shift(uv, xyz)
        assert a[2] == b[2], (case, a, b)
        assert abs(float(a[0]) - float(b[0])) <= 1e-5 * max(1.0, abs(float(a[0]))), (case, a, b)
        assert abs(float(a[1]) - float(b[1])) <= 1e-5 * abs(float(a[1])), (case, a, b)


def test_recon_observation_list_restatement_matches_reference_golden():
    """oracle.recon_oracle (the stage-2 front end's track / observation construction) against the calls the REAL
    reference made on a recording pytheia stand-in (tests/golden/recon_tiny.pt, oracle/make_golden_recon.py)."""
    import os

    import numpy as np
    import torch

    from oracle.recon_oracle import observation_list

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    chunk = torch.load(os.path.join(root, "tests", "golden", "chunk_tiny.pt"), weights_only=False)["chunk"]
    gold = torch.load(os.path.join(root, "tests", "golden", "recon_tiny.pt"), weights_only=False)["cases"]
    for max_obs, ref in gold.items():
        mine = observation_list(chunk, max_obs)
        for k in ref:
            assert np.array_equal(ref[k], mine[k]), (max_obs, k)
    assert len(gold[5]["obs_view"]) > len(gold[2]["obs_view"]) > 120
