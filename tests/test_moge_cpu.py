"""CPU: the MoGe oracle against the goldens produced by the REAL reference MoGeModel (oracle/make_golden_moge.py),
the synthetic model_config / state_dict layout, and the host-built resampling tap tables against F.interpolate."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("case", ["up", "down"])
def test_moge_oracle_matches_reference_golden(case):
    from oracle import moge_oracle
    from pi3_slam_b200.synthetic import make_frames, make_moge_state_dict, moge_synth_config
    gold = torch.load(os.path.join(GOLDEN, f"moge_{case}.pt"))
    cfg = moge_synth_config()
    assert cfg == gold["model_config"]
    sd = make_moge_state_dict(cfg, gold["seed"])
    c = gold["case"]
    img = make_frames(1, c["H"], c["W"], start_index=c["frame"])[0]
    out = moge_oracle.moge_infer(sd, cfg, img, num_tokens=c["num_tokens"])
    assert (out["fwd_points"] - gold["fwd_points"]).abs().max() < 2e-4
    assert (out["fwd_mask"] - gold["fwd_mask"]).abs().max() < 1e-5
    assert abs(float(out["metric_scale"] - gold["metric_scale"])) < 1e-5
    assert torch.equal(out["mask"], gold["mask"])
    fin = torch.isfinite(gold["depth"])
    assert torch.equal(torch.isfinite(out["depth"]), fin) and 0.3 < fin.float().mean() < 0.9
    assert (out["depth"][fin] - gold["depth"][fin]).abs().max() < 2e-4
    assert (out["intrinsics"] - gold["intrinsics"]).abs().max() < 1e-5


def test_moge_state_dict_layout_is_the_reference_one():
    """Key list and shapes recorded from `MoGeModel(**cfg).state_dict()` of the reference (401 tensors, 32.78 M)."""
    from pi3_slam_b200.moge import MoGeModel
    from pi3_slam_b200.synthetic import make_moge_state_dict, moge_param_shapes, moge_synth_config
    cfg = moge_synth_config()
    shapes = moge_param_shapes(cfg)
    assert len(shapes) == 401
    assert sum(int(np.prod(s)) for s in shapes.values()) == 32_777_259
    assert shapes["neck.resamplers.0.0.weight"] == (384, 256, 2, 2)          # ConvTranspose2d: [Cin, Cout, 2, 2]
    assert shapes["points_head.output_blocks.4.weight"] == (3, 32, 1, 1)
    model = MoGeModel(**cfg)
    assert list(model.state_dict().keys()) == list(shapes.keys())
    model.load_state_dict(make_moge_state_dict(cfg, 0), strict=True)
    with pytest.raises(Exception):           # CPU tensors are refused: no fallback
        model.infer(torch.rand(3, 28, 28))


@pytest.mark.parametrize("shape", [(84, 112, 126, 168), (140, 182, 84, 98), (378, 518, 714, 980), (816, 1120, 378, 518)])
def test_tap_tables_match_aten(shape):
    from pi3_slam_b200.host.resample import apply_taps_np, bilinear_aa_taps, bilinear_taps
    hi, wi, ho, wo = shape
    torch.manual_seed(1)
    img = torch.rand(2, hi, wi)
    aa = F.interpolate(img[None], (ho, wo), mode="bilinear", align_corners=False, antialias=True)[0].numpy()
    got = apply_taps_np(img.numpy(), bilinear_aa_taps(hi, ho), bilinear_aa_taps(wi, wo))
    assert np.abs(aa - got).max() < 1e-6
    pl = F.interpolate(img[None], (ho, wo), mode="bilinear", align_corners=False, antialias=False)[0].numpy()
    got = apply_taps_np(img.numpy(), bilinear_taps(hi, ho), bilinear_taps(wi, wo))
    assert np.abs(pl - got).max() < 5e-5     # fp32 rounding of the source coordinate at ~1000-pixel offsets
