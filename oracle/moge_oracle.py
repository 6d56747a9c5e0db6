"""ORACLE — test infrastructure only. Never imported by the product path (pi3_slam_b200/).

CPU restatement (plain PyTorch fp32, functional, state_dict + model_config in -> dict out) of the reference's MoGe-2
metric-depth pass: `MoGeModel.forward` and `MoGeModel.infer` of moge/model/v2.py with the
modules of moge/model/modules.py and the DINOv2 backbone of moge/model/dinov2/. Only tests/, __graft_entry__.smoke()
and bench.py's CPU legs may use it.

Pinned: `oracle/make_golden_moge.py` runs the *real* reference class (imported from /root/reference in the build
container, with a two-function stand-in for the un-vendored `utils3d`) on the same synthetic model_config /
state_dict / frames, checks this restatement against it and commits the reference's outputs as
tests/golden/moge_*.pt; tests/test_oracle.py re-checks the restatement against those files.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

PATCH = 14
BACKBONES = {"dinov2_vits14": (384, 12, 6), "dinov2_vitb14": (768, 12, 12), "dinov2_vitl14": (1024, 24, 16)}


def _ln(x, sd, prefix, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


# ---- DINOv2 block: moge/model/dinov2/layers/block.py (eval path), attention.py (MemEffAttention -> SDPA) ----
def vit_block(x, sd, p, heads):
    B, S, C = x.shape
    qkv = _lin(_ln(x, sd, p + ".norm1"), sd, p + ".attn.qkv").reshape(B, S, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(B, S, C)
    x = x + _lin(o, sd, p + ".attn.proj") * sd[p + ".ls1.gamma"]
    h = _lin(F.gelu(_lin(_ln(x, sd, p + ".norm2"), sd, p + ".mlp.fc1")), sd, p + ".mlp.fc2")
    return x + h * sd[p + ".ls2.gamma"]


# ---- pos-embed: moge/model/dinov2/models/vision_transformer.py:187-221 (interpolate_offset 0.1, bicubic, no antialias) ----
def interpolate_pos_encoding(pos_embed: torch.Tensor, h0: int, w0: int, offset: float = 0.1) -> torch.Tensor:
    pos_embed = pos_embed.float()
    n = pos_embed.shape[1] - 1
    m = int(math.sqrt(n))
    if h0 * w0 == n and h0 == w0:
        return pos_embed
    dim = pos_embed.shape[-1]
    sx, sy = float(w0 + offset) / m, float(h0 + offset) / m
    patch = F.interpolate(pos_embed[:, 1:].reshape(1, m, m, dim).permute(0, 3, 1, 2), mode="bicubic", antialias=False,
                          scale_factor=(sy, sx))
    assert patch.shape[-2:] == (h0, w0)
    return torch.cat((pos_embed[:, :1], patch.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)


# ---- DINOv2Encoder.forward: moge/model/modules.py:120-136 ----
def encoder(sd, cfg, image, rows, cols):
    dim, depth, heads = BACKBONES[cfg["backbone"]]
    b = "encoder.backbone"
    x = F.interpolate(image, (rows * PATCH, cols * PATCH), mode="bilinear", align_corners=False, antialias=True)
    x = (x - sd["encoder.image_mean"]) / sd["encoder.image_std"]
    x = F.conv2d(x, sd[f"{b}.patch_embed.proj.weight"], sd[f"{b}.patch_embed.proj.bias"], stride=PATCH).flatten(2).transpose(1, 2)
    x = torch.cat((sd[f"{b}.cls_token"].expand(x.shape[0], -1, -1), x), dim=1)
    x = x + interpolate_pos_encoding(sd[f"{b}.pos_embed"], rows, cols)
    il = cfg["intermediate_layers"]
    take = list(range(depth - il, depth)) if isinstance(il, int) else list(il)
    outs = []
    for i in range(depth):
        x = vit_block(x, sd, f"{b}.blocks.{i}", heads)
        if i in take:
            outs.append(_ln(x, sd, f"{b}.norm"))             # get_intermediate_layers(norm=True)
    feat = 0
    for k, o in enumerate(outs):
        f = o[:, 1:].permute(0, 2, 1).unflatten(2, (rows, cols))
        feat = feat + F.conv2d(f, sd[f"encoder.output_projections.{k}.weight"], sd[f"encoder.output_projections.{k}.bias"])
    return feat, outs[-1][:, 0]


# ---- normalized_view_plane_uv: moge/utils/geometry_torch.py:40-52 ----
def view_plane_uv(width, height, aspect_ratio=None):
    if aspect_ratio is None:
        aspect_ratio = width / height
    sx = aspect_ratio / (1 + aspect_ratio ** 2) ** 0.5
    sy = 1 / (1 + aspect_ratio ** 2) ** 0.5
    u = torch.linspace(-sx * (width - 1) / width, sx * (width - 1) / width, width, dtype=torch.float32)
    v = torch.linspace(-sy * (height - 1) / height, sy * (height - 1) / height, height, dtype=torch.float32)
    u, v = torch.meshgrid(u, v, indexing="xy")
    return torch.stack([u, v], dim=-1)


def _per_level(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


def _conv3(x, sd, p):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), sd[p + ".weight"], sd[p + ".bias"])


# ---- ResidualConvBlock: moge/model/modules.py:18-68 (relu; in_norm GroupNorm(1,C), hidden_norm GroupNorm(C/32,C)) ----
def res_block(x, sd, p, in_norm="layer_norm", hidden_norm="group_norm"):
    def norm(t, kind, key):
        c = t.shape[1]
        if kind == "none":
            return t
        groups = 1 if kind == "layer_norm" else c // 32
        return F.group_norm(t, groups, sd[key + ".weight"], sd[key + ".bias"], 1e-5)
    h = _conv3(F.relu(norm(x, in_norm, p + ".layers.0")), sd, p + ".layers.2")
    h = _conv3(F.relu(norm(h, hidden_norm, p + ".layers.3")), sd, p + ".layers.5")
    return x + h       # in_channels == out_channels inside ConvStack -> identity skip


# ---- ConvStack.forward: moge/model/modules.py:242-254 (resampler 'conv_transpose' :160-165) ----
def conv_stack(sd, prefix, cfg, in_features: List[Optional[torch.Tensor]]) -> List[torch.Tensor]:
    res = cfg["dim_res_blocks"]
    n = len(res)
    dim_in, dim_out = _per_level(cfg["dim_in"], n), _per_level(cfg["dim_out"], n)
    nblocks = _per_level(cfg.get("num_res_blocks", 1), n)
    outs = []
    x = None
    for i in range(n):
        feat = in_features[i]
        if dim_in[i] is not None:
            feat = F.conv2d(feat, sd[f"{prefix}.input_blocks.{i}.weight"], sd[f"{prefix}.input_blocks.{i}.bias"])
        if i == 0:
            x = feat
        elif feat is not None:
            x = x + feat
        for j in range(nblocks[i]):
            x = res_block(x, sd, f"{prefix}.res_blocks.{i}.{j}", cfg.get("res_block_in_norm", "layer_norm"),
                          cfg.get("res_block_hidden_norm", "group_norm"))
        if dim_out[i] is not None:
            outs.append(F.conv2d(x, sd[f"{prefix}.output_blocks.{i}.weight"], sd[f"{prefix}.output_blocks.{i}.bias"]))
        else:
            outs.append(x)
        if i < n - 1:
            x = F.conv_transpose2d(x, sd[f"{prefix}.resamplers.{i}.0.weight"], sd[f"{prefix}.resamplers.{i}.0.bias"], stride=2)
            x = _conv3(x, sd, f"{prefix}.resamplers.{i}.1")
    return outs


def remap_points(points, mode):
    if mode == "linear":
        return points
    if mode == "sinh":
        return torch.sinh(points)
    xy, z = points.split([2, 1], dim=-1)
    if mode == "exp":
        z = torch.exp(z)
        return torch.cat([xy * z, z], dim=-1)
    if mode == "sinh_exp":
        return torch.cat([torch.sinh(xy), torch.exp(z)], dim=-1)
    raise ValueError(mode)


# ---- MoGeModel.forward: moge/model/v2.py:128-179 ----
@torch.no_grad()
def moge_forward(sd: Dict[str, torch.Tensor], cfg: dict, image: torch.Tensor, num_tokens: int) -> Dict[str, torch.Tensor]:
    B, _, H, W = image.shape
    aspect = W / H
    base_h, base_w = int((num_tokens / aspect) ** 0.5), int((num_tokens * aspect) ** 0.5)
    feat, cls = encoder(sd, cfg["encoder"], image, base_h, base_w)
    feats: List[Optional[torch.Tensor]] = [feat, None, None, None, None]
    for lvl in range(5):
        uv = view_plane_uv(base_w * 2 ** lvl, base_h * 2 ** lvl, aspect).permute(2, 0, 1).unsqueeze(0).expand(B, -1, -1, -1)
        feats[lvl] = uv if feats[lvl] is None else torch.cat([feats[lvl], uv], dim=1)
    feats = conv_stack(sd, "neck", cfg["neck"], feats)
    out = {}
    for name, head in (("points", "points_head"), ("normal", "normal_head"), ("mask", "mask_head")):
        if cfg.get(head) is not None:
            v = conv_stack(sd, head, cfg[head], feats)[-1]
            out[name] = F.interpolate(v, (H, W), mode="bilinear", align_corners=False, antialias=False)
    if "points" in out:
        out["points"] = remap_points(out["points"].permute(0, 2, 3, 1), cfg.get("remap_output", "linear"))
    if "normal" in out:
        out["normal"] = F.normalize(out["normal"].permute(0, 2, 3, 1), dim=-1)
    if "mask" in out:
        out["mask_logit"] = out["mask"].squeeze(1)
        out["mask"] = out["mask"].squeeze(1).sigmoid()
    if cfg.get("scale_head") is not None:
        dims = cfg["scale_head"]["dims"]
        h = cls
        for i in range(len(dims) - 1):
            h = _lin(h, sd, f"scale_head.{2 * i}")
            if i < len(dims) - 2:
                h = F.relu(h)
        out["metric_scale"] = h.squeeze(1).exp()
    return out


# ---- recover_focal_shift: moge/utils/geometry_torch.py:115-170 + geometry_numpy.py solve_optimal_focal_shift ----
def recover_focal_shift(points: torch.Tensor, mask: Optional[torch.Tensor]):
    from functools import partial
    from scipy.optimize import least_squares
    B, H, W = points.shape[:3]
    uv = view_plane_uv(W, H)
    pts_lr = F.interpolate(points.permute(0, 3, 1, 2), (64, 64), mode="nearest").permute(0, 2, 3, 1).numpy()
    uv_lr = F.interpolate(uv.unsqueeze(0).permute(0, 3, 1, 2), (64, 64), mode="nearest").squeeze(0).permute(1, 2, 0).numpy()
    m_lr = None if mask is None else (F.interpolate(mask.float().unsqueeze(1), (64, 64), mode="nearest").squeeze(1) > 0).numpy()
    focal, shift = [], []
    for i in range(B):
        p = pts_lr[i] if m_lr is None else pts_lr[i][m_lr[i]]
        u = uv_lr if m_lr is None else uv_lr[m_lr[i]]
        if u.shape[0] < 2:
            focal.append(1.0); shift.append(0.0)
            continue
        uvf, xy, z = u.reshape(-1, 2), p[..., :2].reshape(-1, 2), p[..., 2].reshape(-1)

        def fn(uvf, xy, z, s):
            xp = xy / (z + s)[:, None]
            f = (xp * uvf).sum() / np.square(xp).sum()
            return (f * xp - uvf).ravel()
        sol = least_squares(partial(fn, uvf, xy, z), x0=0, ftol=1e-3, method="lm")
        s = sol["x"].squeeze().astype(np.float32)
        xp = xy / (z + s)[:, None]
        focal.append(float((xp * uvf).sum() / np.square(xp).sum()))
        shift.append(float(s))
    return torch.tensor(focal, dtype=torch.float32), torch.tensor(shift, dtype=torch.float32)


# ---- MoGeModel.infer: moge/model/v2.py:181-290 (fp32 branch: use_fp16=False; fov_x=None) ----
@torch.no_grad()
def moge_infer(sd, cfg, image: torch.Tensor, num_tokens: Optional[int] = None, resolution_level: int = 9,
               apply_mask: bool = True) -> Dict[str, torch.Tensor]:
    omit = image.dim() == 3
    if omit:
        image = image.unsqueeze(0)
    H, W = image.shape[-2:]
    aspect = W / H
    if num_tokens is None:
        lo, hi = cfg.get("num_tokens_range", [1200, 3600])
        num_tokens = int(lo + (resolution_level / 9) * (hi - lo))
    out = moge_forward(sd, cfg, image.float(), num_tokens)
    points, mask, scale = out.get("points"), out.get("mask"), out.get("metric_scale")
    mask_binary = mask > 0.5 if mask is not None else None
    focal, shift = recover_focal_shift(points, mask_binary)
    fx = focal / 2 * (1 + aspect ** 2) ** 0.5 / aspect
    fy = focal / 2 * (1 + aspect ** 2) ** 0.5
    K = torch.zeros(points.shape[0], 3, 3)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = fx, fy, 0.5, 0.5, 1.0
    points = points.clone()
    points[..., 2] += shift[..., None, None]
    if mask_binary is not None:
        mask_binary = mask_binary & (points[..., 2] > 0)
    depth = points[..., 2].clone()
    if scale is not None:
        depth = depth * scale[:, None, None]
    if apply_mask and mask_binary is not None:
        depth = torch.where(mask_binary, depth, torch.inf)
    res = {"depth": depth, "mask": mask_binary, "intrinsics": K, "focal": focal, "shift": shift,
           "fwd_points": out["points"], "fwd_mask": out.get("mask"), "metric_scale": scale}
    res = {k: v for k, v in res.items() if v is not None}
    if omit:
        res = {k: v.squeeze(0) for k, v in res.items()}
    return res
