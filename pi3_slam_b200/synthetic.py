"""Deterministic synthetic weights and frames for parity tests and benchmarks (no network, no checkpoint).

`pi3_param_shapes()` is the reference `Pi3().state_dict()` layout (1210 entries; reference:
pi3/models/pi3.py:17-129, dinov2/models/vision_transformer.py:88-179, layers/transformer_head.py:9-47,
layers/camera_head.py:34-46). `make_state_dict(seed)` fills it from per-key seeded CPU generators, so the
same bytes can be produced in this container (to feed the real reference when generating golden vectors)
and on the GPU box (to feed the oracle and the CUDA engine). Distributions follow the reference init
(trunc-normal 0.02 Linear weights, LayerScale 1.0 / 0.01, tokens at 1e-6) but biases, LayerNorm affine
parameters and LayerScale get small random perturbations so that a dropped bias or gamma is visible.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

D = 1024
ENC_DEPTH = 24
DEC_DEPTH = 36
HEAD_DEPTH = 5


def _block(prefix: str, shapes: Dict[str, Tuple[int, ...]], ls: bool, qknorm: bool) -> None:
    shapes[f"{prefix}.norm1.weight"] = (D,)
    shapes[f"{prefix}.norm1.bias"] = (D,)
    shapes[f"{prefix}.attn.qkv.weight"] = (3 * D, D)
    shapes[f"{prefix}.attn.qkv.bias"] = (3 * D,)
    shapes[f"{prefix}.attn.proj.weight"] = (D, D)
    shapes[f"{prefix}.attn.proj.bias"] = (D,)
    if qknorm:
        shapes[f"{prefix}.attn.q_norm.weight"] = (64,)
        shapes[f"{prefix}.attn.q_norm.bias"] = (64,)
        shapes[f"{prefix}.attn.k_norm.weight"] = (64,)
        shapes[f"{prefix}.attn.k_norm.bias"] = (64,)
    if ls:
        shapes[f"{prefix}.ls1.gamma"] = (D,)
    shapes[f"{prefix}.norm2.weight"] = (D,)
    shapes[f"{prefix}.norm2.bias"] = (D,)
    shapes[f"{prefix}.mlp.fc1.weight"] = (4 * D, D)
    shapes[f"{prefix}.mlp.fc1.bias"] = (4 * D,)
    shapes[f"{prefix}.mlp.fc2.weight"] = (D, 4 * D)
    shapes[f"{prefix}.mlp.fc2.bias"] = (D,)
    if ls:
        shapes[f"{prefix}.ls2.gamma"] = (D,)


def pi3_param_shapes() -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["register_token"] = (1, 1, 5, D)
    s["image_mean"] = (1, 3, 1, 1)
    s["image_std"] = (1, 3, 1, 1)
    s["encoder.cls_token"] = (1, 1, D)
    s["encoder.pos_embed"] = (1, 1 + 37 * 37, D)
    s["encoder.register_tokens"] = (1, 4, D)
    s["encoder.patch_embed.proj.weight"] = (D, 3, 14, 14)
    s["encoder.patch_embed.proj.bias"] = (D,)
    for i in range(ENC_DEPTH):
        _block(f"encoder.blocks.{i}", s, ls=True, qknorm=False)
    s["encoder.norm.weight"] = (D,)
    s["encoder.norm.bias"] = (D,)
    for i in range(DEC_DEPTH):
        _block(f"decoder.{i}", s, ls=True, qknorm=True)
    for name, out_dim in (("point_decoder", 1024), ("conf_decoder", 1024), ("camera_decoder", 512)):
        s[f"{name}.projects.weight"] = (D, 2 * D)
        s[f"{name}.projects.bias"] = (D,)
        for i in range(HEAD_DEPTH):
            _block(f"{name}.blocks.{i}", s, ls=False, qknorm=False)
        s[f"{name}.linear_out.weight"] = (out_dim, D)
        s[f"{name}.linear_out.bias"] = (out_dim,)
    s["point_head.proj.weight"] = (3 * 196, D)
    s["point_head.proj.bias"] = (3 * 196,)
    s["conf_head.proj.weight"] = (196, D)
    s["conf_head.proj.bias"] = (196,)
    for i in range(2):
        for j in (1, 2, 3):
            s[f"camera_head.res_conv.{i}.res_conv{j}.weight"] = (512, 512)
            s[f"camera_head.res_conv.{i}.res_conv{j}.bias"] = (512,)
    for i in (0, 2):
        s[f"camera_head.more_mlps.{i}.weight"] = (512, 512)
        s[f"camera_head.more_mlps.{i}.bias"] = (512,)
    s["camera_head.fc_t.weight"] = (3, 512)
    s["camera_head.fc_t.bias"] = (3,)
    s["camera_head.fc_rot.weight"] = (9, 512)
    s["camera_head.fc_rot.bias"] = (9,)
    return s


def _gen(key: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def _fill(key: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    g = _gen(key, seed)
    if key == "image_mean":
        return torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    if key == "image_std":
        return torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    if key in ("register_token", "encoder.cls_token", "encoder.register_tokens"):
        # reference: std 1e-6; scaled up so that these rows matter numerically in parity tests
        return torch.randn(shape, generator=g) * 0.02
    if key == "encoder.pos_embed":
        return torch.randn(shape, generator=g).clamp_(-2, 2) * 0.02
    if key.endswith("ls1.gamma") or key.endswith("ls2.gamma"):
        base = 1.0 if key.startswith("encoder.") else 0.01
        return base * (1.0 + 0.1 * torch.randn(shape, generator=g))
    if ".norm" in key or "q_norm" in key or "k_norm" in key:
        if key.endswith(".weight"):
            return 1.0 + 0.05 * torch.randn(shape, generator=g)
        return 0.02 * torch.randn(shape, generator=g)
    if key == "camera_head.fc_rot.bias":
        # a trained camera head emits near-orthonormal 3x3 blocks; a zero-mean random 9-vector would make the
        # SVD orthogonalisation arbitrarily ill-conditioned and turn bf16 noise into large rotation errors
        return torch.eye(3).flatten() + 0.05 * torch.randn(shape, generator=g)
    if key == "camera_head.fc_t.bias":
        return 0.5 * torch.randn(shape, generator=g)
    if key == "point_head.proj.bias":
        b = 0.02 * torch.randn(shape, generator=g)
        b[2 * 196:] *= 0.015
        return b
    if key.endswith(".bias"):
        return 0.02 * torch.randn(shape, generator=g)
    if key == "encoder.patch_embed.proj.weight":
        return torch.randn(shape, generator=g) * 0.024
    # every nn.Linear weight: trunc-normal(std 0.02), as timm's init in the reference
    w = torch.randn(shape, generator=g).clamp_(-2, 2) * 0.02
    if key == "point_head.proj.weight":
        # log-depth channel (rows 392..587): keep neighbouring pixels within a few percent of each other so
        # that the depth-edge mask (rtol 0.03) is a non-trivial mix of true and false on synthetic weights
        w[2 * 196:] *= 0.015
    return w


def make_state_dict(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """fp32 CPU state_dict with the reference's 1210 keys; identical bytes on every machine."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, shp in pi3_param_shapes().items():
        sd[k] = _fill(k, shp, seed).float().contiguous()
    return sd


def make_frames(num_frames: int, height: int, width: int, start_index: int = 0) -> torch.Tensor:
    """Seeded uniform [0,1) frames, (N,3,H,W) fp32 — frame i depends only on its global index."""
    out = torch.empty(num_frames, 3, height, width)
    for i in range(num_frames):
        g = torch.Generator(device="cpu")
        g.manual_seed(1234 + start_index + i)
        out[i] = torch.rand(3, height, width, generator=g)
    return out
