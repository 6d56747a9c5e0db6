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


# ---------------------------------------------------------------------------------------------------------------
# MoGe-2 (metric-depth pass). The reference builds MoGeModel(**checkpoint['model_config'])
# (moge/model/v2.py:89-94) and the checkpoint is a hub download that is not in the repo, so the architecture
# hyper-parameters below are a recorded synthetic choice in the shape of the published ViT-S model: DINOv2 ViT-S/14
# encoder with the last 4 blocks projected and summed, a 5-level ConvStack neck fed with view-plane uv at every level,
# points + mask ConvStack heads, an MLP scale head on the class token, 'exp' output remap.
# ---------------------------------------------------------------------------------------------------------------
MOGE_BACKBONES = {"dinov2_vits14": (384, 12, 6), "dinov2_vitb14": (768, 12, 12), "dinov2_vitl14": (1024, 24, 16)}


def moge_synth_config() -> dict:
    pyramid = [256, 128, 64, 32, 32]
    stack = dict(resamplers=["conv_transpose"] * 4, num_res_blocks=[1, 1, 1, 1, 1])
    return dict(
        encoder=dict(backbone="dinov2_vits14", intermediate_layers=4, dim_out=384),
        neck=dict(dim_in=[386, 2, 2, 2, 2], dim_res_blocks=[384, 256, 128, 64, 32], dim_out=list(pyramid), **stack),
        points_head=dict(dim_in=list(pyramid), dim_res_blocks=list(pyramid), dim_out=[None, None, None, None, 3], **stack),
        mask_head=dict(dim_in=list(pyramid), dim_res_blocks=list(pyramid), dim_out=[None, None, None, None, 1], **stack),
        scale_head=dict(dims=[384, 384, 1]),
        remap_output="exp",
        num_tokens_range=[1200, 3600],
    )


def _per_level(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


def _conv_stack_shapes(prefix: str, cfg: dict, s: Dict[str, Tuple[int, ...]]) -> None:
    """Keys of reference ConvStack (moge/model/modules.py:195-233) for resampler type 'conv_transpose'."""
    res = cfg["dim_res_blocks"]
    n = len(res)
    hidden_mul = cfg.get("dim_times_res_block_hidden", 1)
    for i, (din, dres) in enumerate(zip(_per_level(cfg["dim_in"], n), res)):
        if din is not None:
            s[f"{prefix}.input_blocks.{i}.weight"] = (dres, din, 1, 1)
            s[f"{prefix}.input_blocks.{i}.bias"] = (dres,)
    for i, (a, b, kind) in enumerate(zip(res[:-1], res[1:], _per_level(cfg["resamplers"], n - 1))):
        if kind != "conv_transpose":
            raise NotImplementedError(f"resampler '{kind}' (only 'conv_transpose' is built)")
        s[f"{prefix}.resamplers.{i}.0.weight"] = (a, b, 2, 2)
        s[f"{prefix}.resamplers.{i}.0.bias"] = (b,)
        s[f"{prefix}.resamplers.{i}.1.weight"] = (b, b, 3, 3)
        s[f"{prefix}.resamplers.{i}.1.bias"] = (b,)
    for i, (dres, nb) in enumerate(zip(res, _per_level(cfg.get("num_res_blocks", 1), n))):
        hid = hidden_mul * dres
        for j in range(nb):
            p = f"{prefix}.res_blocks.{i}.{j}.layers"
            if cfg.get("res_block_in_norm", "layer_norm") != "none":
                s[f"{p}.0.weight"] = (dres,)
                s[f"{p}.0.bias"] = (dres,)
            s[f"{p}.2.weight"] = (hid, dres, 3, 3)
            s[f"{p}.2.bias"] = (hid,)
            if cfg.get("res_block_hidden_norm", "group_norm") != "none":
                s[f"{p}.3.weight"] = (hid,)
                s[f"{p}.3.bias"] = (hid,)
            s[f"{p}.5.weight"] = (dres, hid, 3, 3)
            s[f"{p}.5.bias"] = (dres,)
    for i, (dout, dres) in enumerate(zip(_per_level(cfg["dim_out"], n), res)):
        if dout is not None:
            s[f"{prefix}.output_blocks.{i}.weight"] = (dout, dres, 1, 1)
            s[f"{prefix}.output_blocks.{i}.bias"] = (dout,)


def moge_param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """Reference `MoGeModel(**cfg).state_dict()` layout (moge/model/v2.py:29-56, modules.py:71-97, 184-233)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    enc = cfg["encoder"]
    dim, depth, _ = MOGE_BACKBONES[enc["backbone"]]
    s["encoder.image_mean"] = (1, 3, 1, 1)
    s["encoder.image_std"] = (1, 3, 1, 1)
    b = "encoder.backbone"
    s[f"{b}.cls_token"] = (1, 1, dim)
    s[f"{b}.pos_embed"] = (1, 1 + 37 * 37, dim)
    s[f"{b}.mask_token"] = (1, dim)
    s[f"{b}.patch_embed.proj.weight"] = (dim, 3, 14, 14)
    s[f"{b}.patch_embed.proj.bias"] = (dim,)
    for i in range(depth):
        p = f"{b}.blocks.{i}"
        s[f"{p}.norm1.weight"] = (dim,)
        s[f"{p}.norm1.bias"] = (dim,)
        s[f"{p}.attn.qkv.weight"] = (3 * dim, dim)
        s[f"{p}.attn.qkv.bias"] = (3 * dim,)
        s[f"{p}.attn.proj.weight"] = (dim, dim)
        s[f"{p}.attn.proj.bias"] = (dim,)
        s[f"{p}.ls1.gamma"] = (dim,)
        s[f"{p}.norm2.weight"] = (dim,)
        s[f"{p}.norm2.bias"] = (dim,)
        s[f"{p}.mlp.fc1.weight"] = (4 * dim, dim)
        s[f"{p}.mlp.fc1.bias"] = (4 * dim,)
        s[f"{p}.mlp.fc2.weight"] = (dim, 4 * dim)
        s[f"{p}.mlp.fc2.bias"] = (dim,)
        s[f"{p}.ls2.gamma"] = (dim,)
    s[f"{b}.norm.weight"] = (dim,)
    s[f"{b}.norm.bias"] = (dim,)
    il = enc["intermediate_layers"]
    for i in range(il if isinstance(il, int) else len(il)):
        s[f"encoder.output_projections.{i}.weight"] = (enc["dim_out"], dim, 1, 1)
        s[f"encoder.output_projections.{i}.bias"] = (enc["dim_out"],)
    for name in ("neck", "points_head", "mask_head", "normal_head"):
        if cfg.get(name) is not None:
            _conv_stack_shapes(name, cfg[name], s)
    if cfg.get("scale_head") is not None:
        dims = cfg["scale_head"]["dims"]
        for i, (a, c) in enumerate(zip(dims[:-1], dims[1:])):
            s[f"scale_head.{2 * i}.weight"] = (c, a)
            s[f"scale_head.{2 * i}.bias"] = (c,)
    return s


_MOGE_UV_QUIET = ("neck.input_blocks.4.weight", "neck.resamplers.3.1.weight", "neck.res_blocks.4.0.layers.5.weight",
                  "neck.output_blocks.4.weight", "points_head.input_blocks.4.weight", "points_head.resamplers.3.1.weight",
                  "points_head.res_blocks.4.0.layers.5.weight")


def _fill_moge(key: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    g = _gen("moge:" + key, seed)
    if key == "encoder.image_mean":
        return torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    if key == "encoder.image_std":
        return torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    if key.endswith("cls_token") or key.endswith("mask_token"):
        return torch.randn(shape, generator=g) * 0.02
    if key.endswith("pos_embed"):
        return torch.randn(shape, generator=g).clamp_(-2, 2) * 0.02
    if key.endswith("ls1.gamma") or key.endswith("ls2.gamma"):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if len(shape) == 1 and (".norm" in key or ".layers.0." in key or ".layers.3." in key):
        if key.endswith(".weight"):
            return 1.0 + 0.05 * torch.randn(shape, generator=g)
        return 0.02 * torch.randn(shape, generator=g)
    if key.endswith(".bias"):
        return 0.02 * torch.randn(shape, generator=g)
    if len(shape) == 4 and "patch_embed" not in key:
        # conv / transposed-conv kernels: unit-gain fan-in scaling so activations stay O(1) down the pyramid; the 2-channel
        # uv input convs get O(1) weights so the view-plane coordinates visibly steer the output
        fan_in = shape[1] * shape[2] * shape[3] if ".resamplers." not in key or key.endswith(".1.weight") else shape[0]
        std = 1.0 / fan_in ** 0.5
        if key.startswith("points_head.output_blocks"):
            std *= 0.1      # keeps z = exp(raw) within about [0.3, 3] like a real scene
        elif key.startswith("mask_head.output_blocks"):
            std *= 0.5
        w = torch.randn(shape, generator=g) * std
        # A trained MoGe emits a perspective-consistent point map (xy / z ~ uv / focal); random weights do not, and the
        # focal/shift solve of infer() (a 1-D LM with ftol 1e-3) is then degenerate — its answer is decided by rounding
        # noise, in the reference as much as here. Give the synthetic net the linear uv -> xy path a trained one has:
        # two channels of the finest level carry uv through the residual stream to the xy outputs.
        if key in _MOGE_UV_QUIET:
            w[:2] *= 0.02       # nothing else writes into the two uv-carrying channels
        if key == "neck.input_blocks.4.weight":
            w[0, 0, 0, 0], w[1, 1, 0, 0] = 3.0, 3.0
        elif key in ("neck.output_blocks.4.weight", "points_head.input_blocks.4.weight"):
            w[0, 0, 0, 0] = w[1, 1, 0, 0] = 1.0
        elif key == "points_head.output_blocks.4.weight":
            w[:2] *= 0.3
            w[0, 0, 0, 0] = w[1, 1, 0, 0] = 1.0 / (3.0 * 1.2)       # focal 1.2 (relative to the half diagonal)
        return w
    if "patch_embed" in key:
        return torch.randn(shape, generator=g) * 0.024
    return torch.randn(shape, generator=g).clamp_(-2, 2) * 0.02


def make_moge_state_dict(cfg: dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, shp in moge_param_shapes(cfg).items():
        sd[k] = _fill_moge(k, shp, seed).float().contiguous()
    return sd
