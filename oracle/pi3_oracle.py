"""ORACLE — test infrastructure only. Never imported by the product path (pi3_slam_b200/).

CPU restatement (plain PyTorch fp32, functional, state_dict in -> dict out) of the reference's pi3
chunk forward. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may use it. Each function cites the reference code it restates (paths relative to the reference repo).

Pinned: `oracle/make_golden.py` runs the *real* reference (`pi3.models.pi3.Pi3`, imported from
/root/reference in the build container) on the same synthetic state_dict and frames, checks this
restatement against it (max abs diff < 1e-4 at fp32) and commits the reference's outputs as
tests/golden/pi3_forward_*.pt; tests/test_oracle.py re-checks the restatement against those files.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

PATCH = 14
HEADS = 16
HD = 64
N_SPECIAL = 5  # decoder register tokens (pi3.py:86-91); the encoder has 1 cls + 4 registers = 5 as well


def _linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


# ---- RoPE-2D: pos_embed.py:112-159 (PyTorch RoPE2D, the path the reference actually takes), freq 100 ----
def rope_cos_sin(max_pos: int, dtype=torch.float32, base: float = 100.0, d: int = 32, device=None):
    inv_freq = 1.0 / (base ** (torch.arange(0, d, 2, device=device).float() / d))
    t = torch.arange(max_pos + 1, dtype=inv_freq.dtype, device=device)
    freqs = torch.einsum("i,j->ij", t, inv_freq).to(dtype)
    freqs = torch.cat((freqs, freqs), dim=-1)
    return freqs.cos(), freqs.sin()


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def rope2d(tokens: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
    """tokens (B, H, S, 64); positions (B, S, 2) int64 (y, x)."""
    cos, sin = rope_cos_sin(int(positions.max()), tokens.dtype, device=tokens.device)
    y, x = tokens.chunk(2, dim=-1)

    def apply1d(tok, pos1d):
        c = F.embedding(pos1d, cos)[:, None, :, :]
        s = F.embedding(pos1d, sin)[:, None, :, :]
        return tok * c + _rotate_half(tok) * s

    return torch.cat((apply1d(y, positions[:, :, 0]), apply1d(x, positions[:, :, 1])), dim=-1)


def _sdpa(q, k, v):
    """F.scaled_dot_product_attention(q, k, v) (attention.py:102-104, :336-338), (B, H, S, 64) fp32.

    On the CPU this IS the torch call the reference makes. On a CUDA device (the GPU-side fp32 oracle used by the parity
    tests at the benchmarked sizes, S up to 100 400) the same softmax(q k^T / sqrt(d)) v is evaluated explicitly in
    fp32 over blocks of query rows: torch's fused CUDA backends either do not take fp32 (flash) or may use TF32 tensor
    cores internally (memory-efficient), and the math backend would materialise H*S*S scores at once. Callers switch
    torch.backends.cuda.matmul.allow_tf32 off, so every product here is a true fp32 GEMM."""
    if not q.is_cuda:
        return F.scaled_dot_product_attention(q, k, v)
    B, H, S, d = q.shape
    out = torch.empty_like(q)
    scale = 1.0 / math.sqrt(d)
    blk = max(64, min(S, (1 << 30) // max(1, B * H * k.shape[2])))     # <= 4 GiB of fp32 scores per block
    kt = k.transpose(-1, -2)
    for i in range(0, S, blk):
        s_blk = torch.matmul(q[:, :, i:i + blk] * scale, kt)
        out[:, :, i:i + blk] = torch.matmul(torch.softmax(s_blk, dim=-1), v)
        del s_blk
    return out


# ---- attention: layers/attention.py:94-113 (FlashAttention) and :323-347 (FlashAttentionRope) ----
def attention(x, sd, prefix, xpos=None, qk_norm=False):
    B, S, C = x.shape
    qkv = _linear(x, sd, prefix + ".qkv").reshape(B, S, 3, HEADS, C // HEADS).transpose(1, 3)
    q, k, v = [qkv[:, :, i] for i in range(3)]
    if qk_norm:
        q = _ln(q, sd, prefix + ".q_norm", 1e-5).to(v.dtype)
        k = _ln(k, sd, prefix + ".k_norm", 1e-5).to(v.dtype)
    if xpos is not None:
        q = rope2d(q, xpos)
        k = rope2d(k, xpos)
    o = _sdpa(q, k, v)
    o = o.transpose(1, 2).reshape(B, S, C)
    return _linear(o, sd, prefix + ".proj")


# ---- block: dinov2/layers/block.py:88-113 (eval) and layers/block.py:310-335 (eval), LN eps 1e-6 ----
def block(x, sd, prefix, xpos=None, qk_norm=False, layerscale=True):
    a = attention(_ln(x, sd, prefix + ".norm1", 1e-6), sd, prefix + ".attn", xpos, qk_norm)
    if layerscale:
        a = a * sd[prefix + ".ls1.gamma"]
    x = x + a
    h = _linear(_ln(x, sd, prefix + ".norm2", 1e-6), sd, prefix + ".mlp.fc1")
    h = _linear(F.gelu(h), sd, prefix + ".mlp.fc2")
    if layerscale:
        h = h * sd[prefix + ".ls2.gamma"]
    return x + h


# ---- encoder: dinov2/models/vision_transformer.py:181-234 (tokens) and :259-278 (forward_features) ----
def interpolate_pos_encoding(sd, h0: int, w0: int) -> torch.Tensor:
    """Bicubic + antialias resize of the 37x37 pos-embed grid to (h0, w0) patches (interpolate_offset 0)."""
    pos_embed = sd["encoder.pos_embed"].float()
    n = pos_embed.shape[1] - 1
    m = int(math.sqrt(n))
    if h0 * w0 == n and h0 == w0:
        return pos_embed
    class_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    patch_pos = F.interpolate(patch_pos.reshape(1, m, m, dim).permute(0, 3, 1, 2), mode="bicubic", antialias=True,
                              size=(h0, w0))
    patch_pos = patch_pos.permute(0, 2, 3, 1).view(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def encoder_tokens(imgs_norm, sd):
    BN, _, H, W = imgs_norm.shape
    x = F.conv2d(imgs_norm, sd["encoder.patch_embed.proj.weight"], sd["encoder.patch_embed.proj.bias"], stride=PATCH)
    x = x.flatten(2).transpose(1, 2)  # patch_embed.py:75-78
    x = torch.cat((sd["encoder.cls_token"].expand(BN, -1, -1), x), dim=1)
    x = x + interpolate_pos_encoding(sd, H // PATCH, W // PATCH)
    x = torch.cat((x[:, :1], sd["encoder.register_tokens"].expand(BN, -1, -1), x[:, 1:]), dim=1)
    return x


def encoder(imgs_norm, sd, taps: Optional[dict] = None):
    x = encoder_tokens(imgs_norm, sd)
    if taps is not None:
        taps["enc_tokens"] = x.clone()
    for i in range(24):
        x = block(x, sd, f"encoder.blocks.{i}")
        if taps is not None and i == 0:
            taps["enc_block0"] = x.clone()
    x = _ln(x, sd, "encoder.norm", 1e-6)
    return x[:, 5:]  # x_norm_patchtokens


# ---- Pi3.decode: pi3.py:132-171 ; positions: pos_embed.py:162-174 ----
def positions(bn: int, h: int, w: int, device=None) -> torch.Tensor:
    pos = torch.cartesian_prod(torch.arange(h, device=device), torch.arange(w, device=device))
    pos = pos.view(1, h * w, 2).expand(bn, -1, 2).clone()
    pos = pos + 1
    special = torch.zeros(bn, N_SPECIAL, 2, dtype=pos.dtype, device=device)
    return torch.cat([special, pos], dim=1)


def decode(hidden, sd, N, H, W, taps: Optional[dict] = None):
    BN, hw, _ = hidden.shape
    B = BN // N
    reg = sd["register_token"].repeat(B, N, 1, 1).reshape(BN, N_SPECIAL, -1)
    hidden = torch.cat([reg, hidden], dim=1)
    hw = hidden.shape[1]
    pos = positions(BN, H // PATCH, W // PATCH, hidden.device)
    outs = []
    for i in range(36):
        if i % 2 == 0:
            pos = pos.reshape(BN, hw, -1)
            hidden = hidden.reshape(BN, hw, -1)
        else:
            pos = pos.reshape(B, N * hw, -1)
            hidden = hidden.reshape(B, N * hw, -1)
        hidden = block(hidden, sd, f"decoder.{i}", xpos=pos, qk_norm=True)
        if taps is not None and i in (0, 1):
            taps[f"dec_block{i}"] = hidden.reshape(BN, hw, -1).clone()
        if i + 1 in (35, 36):
            outs.append(hidden.reshape(BN, hw, -1))
    return torch.cat(outs, dim=-1), pos.reshape(BN, hw, -1)


# ---- TransformerDecoder: layers/transformer_head.py:48-56 ----
def head_decoder(hidden, sd, name, xpos):
    x = _linear(hidden, sd, name + ".projects")
    for i in range(5):
        x = block(x, sd, f"{name}.blocks.{i}", xpos=xpos, qk_norm=False, layerscale=False)
    return _linear(x, sd, name + ".linear_out")


# ---- LinearPts3d: layers/transformer_head.py:70-81 ----
def linear_pts3d(tokens, sd, name, H, W):
    B, S, Dm = tokens.shape
    feat = _linear(tokens, sd, name + ".proj")
    feat = feat.transpose(-1, -2).view(B, -1, H // PATCH, W // PATCH)
    feat = F.pixel_shuffle(feat, PATCH)
    return feat.permute(0, 2, 3, 1)


# ---- CameraHead: layers/camera_head.py:48-93 ----
def svd_orthogonalize(m):
    m = m.reshape(-1, 3, 3)
    mt = torch.transpose(F.normalize(m, p=2, dim=-1), -1, -2)
    u, s, v = torch.svd(mt)
    det = torch.det(torch.matmul(v, u.transpose(-2, -1)))
    return torch.matmul(torch.cat([v[:, :, :-1], v[:, :, -1:] * det.view(-1, 1, 1)], dim=2), u.transpose(-2, -1))


def camera_head(feat, sd, ph, pw):
    BN, hw, c = feat.shape
    for i in range(2):
        x = F.relu(_linear(feat, sd, f"camera_head.res_conv.{i}.res_conv1"))
        x = F.relu(_linear(x, sd, f"camera_head.res_conv.{i}.res_conv2"))
        x = F.relu(_linear(x, sd, f"camera_head.res_conv.{i}.res_conv3"))
        feat = feat + x
    feat = F.adaptive_avg_pool2d(feat.permute(0, 2, 1).reshape(BN, -1, ph, pw).contiguous(), 1).view(BN, -1)
    feat = F.relu(_linear(feat, sd, "camera_head.more_mlps.0"))
    feat = F.relu(_linear(feat, sd, "camera_head.more_mlps.2"))
    out_t = _linear(feat, sd, "camera_head.fc_t")
    out_r = _linear(feat, sd, "camera_head.fc_rot")
    pose = torch.zeros(BN, 4, 4, device=feat.device)
    pose[:, :3, :3] = svd_orthogonalize(out_r)
    pose[:, :3, 3] = out_t
    pose[:, 3, 3] = 1.0
    return pose


# ---- Pi3.forward: pi3.py:173-216 ----
@torch.no_grad()
def pi3_forward(imgs: torch.Tensor, sd: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """imgs (B,N,3,H,W) fp32 in [0,1]; sd: fp32 state_dict with the reference's keys. CPU tensors: the pinned CPU
    restatement. CUDA tensors (imgs and sd on the same device): the same functions evaluated on the GPU in fp32 —
    used by the parity tests at the benchmarked chunk sizes, where the CPU would need hours (see `_sdpa`)."""
    imgs = (imgs - sd["image_mean"]) / sd["image_std"]
    B, N, _, H, W = imgs.shape
    ph, pw = H // PATCH, W // PATCH
    hidden = encoder(imgs.reshape(B * N, 3, H, W), sd, taps)
    if taps is not None:
        taps["enc_out"] = hidden.clone()
    hidden, pos = decode(hidden, sd, N, H, W, taps)
    if taps is not None:
        taps["dec_out"] = hidden.clone()
    point_hidden = head_decoder(hidden, sd, "point_decoder", pos)
    conf_hidden = head_decoder(hidden, sd, "conf_decoder", pos)
    camera_hidden = head_decoder(hidden, sd, "camera_decoder", pos)
    if taps is not None:
        taps["point_hidden"] = point_hidden.clone()
        taps["camera_hidden"] = camera_hidden.clone()
    ret = linear_pts3d(point_hidden.float()[:, N_SPECIAL:], sd, "point_head", H, W).reshape(B, N, H, W, -1)
    xy, z = ret.split([2, 1], dim=-1)
    z = torch.exp(z)
    local_points = torch.cat([xy * z, z], dim=-1)
    conf = linear_pts3d(conf_hidden.float()[:, N_SPECIAL:], sd, "conf_head", H, W).reshape(B, N, H, W, -1)
    camera_poses = camera_head(camera_hidden.float()[:, N_SPECIAL:], sd, ph, pw).reshape(B, N, 4, 4)
    homog = torch.cat([local_points, torch.ones_like(local_points[..., :1])], dim=-1)
    points = torch.einsum("bnij, bnhwj -> bnhwi", camera_poses, homog)[..., :3]
    return dict(points=points, local_points=local_points, conf=conf, camera_poses=camera_poses)


class fp32_strict:
    """Context manager for the GPU-side oracle: no TF32 anywhere (cuBLAS matmuls and cuDNN convolutions in true fp32)."""

    def __enter__(self):
        self._mm = torch.backends.cuda.matmul.allow_tf32
        self._cd = torch.backends.cudnn.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        return self

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self._mm
        torch.backends.cudnn.allow_tf32 = self._cd
        return False


def pi3_forward_gpu_fp32(imgs: torch.Tensor, sd: Dict[str, torch.Tensor], device="cuda") -> Dict[str, torch.Tensor]:
    """The oracle forward evaluated on a CUDA device in strict fp32 (see `_sdpa`); returns CUDA tensors."""
    sd_dev = {k: v.to(device=device, dtype=torch.float32) for k, v in sd.items()}
    with fp32_strict():
        return pi3_forward(imgs.to(device=device, dtype=torch.float32), sd_dev)


# ---- error metrics used by the parity tests ----
def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def rotation_angle_rad(Ra: torch.Tensor, Rb: torch.Tensor) -> torch.Tensor:
    """Geodesic angle between batches of rotations (...,3,3): |Ra - Rb|_F = 2 sqrt(2) sin(theta/2)
    (well conditioned for small angles, unlike acos of the trace)."""
    d = (Ra.double() - Rb.double()).flatten(-2).norm(dim=-1)
    return 2.0 * torch.asin((d / (2.0 * math.sqrt(2.0))).clamp(max=1.0))


def pose_errors(pa: torch.Tensor, pb: torch.Tensor) -> Tuple[float, float]:
    """(max rotation error in rad, max relative translation error) between pose batches (…,4,4)."""
    ang = rotation_angle_rad(pa[..., :3, :3], pb[..., :3, :3]).max()
    ta, tb = pa[..., :3, 3].double(), pb[..., :3, 3].double()
    rel_t = ((ta - tb).norm(dim=-1) / tb.norm(dim=-1).clamp_min(1e-12)).max()
    return float(ang), float(rel_t)
