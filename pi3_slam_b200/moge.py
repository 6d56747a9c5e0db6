"""`MoGeModel` — drop-in for the reference's MoGe-2 class on the metric-scale step of the chunk path
(reference moge/model/v2.py:22-290, moge/model/modules.py:18-254).

Same constructor arguments (`MoGeModel(**checkpoint['model_config'])`), same `state_dict` keys (reference
checkpoints load), `from_pretrained(path_or_repo)`, `forward(image, num_tokens)` and `infer(image, ...)` with the
reference's return keys; `OfflineChunkCreator` only reads `infer(frame0)["depth"]` (slam/offline_chunk_creator.py:184).
The computation is a flat schedule of C-ABI launches (libpi3b200.so), no nn.Conv2d / autocast / CPU fallback:

  encoder   F.interpolate(antialias) -> pi3b_resample2d; patch embed, 12/24 DINOv2 blocks, final norm and the summed
            1x1 projections of the last 4 blocks -> the same tcgen05 GEMM / attention / LayerNorm kernels as the pi3 trunk
  ConvStack channels-last fp32 maps [H*W, C]; 1x1 conv = GEMM; 3x3 replicate conv = pi3b_im2col3x3 (GroupNorm + ReLU
            fused into the gather) + GEMM with the residual add fused in the epilogue; ConvTranspose 2x2 = GEMM whose
            blocked output the next im2col reads in place; the uv concat of every level = pi3b_uv_affine
  outputs   pi3b_resample2d (bilinear) -> pi3b_moge_remap; infer: 64x64 samples -> pi3b_focal_shift_lm (the reference's scipy LM, restated;
            moge/utils/geometry_torch.py:115-170) -> pi3b_moge_depth

Numerics: bf16 tensor-core operands, fp32 accumulation, fp32 maps / residuals / GroupNorm statistics (fp64 sums); the
reference runs this pass under fp16 autocast (v2.py:228). Built for resampler type 'conv_transpose', ReLU blocks,
in_norm 'layer_norm' | 'none', hidden_norm 'group_norm' | 'layer_norm' | 'none' (the released MoGe-2 configuration);
other ConvStack variants raise NotImplementedError.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .host import camera_estimation as camest
from .host.resample import bilinear_aa_taps, bilinear_taps
from .pi3 import K_PATCH, K_PATCH_PAD, PATCH, _set_nested
from .synthetic import MOGE_BACKBONES, _per_level, moge_param_shapes


class MoGeModel(nn.Module):
    def __init__(self, encoder: Dict[str, Any], neck: Dict[str, Any], points_head: Dict[str, Any] = None,
                 mask_head: Dict[str, Any] = None, normal_head: Dict[str, Any] = None,
                 scale_head: Dict[str, Any] = None, remap_output: str = "linear",
                 num_tokens_range: List[int] = [1200, 3600], **deprecated_kwargs):
        super().__init__()
        if deprecated_kwargs:
            warnings.warn(f"The following deprecated/invalid arguments are ignored: {deprecated_kwargs}")
        if remap_output not in ops.MOGE_REMAP:
            raise ValueError(f"Invalid remap output type: {remap_output}")
        self.remap_output = remap_output
        self.num_tokens_range = list(num_tokens_range)
        self.model_config = dict(encoder=encoder, neck=neck, points_head=points_head, mask_head=mask_head,
                                 normal_head=normal_head, scale_head=scale_head, remap_output=remap_output,
                                 num_tokens_range=self.num_tokens_range)
        for name in ("neck", "points_head", "mask_head", "normal_head"):
            c = self.model_config[name]
            if c is None:
                continue
            if c.get("activation", "relu") != "relu":
                raise NotImplementedError("ConvStack activation other than 'relu'")
            if c.get("res_block_in_norm", "layer_norm") not in ("layer_norm", "none") or \
                    c.get("res_block_hidden_norm", "group_norm") not in ("group_norm", "layer_norm", "none"):
                raise NotImplementedError("ConvStack norm variant (instance_norm) is not built")
        self.dim, self.depth, self.heads = MOGE_BACKBONES[encoder["backbone"]]
        for key, shape in moge_param_shapes(self.model_config).items():
            _set_nested(self, key, torch.zeros(shape), buffer=key in ("encoder.image_mean", "encoder.image_std"))
        with torch.no_grad():
            self.encoder.image_mean.copy_(torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
            self.encoder.image_std.copy_(torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        self._packed: Optional[dict] = None
        self._shape_cache: Dict[tuple, dict] = {}

    # ------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, model_kwargs: Optional[Dict[str, Any]] = None,
                        **hf_kwargs) -> "MoGeModel":
        """Checkpoint file `model.pt` = {'model_config': ..., 'model': state_dict} (v2.py:66-97), local path or hub id."""
        path = pretrained_model_name_or_path
        if not os.path.exists(str(path)):
            from huggingface_hub import hf_hub_download  # no network on the build boxes: raises there
            path = hf_hub_download(repo_id=pretrained_model_name_or_path, repo_type="model", filename="model.pt", **hf_kwargs)
        checkpoint = torch.load(path, map_location="cpu", weights_only=True)
        cfg = checkpoint["model_config"]
        if model_kwargs is not None:
            cfg.update(model_kwargs)
        model = cls(**cfg)
        model.load_state_dict(checkpoint["model"], strict=False)
        return model

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._packed, self._shape_cache = None, {}
        return out

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._packed, self._shape_cache = None, {}
        return out

    def _p(self, key: str) -> torch.Tensor:
        mod = self
        for part in key.split("."):
            mod = getattr(mod, part)
        return mod

    # ------------------------------------------------------------------ packing
    def _pack(self, device: torch.device) -> dict:
        if self._packed is not None and self._packed["device"] == device:
            return self._packed
        _lib.load()
        dim = self.dim

        def f32(key):
            return self._p(key).detach().to(device=device, dtype=torch.float32).contiguous()

        def bf(t):
            return t.to(device=device, dtype=torch.bfloat16).contiguous()

        def lin(key):           # nn.Linear or 1x1 conv -> [out, in] bf16
            w = self._p(key + ".weight").detach()
            return bf(w.reshape(w.shape[0], -1)), f32(key + ".bias")

        def conv3(key):         # [Cout, Cin, 3, 3] -> [Cout, (ky, kx, ci)] matching pi3b_im2col3x3's column order
            w = self._p(key + ".weight").detach()
            return bf(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)), f32(key + ".bias")

        pk: dict = {"device": device}
        b = "encoder.backbone"
        wp = torch.zeros(dim, K_PATCH_PAD, device=device, dtype=torch.bfloat16)
        wp[:, :K_PATCH] = self._p(f"{b}.patch_embed.proj.weight").detach().to(device).reshape(dim, K_PATCH).to(torch.bfloat16)
        pk["patch_w"], pk["patch_b"] = wp, f32(f"{b}.patch_embed.proj.bias")
        blocks = []
        for i in range(self.depth):
            p = f"{b}.blocks.{i}"
            blocks.append({
                "n1": (f32(p + ".norm1.weight"), f32(p + ".norm1.bias")), "n2": (f32(p + ".norm2.weight"), f32(p + ".norm2.bias")),
                "qkv": lin(p + ".attn.qkv"), "proj": lin(p + ".attn.proj"), "fc1": lin(p + ".mlp.fc1"), "fc2": lin(p + ".mlp.fc2"),
                "ls1": f32(p + ".ls1.gamma"), "ls2": f32(p + ".ls2.gamma")})
        pk["blocks"] = blocks
        pk["norm"] = (f32(f"{b}.norm.weight"), f32(f"{b}.norm.bias"))
        il = self.model_config["encoder"]["intermediate_layers"]
        pk["take"] = list(range(self.depth - il, self.depth)) if isinstance(il, int) else list(il)
        pk["proj"] = [lin(f"encoder.output_projections.{k}") for k in range(len(pk["take"]))]
        pk["mean"] = self.encoder.image_mean.detach().flatten().tolist()
        pk["std"] = self.encoder.image_std.detach().flatten().tolist()

        def stack(name, n_uv_first: int, n_uv_rest: int):
            cfg = self.model_config[name]
            res = cfg["dim_res_blocks"]
            n = len(res)
            s = {"res": res, "dim_in": _per_level(cfg["dim_in"], n), "dim_out": _per_level(cfg["dim_out"], n),
                 "in": [], "out": [], "up": [], "blocks": []}
            in_norm, hid_norm = cfg.get("res_block_in_norm", "layer_norm"), cfg.get("res_block_hidden_norm", "group_norm")
            for i in range(n):
                ent = None
                if s["dim_in"][i] is not None:
                    w = self._p(f"{name}.input_blocks.{i}.weight").detach().reshape(res[i], -1)
                    n_uv = n_uv_first if i == 0 else n_uv_rest
                    n_feat = w.shape[1] - n_uv
                    ent = {"w": bf(w[:, :n_feat]) if n_feat > 0 else None,
                           "w_uv": w[:, n_feat:].to(device=device, dtype=torch.float32).contiguous() if n_uv else None,
                           "b": f32(f"{name}.input_blocks.{i}.bias")}
                s["in"].append(ent)
                s["out"].append(lin(f"{name}.output_blocks.{i}") if s["dim_out"][i] is not None else None)
                lvl = []
                for j in range(_per_level(cfg.get("num_res_blocks", 1), n)[i]):
                    p = f"{name}.res_blocks.{i}.{j}.layers"
                    hid = cfg.get("dim_times_res_block_hidden", 1) * res[i]
                    lvl.append({
                        "gn1": (f32(p + ".0.weight"), f32(p + ".0.bias"), 1) if in_norm != "none" else None,
                        "c1": conv3(p + ".2"),
                        "gn2": (f32(p + ".3.weight"), f32(p + ".3.bias"), hid // 32 if hid_norm == "group_norm" else 1)
                        if hid_norm != "none" else None,
                        "c2": conv3(p + ".5"), "hid": hid})
                s["blocks"].append(lvl)
                if i < n - 1:
                    wt = self._p(f"{name}.resamplers.{i}.0.weight").detach()      # [Cin, Cout, 2, 2]
                    s["up"].append({"wt": bf(wt.permute(2, 3, 1, 0).reshape(4 * wt.shape[1], wt.shape[0])),
                                    "bt": f32(f"{name}.resamplers.{i}.0.bias").repeat(4).contiguous(),
                                    "c": conv3(f"{name}.resamplers.{i}.1")})
            return s
        pk["neck"] = stack("neck", 2, 2)
        for h in ("points_head", "mask_head", "normal_head"):
            pk[h] = stack(h, 0, 0) if self.model_config[h] is not None else None
        if self.model_config["scale_head"] is not None:
            dims = self.model_config["scale_head"]["dims"]
            pk["scale"] = [lin(f"scale_head.{2 * i}") for i in range(len(dims) - 1)]
        else:
            pk["scale"] = None
        self._packed = pk
        return pk

    def _shape_consts(self, H: int, W: int, rows: int, cols: int, device: torch.device) -> dict:
        """Per-shape constants: token table (cls + bicubic-resized pos-embed, vision_transformer.py:187-221 with the
        0.1 interpolate offset), resize tap tables, per-level view-plane uv tables (geometry_torch.py:40-52)."""
        key = (H, W, rows, cols, str(device))
        c = self._shape_cache.get(key)
        if c is not None:
            return c
        dim = self.dim
        pos = self._p("encoder.backbone.pos_embed").detach().float().cpu()
        n = pos.shape[1] - 1
        m = int(math.sqrt(n))
        if rows * cols == n and rows == cols:
            patch_pos = pos[0, 1:]
        else:
            patch_pos = F.interpolate(pos[:, 1:].reshape(1, m, m, dim).permute(0, 3, 1, 2), mode="bicubic", antialias=False,
                                      scale_factor=(float(rows + 0.1) / m, float(cols + 0.1) / m))
            assert patch_pos.shape[-2:] == (rows, cols)
            patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(-1, dim)
        table = torch.empty(1 + rows * cols, dim)
        table[0] = self._p("encoder.backbone.cls_token").detach().float().cpu().reshape(dim) + pos[0, 0]
        table[1:] = patch_pos

        def dev_taps(t):
            return torch.from_numpy(t[0]).to(device), torch.from_numpy(t[1]).to(device).contiguous()
        aspect = W / H
        sx, sy = aspect / (1 + aspect ** 2) ** 0.5, 1 / (1 + aspect ** 2) ** 0.5
        uv = []
        for lvl in range(5):
            w_l, h_l = cols * 2 ** lvl, rows * 2 ** lvl
            u = torch.linspace(-sx * (w_l - 1) / w_l, sx * (w_l - 1) / w_l, w_l, dtype=torch.float32)
            v = torch.linspace(-sy * (h_l - 1) / h_l, sy * (h_l - 1) / h_l, h_l, dtype=torch.float32)
            uv.append((u.to(device), v.to(device)))
        ho, wo = rows * 16, cols * 16
        ys, xs = camest.nearest_indices(H, 64), camest.nearest_indices(W, 64)
        c = {
            "table": table.to(device).contiguous(),
            "in_y": dev_taps(bilinear_aa_taps(H, rows * PATCH)), "in_x": dev_taps(bilinear_aa_taps(W, cols * PATCH)),
            "out_y": dev_taps(bilinear_taps(ho, H)), "out_x": dev_taps(bilinear_taps(wo, W)),
            "uv": uv, "ys": torch.from_numpy(ys).to(device), "xs": torch.from_numpy(xs).to(device),
            "uv_lr": camest.normalized_view_plane_uv(W, H)[ys][:, xs].reshape(-1, 2),
        }
        self._shape_cache[key] = c
        return c

    # ------------------------------------------------------------------ building blocks
    def _vit_block(self, x, b, h, qkv, att, mlp, T) -> None:
        ops.layernorm(x, b["n1"][0], b["n1"][1], 1e-6, h)
        ops.gemm(h, b["qkv"][0], qkv, bias=b["qkv"][1])
        ops.attention_qkv(qkv, att, 1, T, self.heads, 0.125)
        ops.gemm(att, b["proj"][0], x, bias=b["proj"][1], gamma=b["ls1"], resid=x)
        ops.layernorm(x, b["n2"][0], b["n2"][1], 1e-6, h)
        ops.gemm(h, b["fc1"][0], mlp, bias=b["fc1"][1], act=ops.ACT_GELU)
        ops.gemm(mlp, b["fc2"][0], x, bias=b["fc2"][1], gamma=b["ls2"], resid=x)

    @staticmethod
    def _bf16(x: torch.Tensor) -> torch.Tensor:
        return ops.cast_bf16(x, torch.empty(x.shape, device=x.device, dtype=torch.bfloat16))

    def _conv3(self, x, H, W, C, wb, col, out=None, resid=None, blocked=False, gn=None) -> torch.Tensor:
        """3x3 replicate conv of a channels-last map: (GroupNorm + ReLU) -> im2col -> GEMM (+ bias, + residual)."""
        a = col[: H * W * 9 * C].view(H * W, 9 * C)
        if gn is not None:
            stats = ops.groupnorm_stats(x, gn[2], 1e-5)
            ops.im2col3x3(x, H, W, C, a, blocked=blocked, mean_rstd=stats, gamma=gn[0], beta=gn[1], relu=True)
        else:
            ops.im2col3x3(x, H, W, C, a, blocked=blocked, relu=False)
        if out is None:
            out = torch.empty(H * W, wb[0].shape[0], device=x.device, dtype=torch.float32)
        return ops.gemm(a, wb[0], out, bias=wb[1], resid=resid)

    def _conv_stack(self, s: dict, feats: List[Optional[torch.Tensor]], uv, rows: int, cols: int, col: torch.Tensor,
                    need: Optional[List[int]] = None) -> List[Optional[torch.Tensor]]:
        """ConvStack.forward (modules.py:242-254). feats[i]: fp32 [H_i*W_i, C] or None; uv[i] = (u_tab, v_tab) when the
        level's input carries the 2 view-plane channels (after the feature channels), else None."""
        res = s["res"]
        n = len(res)
        outs: List[Optional[torch.Tensor]] = []
        x = None
        for i in range(n):
            H, W, C = rows << i, cols << i, res[i]
            ent = s["in"][i]
            if ent is None:
                if feats[i] is not None:
                    raise NotImplementedError("ConvStack level with an identity input block and a feature input")
            else:
                first = x is None
                if first:
                    x = torch.empty(H * W, C, device=col.device, dtype=torch.float32)
                bias_pending = ent["b"]
                if ent["w"] is not None:
                    ops.gemm(self._bf16(feats[i]), ent["w"], x, bias=bias_pending, resid=None if first else x)
                    bias_pending, first = None, False
                if ent["w_uv"] is not None:
                    ops.uv_affine(x, H, W, ent["w_uv"], bias_pending, uv[i][0], uv[i][1], accumulate=not first)
            for blk in s["blocks"][i]:
                hmap = self._conv3(x, H, W, C, blk["c1"], col, gn=blk["gn1"])
                self._conv3(hmap, H, W, blk["hid"], blk["c2"], col, out=x, resid=x, gn=blk["gn2"])
            if need is None or i in need:
                outs.append(ops.gemm(self._bf16(x), s["out"][i][0],
                                     torch.empty(H * W, s["dim_out"][i], device=x.device, dtype=torch.float32),
                                     bias=s["out"][i][1]) if s["out"][i] is not None else x)
            else:
                outs.append(None)
            if i < n - 1:
                up = s["up"][i]
                t = ops.gemm(self._bf16(x), up["wt"], torch.empty(H * W, 4 * res[i + 1], device=x.device, dtype=torch.float32),
                             bias=up["bt"])
                x = self._conv3(t, 2 * H, 2 * W, res[i + 1], up["c"], col, blocked=True)
        return outs

    # ------------------------------------------------------------------ forward
    def _forward_one(self, img: torch.Tensor, num_tokens: int, pk: dict, want_normal: bool = True) -> dict:
        device = img.device
        _, H, W = img.shape
        aspect = W / H
        rows, cols = int((num_tokens / aspect) ** 0.5), int((num_tokens * aspect) ** 0.5)
        sc = self._shape_consts(H, W, rows, cols, device)
        dim = self.dim
        P = rows * cols
        T = P + 1
        e = lambda *s, dt=torch.float32: torch.empty(*s, device=device, dtype=dt)  # noqa: E731
        bf = torch.bfloat16
        # ---- encoder (modules.py:120-136)
        img14 = ops.resample2d(img, "chw", e(3, rows * PATCH, cols * PATCH), "chw", sc["in_y"], sc["in_x"])
        patches = ops.patch_pack(img14.unsqueeze(0), e(P, K_PATCH_PAD, dt=bf), pk["mean"], pk["std"])
        x = e(T, dim)
        ops.broadcast_rows(sc["table"], x, 1)
        ops.gemm(patches, pk["patch_w"], x, bias=pk["patch_b"], resid=x, row_remap=(P, T, 1))
        h, qkv, att, mlp = e(T, dim, dt=bf), e(T, 3 * dim, dt=bf), e(T, dim, dt=bf), e(T, 4 * dim, dt=bf)
        dim_out = pk["proj"][0][0].shape[0]
        feat = e(P, dim_out)
        k = 0
        for i, b in enumerate(pk["blocks"]):
            self._vit_block(x, b, h, qkv, att, mlp, T)
            if i in pk["take"]:
                ops.layernorm(x, pk["norm"][0], pk["norm"][1], 1e-6, h)
                ops.gemm(h[1:], pk["proj"][k][0], feat, bias=pk["proj"][k][1], resid=feat if k else None)
                k += 1
        out: dict = {}
        if pk["scale"] is not None:          # MLP on the class token of the last tapped block (v2.py:154, modules.py:134)
            t = h[:1]
            for j, (w, bias) in enumerate(pk["scale"]):
                last = j == len(pk["scale"]) - 1
                t = ops.gemm(t, w, e(1, w.shape[0], dt=torch.float32 if last else bf), bias=bias,
                             act=ops.ACT_NONE if last else ops.ACT_RELU)
            out["log_scale"] = t.reshape(1)
        # ---- neck + heads (v2.py:138-153)
        stacks = [pk["neck"]] + [pk[hd] for hd in ("points_head", "mask_head", "normal_head") if pk[hd] is not None]
        col_elems = 0
        for s in stacks:
            for i, c in enumerate(s["res"]):
                hid = max([blk["hid"] for blk in s["blocks"][i]] + [c])
                col_elems = max(col_elems, (rows << i) * (cols << i) * 9 * hid)
        col = e(col_elems, dt=bf)
        uv = sc["uv"]
        neck = self._conv_stack(pk["neck"], [feat, None, None, None, None], uv, rows, cols, col)
        last = len(pk["neck"]["res"]) - 1
        Ho, Wo = rows << last, cols << last
        for name, hd in (("points", "points_head"), ("mask", "mask_head"), ("normal", "normal_head")):
            if pk[hd] is None or (name == "normal" and not want_normal):
                continue
            o = self._conv_stack(pk[hd], neck, [None] * 5, rows, cols, col, need=[last])[last]
            C = o.shape[1]
            out[name] = ops.resample2d(o.view(Ho, Wo, C), "hwc", e(H, W, C), "hwc", sc["out_y"], sc["out_x"])
        if "points" in out:
            logit = out["mask"].view(H, W) if "mask" in out else None
            prob = e(H, W) if logit is not None else None
            ops.moge_remap(out["points"], logit, prob, self.remap_output)
            if logit is not None:
                out["mask_logit"], out["mask"] = logit, prob
        elif "mask" in out:
            out["mask_logit"] = out["mask"].view(H, W)
            out["mask"] = torch.sigmoid(out["mask_logit"])
        out["_sc"] = sc
        return out

    def _check(self, image: torch.Tensor) -> None:
        if not image.is_cuda:
            raise _lib.Pi3bError("pi3_slam_b200.MoGeModel runs on CUDA (sm_100a) only; there is no CPU fallback")

    @ops.device_guard(lambda self, image, *a: image)
    @torch.no_grad()
    def forward(self, image: torch.Tensor, num_tokens: int) -> Dict[str, torch.Tensor]:
        """image (B,3,H,W) in [0,1] -> points (B,H,W,3), mask (B,H,W) probability, metric_scale (B,) [, normal]."""
        self._check(image)
        pk = self._pack(image.device)
        res: Dict[str, list] = {}
        for b in range(image.shape[0]):
            o = self._forward_one(image[b].to(torch.float32).contiguous(), num_tokens, pk)
            if "normal" in o:
                o["normal"] = F.normalize(o["normal"], dim=-1)
            if "log_scale" in o:
                o["metric_scale"] = o.pop("log_scale").exp()
            for key in ("points", "normal", "mask", "metric_scale"):
                if key in o:
                    res.setdefault(key, []).append(o[key])
        return {k: (torch.cat(v) if k == "metric_scale" else torch.stack(v)) for k, v in res.items()}

    @ops.device_guard(lambda self, image, *a: image)
    @torch.inference_mode()
    def infer(self, image: torch.Tensor, num_tokens: int = None, resolution_level: int = 9, force_projection: bool = True,
              apply_mask: bool = True, fov_x=None, use_fp16: bool = True) -> Dict[str, torch.Tensor]:
        """Reference `MoGeModel.infer` (v2.py:181-290): depth (H,W) metric, +inf where masked out; mask; intrinsics
        (normalised, cx = cy = 0.5). `points` is returned in camera space from the affine point map (z + shift); the
        force_projection re-projection goes through the un-vendored utils3d in the reference and is not reproduced
        (the chunk path reads ["depth"] only). `use_fp16` is accepted and ignored: operands are bf16 on tensor cores."""
        if fov_x is not None:
            raise NotImplementedError("infer(fov_x=...) — the chunk path always recovers the focal from the point map")
        omit = image.dim() == 3
        if omit:
            image = image.unsqueeze(0)
        self._check(image)
        pk = self._pack(image.device)
        H, W = image.shape[-2:]
        aspect = W / H
        if num_tokens is None:
            lo, hi = self.num_tokens_range
            num_tokens = int(lo + (resolution_level / 9) * (hi - lo))
        res: Dict[str, list] = {}
        for b in range(image.shape[0]):
            o = self._forward_one(image[b].to(torch.float32).contiguous(), num_tokens, pk, want_normal=False)
            if "points" not in o:
                raise _lib.Pi3bError("MoGeModel.infer needs a points head")
            sc = o["_sc"]
            pts = o["points"]
            # 64x64 nearest samples of (points, mask > 0.5) -> host LM over the z shift (geometry_torch.py:141-162)
            if "mask_logit" in o:
                s_pts, s_val = ops.sample_grid(pts.unsqueeze(0), o["mask_logit"].unsqueeze(0), sc["ys"], sc["xs"], 0.5)
            else:
                zeros = torch.zeros(1, H, W, device=pts.device)
                s_pts, _ = ops.sample_grid(pts.unsqueeze(0), zeros, sc["ys"], sc["xs"], 0.5)
                s_val = torch.ones(1, s_pts.shape[1], device=pts.device, dtype=torch.uint8)
            # the 1-parameter LM solve on the device (pi3b_focal_shift_lm: the algorithm scipy executes, restated)
            sol = camest.solve_focal_shift_on_device(s_pts, s_val, sc["uv_lr"]).cpu().numpy()
            shift, focal = sol[:, 0], sol[:, 1]
            depth, mask = ops.moge_depth(pts, o.get("mask"), float(shift[0]), o.get("log_scale"), apply_mask)
            f = float(focal[0])
            K = torch.zeros(3, 3, device=pts.device)
            K[0, 0] = f / 2 * (1 + aspect ** 2) ** 0.5 / aspect
            K[1, 1] = f / 2 * (1 + aspect ** 2) ** 0.5
            K[0, 2] = K[1, 2] = 0.5
            K[2, 2] = 1.0
            cam_pts = pts.clone()
            cam_pts[..., 2] += float(shift[0])
            if "log_scale" in o:
                cam_pts *= o["log_scale"].exp()
            for key, val in (("points", cam_pts), ("intrinsics", K), ("depth", depth)):
                res.setdefault(key, []).append(val)
            if "mask" in o:
                res.setdefault("mask", []).append(mask.bool())
                if apply_mask:
                    cam_pts[~res["mask"][-1]] = float("inf")
        out = {k: torch.stack(v) for k, v in res.items()}
        if omit:
            out = {k: v.squeeze(0) for k, v in out.items()}
        return out
