"""`Pi3` — drop-in for the reference model class on the chunk-creation path.

Mirrors reference `pi3/models/pi3.py`: same constructor arguments, same 1210 `state_dict` keys
(so reference checkpoints load with strict=True), `from_pretrained(dir_or_repo)`, and
`forward(imgs (B,N,3,H,W) in [0,1]) -> dict(points, local_points, conf, camera_poses)` all fp32
(pi3.py:173-216). The computation itself is a flat schedule of C-ABI kernel launches
(libpi3b200.so): there is no nn.Linear / SDPA / autocast anywhere on this path and no CPU fallback.

Numerics: bf16 tensor-core operands with fp32 accumulation and an fp32 residual stream everywhere
(also in the three head decoders, where the reference's autocast path degrades to a bf16 residual),
fp32 RoPE angles, TF32 tensor cores for the fp32 heads (the reference enables TF32 for them,
create_offline_chunks.py:22), fp64 for the 3x3 pose orthogonalisation.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .synthetic import pi3_param_shapes

PATCH = 14
HEADS = 16
N_SPECIAL = 5
D = 1024
K_PATCH = 3 * PATCH * PATCH      # 588
K_PATCH_PAD = 592                # 16-byte row pitch for TMA


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def _set_nested(root: nn.Module, key: str, tensor: torch.Tensor, buffer: bool) -> None:
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class Pi3(nn.Module):
    def __init__(self, pos_type: str = "rope100", decoder_size: str = "large"):
        super().__init__()
        if pos_type is None or not pos_type.startswith("rope"):
            raise NotImplementedError("only ropeNNN position encodings exist in the reference (pi3.py:34-42)")
        if decoder_size != "large":
            raise NotImplementedError("the B200 kernels are built for decoder_size='large' (D=1024, 16 heads)")
        self.pos_type = pos_type
        self.rope_freq = float(pos_type[len("rope"):])
        self.patch_size = PATCH
        self.patch_start_idx = N_SPECIAL
        self.dec_embed_dim = D
        for key, shape in pi3_param_shapes().items():
            is_buf = key in ("image_mean", "image_std")
            _set_nested(self, key, torch.zeros(shape), buffer=is_buf)
        with torch.no_grad():
            self.image_mean.copy_(torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
            self.image_std.copy_(torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        self._packed: Optional[dict] = None       # bf16 / padded operands, built lazily on the device
        self._shape_cache: Dict[Tuple[int, int], dict] = {}
        self._arena: Dict[tuple, dict] = {}
        # when set to a list, (start, end) CUDA events of every chunk-global attention launch are appended
        # (bench.py uses them for the live roofline of the dominant kernel)
        self.global_attn_events: Optional[list] = None
        # sequence-parallel chunk-global layers: True = attend to the local K/V block while the remote blocks cross
        # NVLink, then merge through the log-sum-exp (exposed all-gather ~0, result within bf16 tolerance of the
        # single-GPU forward); False = blocking all-gather, one attention over all keys (bit-identical to one GPU)
        self.sp_overlap: bool = True
        # q/k LayerNorm + RoPE fused into the QKV GEMM epilogue (bit-identical to the separate pi3b_qknorm_rope pass, which
        # PI3B_FUSE_QK=0 brings back for A/B runs)
        self.fuse_qk_epilogue: bool = os.environ.get("PI3B_FUSE_QK", "1") != "0"
        # softmax scale * log2(e) folded into q where q is post-processed anyway (decoder and head-decoder blocks): one
        # rounding of q * c instead of q, and the chunk-global attention kernel skips the scale-and-subtract per score
        # (PI3B_PRESCALE_Q=0 for A/B runs and for the parity test that compares both)
        self.prescale_q: bool = os.environ.get("PI3B_PRESCALE_Q", "1") != "0"
        # when set to a list, (local_done, gather_done) CUDA events of every sequence-parallel layer are appended:
        # their distance is the all-gather time the local attention did not hide
        self.sp_events: Optional[list] = None

    # ------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, path_or_repo: str, **kwargs) -> "Pi3":
        """Local directory with config.json + model.safetensors (what the reference's
        PyTorchModelHubMixin.save_pretrained writes), a .safetensors/.pt file, or a hub repo id."""
        path = path_or_repo
        if not os.path.exists(path):
            from huggingface_hub import snapshot_download  # no network on the build boxes: raises there
            path = snapshot_download(path_or_repo)
        cfg = {}
        weights = path
        if os.path.isdir(path):
            cfg_path = os.path.join(path, "config.json")
            if os.path.exists(cfg_path):
                with open(cfg_path) as f:
                    cfg = json.load(f)
            weights = os.path.join(path, "model.safetensors")
        ctor = {k: cfg[k] for k in ("pos_type", "decoder_size") if k in cfg}
        ctor.update(kwargs)
        model = cls(**ctor)
        if weights.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(weights)
        else:
            sd = torch.load(weights, map_location="cpu")
            if isinstance(sd, dict) and "state_dict" in sd:
                sd = sd["state_dict"]
        # the two normalisation buffers are not always serialised by older checkpoints; nothing else may be absent
        res = model.load_state_dict(sd, strict=False)
        missing = [k for k in res.missing_keys if k not in ("image_mean", "image_std")]
        if missing or res.unexpected_keys:
            raise RuntimeError(f"checkpoint does not match Pi3: missing {missing[:8]} unexpected {res.unexpected_keys[:8]}")
        return model

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._invalidate()
        return out

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._invalidate()
        return out

    def _invalidate(self) -> None:
        self._packed = None
        self._shape_cache = {}
        self._arena = {}

    # ------------------------------------------------------------------ packing
    def _p(self, key: str) -> torch.Tensor:
        mod = self
        for part in key.split("."):
            mod = getattr(mod, part)
        return mod

    def _pack(self, device: torch.device) -> dict:
        if self._packed is not None and self._packed["device"] == device:
            return self._packed
        _lib.load()  # fail loudly before touching anything if the CUDA library is missing
        pk: dict = {"device": device}

        def f32(key):
            return self._p(key).detach().to(device=device, dtype=torch.float32).contiguous()

        def bf16(key):
            return self._p(key).detach().to(device=device, dtype=torch.bfloat16).contiguous()

        def block(prefix, ls, qkn):
            b = {
                "n1": (f32(prefix + ".norm1.weight"), f32(prefix + ".norm1.bias")),
                "n2": (f32(prefix + ".norm2.weight"), f32(prefix + ".norm2.bias")),
                "qkv_w": bf16(prefix + ".attn.qkv.weight"), "qkv_b": f32(prefix + ".attn.qkv.bias"),
                "proj_w": bf16(prefix + ".attn.proj.weight"), "proj_b": f32(prefix + ".attn.proj.bias"),
                "fc1_w": bf16(prefix + ".mlp.fc1.weight"), "fc1_b": f32(prefix + ".mlp.fc1.bias"),
                "fc2_w": bf16(prefix + ".mlp.fc2.weight"), "fc2_b": f32(prefix + ".mlp.fc2.bias"),
                "ls1": f32(prefix + ".ls1.gamma") if ls else None,
                "ls2": f32(prefix + ".ls2.gamma") if ls else None,
                "qn": (f32(prefix + ".attn.q_norm.weight"), f32(prefix + ".attn.q_norm.bias")) if qkn else None,
                "kn": (f32(prefix + ".attn.k_norm.weight"), f32(prefix + ".attn.k_norm.bias")) if qkn else None,
            }
            # fused QKV epilogue: [q weight | q bias | k weight | k bias] x 64
            b["qkn"] = torch.cat([b["qn"][0], b["qn"][1], b["kn"][0], b["kn"][1]]).contiguous() if qkn else None
            return b

        wp = torch.zeros(D, K_PATCH_PAD, device=device, dtype=torch.bfloat16)
        wp[:, :K_PATCH] = self._p("encoder.patch_embed.proj.weight").detach().to(device).reshape(D, K_PATCH).to(torch.bfloat16)
        pk["patch_w"] = wp
        pk["patch_b"] = f32("encoder.patch_embed.proj.bias")
        pk["enc"] = [block(f"encoder.blocks.{i}", True, False) for i in range(24)]
        pk["enc_norm"] = (f32("encoder.norm.weight"), f32("encoder.norm.bias"))
        pk["dec"] = [block(f"decoder.{i}", True, True) for i in range(36)]
        pk["dec_reg"] = f32("register_token").reshape(N_SPECIAL, D).contiguous()
        for name in ("point_decoder", "conf_decoder", "camera_decoder"):
            pk[name] = {
                "projects_w": bf16(name + ".projects.weight"), "projects_b": f32(name + ".projects.bias"),
                "blocks": [block(f"{name}.blocks.{i}", False, False) for i in range(5)],
                "out_w": bf16(name + ".linear_out.weight"), "out_b": f32(name + ".linear_out.bias"),
            }
        pk["point_head"] = (f32("point_head.proj.weight"), f32("point_head.proj.bias"))
        pk["conf_head"] = (f32("conf_head.proj.weight"), f32("conf_head.proj.bias"))
        cam = {}
        for i in range(2):
            for j in (1, 2, 3):
                cam[f"res{i}{j}"] = (f32(f"camera_head.res_conv.{i}.res_conv{j}.weight"),
                                     f32(f"camera_head.res_conv.{i}.res_conv{j}.bias"))
        for i in (0, 2):
            cam[f"mlp{i}"] = (f32(f"camera_head.more_mlps.{i}.weight"), f32(f"camera_head.more_mlps.{i}.bias"))
        cam["fc_t"] = (f32("camera_head.fc_t.weight"), f32("camera_head.fc_t.bias"))
        cam["fc_rot"] = (f32("camera_head.fc_rot.weight"), f32("camera_head.fc_rot.bias"))
        pk["cam"] = cam
        pk["mean"] = self.image_mean.detach().flatten().tolist()
        pk["std"] = self.image_std.detach().flatten().tolist()
        self._packed = pk
        return pk

    def _shape_consts(self, H: int, W: int, device: torch.device) -> dict:
        """Per-(H, W) constants: token table (cls+pos / registers / interpolated pos-embed) and RoPE table.
        The bicubic-antialias resize of the pos-embed grid (vision_transformer.py:181-213) is a per-shape
        constant, evaluated once here with the same torch op the reference uses."""
        key = (H, W)
        c = self._shape_cache.get(key)
        if c is not None and c["device"] == device:
            return c
        ph, pw = H // PATCH, W // PATCH
        pos_embed = self._p("encoder.pos_embed").detach().float().cpu()
        n = pos_embed.shape[1] - 1
        m = int(math.sqrt(n))
        if ph * pw == n and ph == pw:
            patch_pos = pos_embed[0, 1:]
        else:
            patch_pos = F.interpolate(pos_embed[:, 1:].reshape(1, m, m, D).permute(0, 3, 1, 2), mode="bicubic",
                                      antialias=True, size=(ph, pw)).permute(0, 2, 3, 1).reshape(-1, D)
        table = torch.empty(N_SPECIAL + ph * pw, D)
        table[0] = self._p("encoder.cls_token").detach().float().cpu().reshape(D) + pos_embed[0, 0]
        table[1:N_SPECIAL] = self._p("encoder.register_tokens").detach().float().cpu().reshape(4, D)
        table[N_SPECIAL:] = patch_pos
        # RoPE table: theta = pos * base^(-j/16), fp32 (pos_embed.py:120-129 evaluated in fp32)
        max_pos = max(ph, pw)
        inv_freq = 1.0 / (self.rope_freq ** (torch.arange(0, 32, 2).float() / 32))
        ang = torch.einsum("i,j->ij", torch.arange(max_pos + 1, dtype=torch.float32), inv_freq)
        cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous()
        c = {"device": device, "table": table.to(device).contiguous(), "cs": cs.to(device), "ph": ph, "pw": pw}
        self._shape_cache[key] = c
        return c

    def _buffers_for(self, BN: int, H: int, W: int, device: torch.device, sp_world: int = 1, sp_max_frames: int = 0) -> dict:
        key = (BN, H, W, str(device), sp_world, sp_max_frames)
        a = self._arena.get(key)
        if a is not None:
            return a
        P = (H // PATCH) * (W // PATCH)
        T = P + N_SPECIAL
        M = BN * T
        e = lambda *s, dt=torch.float32: torch.empty(*s, device=device, dtype=dt)  # noqa: E731
        bf = torch.bfloat16
        a = {
            "patches": e(BN * P, K_PATCH_PAD, dt=bf),
            "x": e(M, D), "h": e(M, D, dt=bf), "qkv": e(M, 3 * D, dt=bf), "att": e(M, D, dt=bf),
            "mlp": e(M, 4 * D, dt=bf), "cat": e(M, 2 * D, dt=bf), "xh": e(M, D),
            "point_hidden": e(M, D), "conf_hidden": e(M, D), "camera_hidden": e(M, 512),
            "cam_t1": e(M, 512), "cam_t2": e(M, 512),
            "feat_pts": e(M, K_PATCH), "feat_conf": e(M, PATCH * PATCH),
            "pooled": e(BN, 512), "pooled2": e(BN, 512), "rot9": e(BN, 12), "trans": e(BN, 4),
        }
        if sp_world > 1:
            m_max = max(sp_max_frames, BN) * T          # rows of the largest rank's block (uneven frame splits are padded)
            a["kv_send"] = e(m_max, 2 * D, dt=bf)
            a["kv_all"] = e(sp_world * m_max, 2 * D, dt=bf)
            a["att_remote"] = e(M, D, dt=bf)
            a["lse"] = e(2, HEADS, M)
        self._arena = {key: a}   # one live shape at a time
        return a

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _sp_plan(group, frame_counts, T: int) -> dict:
        """Per-forward constants of the sequence-parallel chunk-global layers. Every rank's K|V block lands at
        offset g * m_max of the gathered buffer (blocks of ranks with fewer frames are padded). Softmax attention does
        not care about the ORDER of the keys, so the remote rows are made contiguous by moving rows from the end of the
        buffer into the holes (this rank's own block, pad rows): `copies` = (dst, src, n) row ranges."""
        import torch.distributed as dist
        G = dist.get_world_size(group)
        r = dist.get_rank(group)
        rows = [int(c) * T for c in frame_counts]
        m_max = max(rows)
        valid = [(g * m_max, rows[g]) for g in range(G) if g != r and rows[g] > 0]
        R = sum(n for _, n in valid)
        holes, movers, pos = [], [], 0
        for s0, n in valid:
            lo, hi = s0, s0 + n
            if lo >= R:
                movers.append((lo, n))
                continue
            if lo > pos:
                holes.append((pos, lo - pos))
            if hi > R:
                movers.append((R, hi - R))
                hi = R
            pos = hi
        if pos < R:
            holes.append((pos, R - pos))
        copies = []
        movers = movers[::-1]
        for dst, need in holes:
            while need > 0:
                src, have = movers[-1]
                n = min(need, have)
                copies.append((dst, src + have - n, n))       # take from the end of the mover
                if n == have:
                    movers.pop()
                else:
                    movers[-1] = (src, have - n)
                dst += n
                need -= n
        assert not movers
        return dict(group=group, G=G, rank=r, rows=rows, m_max=m_max, remote_rows=R, copies=copies,
                    all_rows=sum(rows), equal=len(set(rows)) == 1)

    def _sp_global_attention(self, qkv, att, bufs, sp: dict, M: int, scale: float = 0.125) -> None:
        """Chunk-global attention of the M local query rows against the K/V rows of every rank of the group
        (`att` = softmax(q K^T / 8) V over all sum(rows) keys). K|V cross NVLink as one NCCL all-gather per layer."""
        import torch.distributed as dist
        kv_send, kv_all = bufs["kv_send"], bufs["kv_all"]
        kv_send[:M].copy_(qkv[:, D:])                     # k | v of the local frames, made contiguous
        if not self.sp_overlap:
            # blocking variant: one attention over the gathered keys in frame order (bit-identical to a single GPU);
            # needs equal blocks so that the gathered buffer has no pad rows inside the sequence
            if not sp["equal"]:
                raise ValueError("sp_overlap=False needs the same number of frames on every rank")
            dist.all_gather_into_tensor(kv_all, kv_send, group=sp["group"])
            ops.attention_q_kv(qkv, kv_all, att, 1, M, sp["all_rows"], HEADS, scale)
            return
        work = dist.all_gather_into_tensor(kv_all, kv_send, group=sp["group"], async_op=True)   # NCCL's own stream
        lse = bufs["lse"]
        ops.attention_q_kv_lse(qkv, kv_send[:M], att, lse[0], 1, M, M, HEADS, scale)             # local block meanwhile
        if self.sp_events is not None:
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
        work.wait()                                       # the current stream waits for the gathered rows
        if self.sp_events is not None:
            e1 = torch.cuda.Event(enable_timing=True); e1.record()
            self.sp_events.append((e0, e1))
        for dst, src, n in sp["copies"]:
            kv_all[dst:dst + n].copy_(kv_all[src:src + n])
        R = sp["remote_rows"]
        ops.attention_q_kv_lse(qkv, kv_all[:R], bufs["att_remote"], lse[1], 1, M, R, HEADS, scale)
        ops.attention_merge([att, bufs["att_remote"]], [lse[0], lse[1]], att, HEADS)

    def _block(self, x, b, bufs, num_seqs, seq_len, T, pw, cs, rope: bool, sp=None) -> None:
        """x += ls1 * proj(attn(LN(x)));  x += ls2 * fc2(GELU(fc1(LN(x))))   (block.py:333-334), in place on x.
        With `sp` (sequence-parallel chunk-global layer, see `_sp_plan`) the local queries attend to the K/V rows of
        every rank of the group, all-gathered over NCCL."""
        h, qkv, att, mlp = bufs["h"], bufs["qkv"], bufs["att"], bufs["mlp"]
        ops.layernorm(x, b["n1"][0], b["n1"][1], 1e-6, h)
        # blocks whose q passes through a post-processing step anyway (q/k norm and / or RoPE) fold the softmax scale
        # into it: exp2 then applies to the attention accumulator as is (see ops.Q_PRESCALE)
        prescale = self.prescale_q and (rope or b["qn"] is not None)
        q_scale, scale = (ops.Q_PRESCALE, ops.ATTN_SCALE_PRESCALED) if prescale else (1.0, 0.125)
        if (rope or b["qn"] is not None) and self.fuse_qk_epilogue:
            # q/k LayerNorm(64) + RoPE-2D inside the QKV GEMM's epilogue (one launch, no extra pass over q and k)
            ops.gemm_qkv(h, b["qkv_w"], qkv, b["qkv_b"], HEADS, b["qkn"], 1e-5, cs if rope else None, T, N_SPECIAL, pw,
                         q_scale)
        else:
            ops.gemm(h, b["qkv_w"], qkv, bias=b["qkv_b"])
            if rope or b["qn"] is not None:
                ops.qknorm_rope(qkv, HEADS, T, N_SPECIAL, pw, cs if rope else None, b["qn"], b["kn"], 1e-5, q_scale)
        if sp is not None:
            self._sp_global_attention(qkv, att, bufs, sp, seq_len, scale)
        elif self.global_attn_events is not None and seq_len != T:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.attention_qkv(qkv, att, num_seqs, seq_len, HEADS, scale)
            e1.record()
            self.global_attn_events.append((e0, e1))
        else:
            ops.attention_qkv(qkv, att, num_seqs, seq_len, HEADS, scale)
        ops.gemm(att, b["proj_w"], x, bias=b["proj_b"], gamma=b["ls1"], resid=x)
        ops.layernorm(x, b["n2"][0], b["n2"][1], 1e-6, h)
        ops.gemm(h, b["fc1_w"], mlp, bias=b["fc1_b"], act=ops.ACT_GELU)
        ops.gemm(mlp, b["fc2_w"], x, bias=b["fc2_b"], gamma=b["ls2"], resid=x)

    @torch.no_grad()
    def forward(self, imgs: torch.Tensor, taps: Optional[dict] = None, sp_group=None,
                sp_frame_counts=None) -> Dict[str, torch.Tensor]:
        """Reference signature `forward(imgs)`. Extension for over-long chunks: with `sp_group` (a
        torch.distributed group, NCCL) `imgs` holds this rank's contiguous slice of the chunk's frames (rank order =
        frame order, B = 1); frame-local layers, MLPs and heads run locally, each of the 18 chunk-global layers
        all-gathers bf16 K and V across the group. `sp_frame_counts`: frames held by each rank of the group (default:
        the same number everywhere). Returns the local frames' outputs."""
        if imgs.dim() != 5 or imgs.shape[2] != 3:
            raise ValueError("imgs must be (B, N, 3, H, W)")
        sp_world = 1
        if sp_group is not None:
            import torch.distributed as dist
            sp_world = dist.get_world_size(sp_group)
            if imgs.shape[0] != 1:
                raise ValueError("sequence-parallel forward needs B == 1")
            if sp_world == 1:
                sp_group = None
        if not imgs.is_cuda:
            raise _lib.Pi3bError("pi3_slam_b200.Pi3 runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, N, _, H, W = imgs.shape
        if H % PATCH or W % PATCH:
            raise ValueError("H and W must be multiples of 14")
        # the C ABI launches on the CURRENT device and stream: make the frames' device current for the whole schedule
        with ops.on_device_of(imgs):
            return self._forward_on_device(imgs, taps, sp_group, sp_world, sp_frame_counts)

    def _forward_on_device(self, imgs, taps, sp_group, sp_world, sp_frame_counts) -> Dict[str, torch.Tensor]:
        B, N, _, H, W = imgs.shape
        device = imgs.device
        pk = self._pack(device)
        sc = self._shape_consts(H, W, device)
        ph, pw = sc["ph"], sc["pw"]
        P = ph * pw
        T = P + N_SPECIAL
        BN = B * N
        M = BN * T
        sp = None
        if sp_group is not None:
            counts = list(sp_frame_counts) if sp_frame_counts is not None else [N] * sp_world
            if len(counts) != sp_world:
                raise ValueError("sp_frame_counts must have one entry per rank of sp_group")
            sp = self._sp_plan(sp_group, counts, T)
            if sp["rows"][sp["rank"]] != M:
                raise ValueError("sp_frame_counts disagrees with the number of local frames")
        bufs = self._buffers_for(BN, H, W, device, sp_world, max(sp_frame_counts) if sp_frame_counts is not None else 0)
        x = bufs["x"]
        imgs_f = imgs.reshape(BN, 3, H, W).to(torch.float32).contiguous()

        # ---- encoder tokens: patch-embed GEMM accumulated onto the (cls | registers | pos-embed) table
        ops.patch_pack(imgs_f, bufs["patches"], pk["mean"], pk["std"])
        ops.broadcast_rows(sc["table"], x, BN)
        ops.gemm(bufs["patches"], pk["patch_w"], x, bias=pk["patch_b"], resid=x, row_remap=(P, T, N_SPECIAL))
        if taps is not None:
            taps["enc_tokens"] = x.view(BN, T, D).clone()
        # ---- DINOv2 encoder: 24 frame-local blocks
        for i, b in enumerate(pk["enc"]):
            self._block(x, b, bufs, BN, T, T, pw, None, rope=False)
            if taps is not None and i == 0:
                taps["enc_block0"] = x.view(BN, T, D).clone()
        # final norm; cls/register rows are dropped and replaced by the decoder's 5 register tokens
        ops.layernorm(x, pk["enc_norm"][0], pk["enc_norm"][1], 1e-6, x, override_rows=pk["dec_reg"], period=T)
        if taps is not None:
            taps["enc_out"] = x.view(BN, T, D)[:, N_SPECIAL:].clone()
        # ---- decoder: 36 blocks alternating frame-local / chunk-global attention (pi3.py:156-166)
        cat = bufs["cat"]
        for i, b in enumerate(pk["dec"]):
            if i % 2 == 0:
                self._block(x, b, bufs, BN, T, T, pw, sc["cs"], rope=True)
            else:
                self._block(x, b, bufs, B, N * T, T, pw, sc["cs"], rope=True, sp=sp)
            if taps is not None and i in (0, 1):
                taps[f"dec_block{i}"] = x.view(BN, T, D).clone()
            if i == 34:
                ops.cast_bf16(x, cat[:, :D])
            elif i == 35:
                ops.cast_bf16(x, cat[:, D:])
        if taps is not None:
            taps["dec_out"] = cat.view(BN, T, 2 * D).float()
        # ---- three head decoders (transformer_head.py:48-56), frame-local attention with RoPE
        xh = bufs["xh"]
        for name, out_key in (("point_decoder", "point_hidden"), ("conf_decoder", "conf_hidden"),
                              ("camera_decoder", "camera_hidden")):
            hd = pk[name]
            ops.gemm(cat, hd["projects_w"], xh, bias=hd["projects_b"])
            for b in hd["blocks"]:
                self._block(xh, b, bufs, BN, T, T, pw, sc["cs"], rope=True)
            ops.cast_bf16(xh, bufs["h"])
            ops.gemm(bufs["h"], hd["out_w"], bufs[out_key], bias=hd["out_b"])
        if taps is not None:
            taps["point_hidden"] = bufs["point_hidden"].view(BN, T, D).clone()
            taps["camera_hidden"] = bufs["camera_hidden"].view(BN, T, 512).clone()
        # ---- fp32 heads on TF32 tensor cores (pi3.py:192-209)
        ops.gemm(bufs["point_hidden"], pk["point_head"][0], bufs["feat_pts"], bias=pk["point_head"][1])
        ops.gemm(bufs["conf_hidden"], pk["conf_head"][0], bufs["feat_conf"], bias=pk["conf_head"][1])
        cam = pk["cam"]
        f = bufs["camera_hidden"]
        for i in range(2):
            ops.gemm(f, cam[f"res{i}1"][0], bufs["cam_t1"], bias=cam[f"res{i}1"][1], act=ops.ACT_RELU)
            ops.gemm(bufs["cam_t1"], cam[f"res{i}2"][0], bufs["cam_t2"], bias=cam[f"res{i}2"][1], act=ops.ACT_RELU)
            ops.gemm(bufs["cam_t2"], cam[f"res{i}3"][0], f, bias=cam[f"res{i}3"][1], act=ops.ACT_RELU, resid=f)
        ops.meanpool_tokens(f, bufs["pooled"], BN, T, N_SPECIAL)
        ops.gemm(bufs["pooled"], cam["mlp0"][0], bufs["pooled2"], bias=cam["mlp0"][1], act=ops.ACT_RELU)
        ops.gemm(bufs["pooled2"], cam["mlp2"][0], bufs["pooled"], bias=cam["mlp2"][1], act=ops.ACT_RELU)
        ops.gemm(bufs["pooled"], cam["fc_t"][0], bufs["trans"][:, :3], bias=cam["fc_t"][1])
        ops.gemm(bufs["pooled"], cam["fc_rot"][0], bufs["rot9"][:, :9], bias=cam["fc_rot"][1])
        poses = torch.empty(BN, 4, 4, device=device, dtype=torch.float32)
        ops.pose_from_9d(bufs["rot9"], bufs["trans"], poses)
        local_points = torch.empty(BN, H, W, 3, device=device, dtype=torch.float32)
        points = torch.empty(BN, H, W, 3, device=device, dtype=torch.float32)
        conf = torch.empty(BN, H, W, 1, device=device, dtype=torch.float32)
        ops.unpatchify_points_conf(bufs["feat_pts"], bufs["feat_conf"], poses, local_points, points, conf, BN, H, W, T,
                                   N_SPECIAL)
        return dict(
            points=points.view(B, N, H, W, 3),
            local_points=local_points.view(B, N, H, W, 3),
            conf=conf.view(B, N, H, W, 1),
            camera_poses=poses.view(B, N, 4, 4),
        )

    # algorithmic FLOPs of one forward — used by bench.py for the roofline
    @staticmethod
    def flops(N: int, H: int, W: int) -> Dict[str, float]:
        P = (H // PATCH) * (W // PATCH)
        T = P + N_SPECIAL
        lin = 24.0 * D * D
        f = {
            "patch": 2.0 * N * P * K_PATCH * D,
            "trunk_linear": 60.0 * N * T * lin,
            "frame_attn": (24 + 18) * N * 4.0 * T * T * D,
            "global_attn": 18 * 4.0 * (N * T) ** 2 * D,
            "head_decoders": N * T * (3 * 2.0 * 2048 * D + 15 * lin + 2.0 * D * (1024 + 1024 + 512)) + 15 * N * 4.0 * T * T * D,
            "heads": N * P * (2.0 * D * (588 + 196) + 6 * 2.0 * 512 * 512),
        }
        f["total"] = sum(f.values())
        return f
