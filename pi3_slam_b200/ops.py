"""Host-side launchers: torch tensors in, raw pointers + shapes through the C ABI, nothing else.

Each function mirrors one entry point of include/pi3b200.h. Tensors must already live on the CUDA
device; the kernels are enqueued on torch's current stream and never synchronise.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
# Softmax scale folded into the queries: q leaves the QKV epilogue multiplied by (1/8) * log2(e) and the attention is
# called with scale = ln 2, i.e. softmax_2(q' k) == softmax_e(q k / 8). The long-sequence kernel then takes the
# exponent straight from the accumulator (attn.cu, "zero reference").
Q_PRESCALE = 0.125 * 1.4426950408889634
ATTN_SCALE_PRESCALED = 0.6931471805599453


def device_guard(get_device):
    """Method decorator: run the method with `get_device(self, *args)` (a CUDA torch.device / str / tensor) as the
    current CUDA device. Every C-ABI launch inside then goes to that device's stream, function attributes and SM count."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(self, *args, **kwargs):
            d = get_device(self, *args)
            if torch.is_tensor(d):
                d = d.device
            d = torch.device(d) if d is not None else None
            if d is None or d.type != "cuda":
                return fn(self, *args, **kwargs)      # the method itself rejects non-CUDA inputs loudly
            with torch.cuda.device(d):
                return fn(self, *args, **kwargs)
        return wrapper
    return deco


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def on_device_of(t: torch.Tensor):
    """Context manager that makes `t`'s device current. The C ABI launches on the CURRENT device (the stream handle,
    the opt-in shared-memory attributes and the SM count are all per device), so callers that may be handed tensors of
    another GPU (`OfflineCreatorConfig.device='cuda:1'`, `Pi3.forward` on a non-default device) wrap their launch
    sequence in this once instead of paying a device switch per kernel."""
    if not t.is_cuda:
        raise _lib.Pi3bError("expected a CUDA tensor (no CPU fallback)")
    return torch.cuda.device(t.device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.Pi3bError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.Pi3bError(f"{name} must be {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, bias: Optional[torch.Tensor] = None,
         gamma: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         k: Optional[int] = None, row_remap: Optional[tuple] = None) -> torch.Tensor:
    """out[M,N] = [resid +] gamma * act(a[M,K] @ w[N,K]^T + bias).

    a, w: bf16 (tcgen05 kind::f16) or fp32 (read as TF32); row-major with unit inner stride, row pitch
    a multiple of 16 bytes. out: bf16 or fp32. bias/gamma: fp32 [N]. resid: fp32 (may alias out).
    `k` overrides the contraction length (for K-padded operands). `row_remap=(grp_in, grp_out, off)`
    scatters GEMM row m to row (m // grp_in) * grp_out + off + m % grp_in of out/resid.
    """
    lib = _lib.load()
    assert a.dim() == 2 and w.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    assert a.dtype == w.dtype and a.dtype in (torch.bfloat16, torch.float32)
    _req(a, a.dtype, "a"); _req(w, a.dtype, "w")
    M = a.shape[0]
    N = w.shape[0]
    K = k if k is not None else a.shape[1]
    grp = row_remap if row_remap is not None else (0, 0, 0)
    m_out = M if row_remap is None else (M // grp[0]) * grp[1]
    assert out.shape[0] == m_out and out.shape[1] == N
    assert out.dtype in (torch.bfloat16, torch.float32) and out.is_cuda
    for t, nm in ((bias, "bias"), (gamma, "gamma"), (resid, "resid")):
        if t is not None:
            _req(t, torch.float32, nm)
    rc = lib.pi3b_gemm(_stream(), 0 if a.dtype == torch.bfloat16 else 1, a.data_ptr(), a.stride(0), w.data_ptr(),
                       w.stride(0), M, N, K, _ptr(bias), _ptr(gamma), _ptr(resid),
                       resid.stride(0) if resid is not None else 0, out.data_ptr(), out.stride(0),
                       1 if out.dtype == torch.float32 else 0, act, grp[0], grp[1], grp[2])
    _lib.check(rc, "pi3b_gemm")
    return out


_ATTN_WS = {}


def _attn_workspace(device: torch.device) -> torch.Tensor:
    """Scratch buffer of the attention launches (max |k|^2 per (sequence, head) for the bounded-reference softmax): one per
    (device, stream), because launches on different streams may run concurrently."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ATTN_WS.get(key)
    if ws is None:
        ws = torch.zeros(16384, device=device, dtype=torch.float32)
        _ATTN_WS[key] = ws
    return ws


def gemm_qkv(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, bias: Optional[torch.Tensor], num_heads: int,
             qk_norm: Optional[torch.Tensor], eps: float, cs_table: Optional[torch.Tensor], tokens_per_frame: int,
             n_special: int, patch_w: int, q_scale: float = 1.0) -> torch.Tensor:
    """QKV projection with q/k LayerNorm(64) and RoPE-2D fused into the GEMM epilogue (bit-identical to `gemm` followed by
    `qknorm_rope`). a [M,K], w [3*64*heads, K] bf16; out [M, 3*64*heads] bf16; qk_norm fp32 [4,64] (q weight, q bias,
    k weight, k bias) or None; cs_table fp32 [max_pos+1, 16, 2] or None. q_scale: the queries are multiplied by it
    before their final rounding (`Q_PRESCALE` with attention `scale=ATTN_SCALE_PRESCALED`: the softmax scale rides on q)."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w"); _req(out, torch.bfloat16, "out")
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    M, K = a.shape
    assert w.shape == (3 * 64 * num_heads, K) and out.shape == (M, 3 * 64 * num_heads)
    rows = 0
    if qk_norm is not None:
        _req(qk_norm, torch.float32, "qk_norm")
        assert qk_norm.is_contiguous() and qk_norm.numel() == 256
    if cs_table is not None:
        _req(cs_table, torch.float32, "cs_table")
        assert cs_table.is_contiguous() and cs_table.shape[1:] == (16, 2)
        rows = cs_table.shape[0]
    if bias is not None:
        _req(bias, torch.float32, "bias")
    rc = lib.pi3b_gemm_qkv(_stream(), a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, num_heads, K, _ptr(bias),
                           out.data_ptr(), out.stride(0), _ptr(qk_norm), eps, _ptr(cs_table), rows, tokens_per_frame,
                           n_special, patch_w, q_scale)
    _lib.check(rc, "pi3b_gemm_qkv")
    return out


def attention_qkv(qkv: torch.Tensor, out: torch.Tensor, num_seqs: int, seq_len: int, num_heads: int = 16,
                  scale: float = 0.125) -> torch.Tensor:
    """softmax(q k^T * scale) v per (sequence, head), head_dim 64.

    qkv: bf16 [num_seqs*seq_len, 3*num_heads*64] packed as the qkv Linear emits it (q | k | v, head-major
    inside each third — reference layout, pi3/models/layers/attention.py:326-329). out: bf16
    [num_seqs*seq_len, num_heads*64].
    """
    lib = _lib.load()
    _req(qkv, torch.bfloat16, "qkv"); _req(out, torch.bfloat16, "out")
    C = num_heads * 64
    assert qkv.dim() == 2 and qkv.shape[1] == 3 * C and qkv.stride(1) == 1
    assert qkv.shape[0] == num_seqs * seq_len and out.shape[0] == qkv.shape[0] and out.shape[1] == C
    ld = qkv.stride(0)
    ws = _attn_workspace(qkv.device)
    rc = lib.pi3b_attn_fwd_hd64_ws(_stream(), qkv.data_ptr(), qkv.data_ptr(), qkv.data_ptr(), ld, ld, ld, 0, C, 2 * C,
                                   out.data_ptr(), out.stride(0), None, num_seqs, seq_len, seq_len, num_heads, scale,
                                   ws.data_ptr(), ws.numel() * 4)
    _lib.check(rc, "pi3b_attn_fwd_hd64_ws")
    return out


def attention_q_kv(q: torch.Tensor, kv: torch.Tensor, out: torch.Tensor, num_seqs: int, q_len: int, kv_len: int,
                   num_heads: int = 16, scale: float = 0.125, q_col0: int = 0) -> torch.Tensor:
    """Attention of local queries against a separate key/value matrix (sequence-parallel global attention):
    q: bf16 [num_seqs*q_len, >= q_col0 + heads*64] (queries at column q_col0 + 64*h), kv: bf16
    [num_seqs*kv_len, 2*heads*64] (k | v, head-major), out: bf16 [num_seqs*q_len, heads*64]."""
    lib = _lib.load()
    _req(q, torch.bfloat16, "q"); _req(kv, torch.bfloat16, "kv"); _req(out, torch.bfloat16, "out")
    C = num_heads * 64
    assert q.stride(1) == 1 and kv.stride(1) == 1 and out.stride(1) == 1
    assert q.shape[0] == num_seqs * q_len and kv.shape == (num_seqs * kv_len, 2 * C) and out.shape == (q.shape[0], C)
    ws = _attn_workspace(q.device)
    rc = lib.pi3b_attn_fwd_hd64_ws(_stream(), q.data_ptr(), kv.data_ptr(), kv.data_ptr(), q.stride(0), kv.stride(0),
                                   kv.stride(0), q_col0, 0, C, out.data_ptr(), out.stride(0), None, num_seqs, q_len, kv_len,
                                   num_heads, scale, ws.data_ptr(), ws.numel() * 4)
    _lib.check(rc, "pi3b_attn_fwd_hd64_ws")
    return out


def attention_q_kv_lse(q: torch.Tensor, kv: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, num_seqs: int,
                       q_len: int, kv_len: int, num_heads: int = 16, scale: float = 0.125, q_col0: int = 0) -> None:
    """`attention_q_kv` that also writes the log2-domain log-sum-exp of every (head, query row) into `lse`
    (fp32 [num_heads, num_seqs*q_len]) so that results over disjoint key/value sets can be merged."""
    lib = _lib.load()
    _req(q, torch.bfloat16, "q"); _req(kv, torch.bfloat16, "kv"); _req(out, torch.bfloat16, "out")
    _req(lse, torch.float32, "lse")
    C = num_heads * 64
    assert q.stride(1) == 1 and kv.stride(1) == 1 and out.stride(1) == 1 and lse.is_contiguous()
    assert q.shape[0] == num_seqs * q_len and kv.shape[0] == num_seqs * kv_len and kv.shape[1] == 2 * C
    assert out.shape == (q.shape[0], C) and lse.shape == (num_heads, q.shape[0])
    ws = _attn_workspace(q.device)
    rc = lib.pi3b_attn_fwd_hd64_ws(_stream(), q.data_ptr(), kv.data_ptr(), kv.data_ptr(), q.stride(0), kv.stride(0),
                                   kv.stride(0), q_col0, 0, C, out.data_ptr(), out.stride(0), lse.data_ptr(), num_seqs,
                                   q_len, kv_len, num_heads, scale, ws.data_ptr(), ws.numel() * 4)
    _lib.check(rc, "pi3b_attn_fwd_hd64_ws")


def attention_merge(parts, lses, out: torch.Tensor, num_heads: int = 16, lse_out: Optional[torch.Tensor] = None):
    """out = exact softmax-weighted combination of 1..4 partial attention results (bf16 [rows, 64*heads], equal row
    pitch) with their log-sum-exps (fp32 [heads, rows]). `out` may alias parts[0]."""
    import ctypes
    lib = _lib.load()
    n = len(parts)
    assert 1 <= n <= 4 and len(lses) == n
    rows = out.shape[0]
    ld = parts[0].stride(0)
    for t, l in zip(parts, lses):
        _req(t, torch.bfloat16, "part"); _req(l, torch.float32, "lse")
        assert t.shape == out.shape and t.stride(1) == 1 and t.stride(0) == ld
        assert l.is_contiguous() and l.shape == (num_heads, rows)
    _req(out, torch.bfloat16, "out")
    pa = (ctypes.c_void_p * n)(*[t.data_ptr() for t in parts])
    la = (ctypes.c_void_p * n)(*[l.data_ptr() for l in lses])
    rc = lib.pi3b_attn_merge(_stream(), n, ctypes.cast(pa, ctypes.c_void_p), ctypes.cast(la, ctypes.c_void_p), ld,
                             out.data_ptr(), out.stride(0), _ptr(lse_out), rows, num_heads)
    _lib.check(rc, "pi3b_attn_merge")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out: torch.Tensor,
              override_rows: Optional[torch.Tensor] = None, period: int = 0) -> torch.Tensor:
    """out = LayerNorm_1024(x) (fp32 in, bf16 or fp32 out). Rows r with r % period < len(override_rows) are
    replaced by override_rows[r % period] instead (token re-labelling between encoder and decoder)."""
    lib = _lib.load()
    _req(x, torch.float32, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    assert x.dim() == 2 and x.stride(1) == 1 and out.shape == x.shape and out.stride(1) == 1 and out.is_cuda
    n_ov = 0
    if override_rows is not None:
        _req(override_rows, torch.float32, "override_rows")
        assert override_rows.is_contiguous() and override_rows.shape[1] == x.shape[1]
        n_ov = override_rows.shape[0]
    rc = lib.pi3b_layernorm(_stream(), x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), eps,
                            out.data_ptr(), out.stride(0), 1 if out.dtype == torch.float32 else 0, x.shape[0],
                            x.shape[1], _ptr(override_rows), period, n_ov)
    _lib.check(rc, "pi3b_layernorm")
    return out


def qknorm_rope(qkv: torch.Tensor, num_heads: int, tokens_per_frame: int, n_special: int, patch_w: int,
                cs_table: Optional[torch.Tensor], qn: Optional[tuple] = None, kn: Optional[tuple] = None,
                eps: float = 1e-5, q_scale: float = 1.0) -> torch.Tensor:
    """In place on the q and k thirds of packed qkv: optional LayerNorm(64) then RoPE-2D (if cs_table given).
    cs_table: fp32 [max_pos+1, 16, 2] = (cos, sin) of pos * base^(-j/16). q_scale as in `gemm_qkv`."""
    lib = _lib.load()
    _req(qkv, torch.bfloat16, "qkv")
    assert qkv.dim() == 2 and qkv.stride(1) == 1
    ptrs = [None, None, None, None]
    if qn is not None:
        for t in (*qn, *kn):
            _req(t, torch.float32, "qk-norm parameter")
        ptrs = [qn[0].data_ptr(), qn[1].data_ptr(), kn[0].data_ptr(), kn[1].data_ptr()]
    rows = 0
    if cs_table is not None:
        _req(cs_table, torch.float32, "cs_table")
        assert cs_table.is_contiguous() and cs_table.shape[1:] == (16, 2)
        rows = cs_table.shape[0]
    rc = lib.pi3b_qknorm_rope(_stream(), qkv.data_ptr(), qkv.stride(0), qkv.shape[0], num_heads, tokens_per_frame,
                              n_special, patch_w, *ptrs, eps, _ptr(cs_table), rows, 1 if cs_table is not None else 0,
                              q_scale)
    _lib.check(rc, "pi3b_qknorm_rope")
    return qkv


def patch_pack(imgs: torch.Tensor, out: torch.Tensor, mean, std) -> torch.Tensor:
    """imgs fp32 [N,3,H,W] -> out bf16 [N*P, ldk] (normalised im2col of the 14x14/s14 patch conv)."""
    import ctypes
    lib = _lib.load()
    _req(imgs, torch.float32, "imgs"); _req(out, torch.bfloat16, "out")
    assert imgs.is_contiguous() and out.is_contiguous()
    N, _, H, W = imgs.shape
    assert out.shape[0] == N * (H // 14) * (W // 14)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    rc = lib.pi3b_patch_pack(_stream(), imgs.data_ptr(), out.data_ptr(), N, H, W, out.shape[1],
                             ctypes.cast(m3, ctypes.c_void_p), ctypes.cast(s3, ctypes.c_void_p))
    _lib.check(rc, "pi3b_patch_pack")
    return out


def broadcast_rows(table: torch.Tensor, out: torch.Tensor, num_groups: int) -> torch.Tensor:
    lib = _lib.load()
    _req(table, torch.float32, "table"); _req(out, torch.float32, "out")
    assert table.is_contiguous() and out.is_contiguous() and out.shape[0] == num_groups * table.shape[0]
    rc = lib.pi3b_broadcast_rows(_stream(), table.data_ptr(), out.data_ptr(), num_groups, table.shape[0], table.shape[1])
    _lib.check(rc, "pi3b_broadcast_rows")
    return out


def cast_bf16(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _req(x, torch.float32, "x"); _req(out, torch.bfloat16, "out")
    assert x.shape == out.shape and x.stride(1) == 1 and out.stride(1) == 1
    rc = lib.pi3b_cast_f32_bf16(_stream(), x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0], x.shape[1])
    _lib.check(rc, "pi3b_cast_f32_bf16")
    return out


def unpatchify_points_conf(feat_pts, feat_conf, poses, local_points, points, conf, N, H, W, tokens_per_frame,
                           n_special) -> None:
    lib = _lib.load()
    for t, nm in ((feat_pts, "feat_pts"), (feat_conf, "feat_conf"), (poses, "poses"), (local_points, "local_points"),
                  (points, "points"), (conf, "conf")):
        _req(t, torch.float32, nm)
    assert poses.is_contiguous() and local_points.is_contiguous() and points.is_contiguous() and conf.is_contiguous()
    rc = lib.pi3b_unpatchify_points_conf(_stream(), feat_pts.data_ptr(), feat_pts.stride(0), feat_conf.data_ptr(),
                                         feat_conf.stride(0), poses.data_ptr(), local_points.data_ptr(),
                                         points.data_ptr(), conf.data_ptr(), N, H, W, tokens_per_frame, n_special)
    _lib.check(rc, "pi3b_unpatchify_points_conf")


def meanpool_tokens(x: torch.Tensor, out: torch.Tensor, N: int, tokens_per_frame: int, n_special: int) -> torch.Tensor:
    lib = _lib.load()
    _req(x, torch.float32, "x"); _req(out, torch.float32, "out")
    assert x.stride(1) == 1 and out.is_contiguous() and out.shape == (N, x.shape[1])
    rc = lib.pi3b_meanpool_tokens(_stream(), x.data_ptr(), x.stride(0), out.data_ptr(), N, x.shape[1], tokens_per_frame,
                                  n_special)
    _lib.check(rc, "pi3b_meanpool_tokens")
    return out


def pose_from_9d(rot9: torch.Tensor, trans: torch.Tensor, poses: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _req(rot9, torch.float32, "rot9"); _req(trans, torch.float32, "trans"); _req(poses, torch.float32, "poses")
    assert poses.is_contiguous() and rot9.stride(1) == 1 and trans.stride(1) == 1
    rc = lib.pi3b_pose_from_9d(_stream(), rot9.data_ptr(), rot9.stride(0), trans.data_ptr(), trans.stride(0),
                               poses.data_ptr(), rot9.shape[0])
    _lib.check(rc, "pi3b_pose_from_9d")
    return poses


def mask_depth_edge(local_points: torch.Tensor, conf: torch.Tensor, mask: torch.Tensor, conf_thr: float = 0.1,
                    rtol: float = 0.03) -> torch.Tensor:
    """mask (N,H,W) uint8 = sigmoid(conf) > conf_thr and not depth_edge(z, rtol); z = local_points[..., 2]."""
    lib = _lib.load()
    _req(local_points, torch.float32, "local_points"); _req(conf, torch.float32, "conf")
    assert local_points.is_contiguous() and conf.is_contiguous() and mask.is_contiguous() and mask.dtype == torch.uint8
    N, H, W = local_points.shape[:3]
    rc = lib.pi3b_mask_depth_edge(_stream(), local_points.data_ptr() + 8, 3, conf.data_ptr(), mask.data_ptr(), N, H, W,
                                  conf_thr, rtol)
    _lib.check(rc, "pi3b_mask_depth_edge")
    return mask


def keypoint_gather(points, local_points, conf, mask, images, keypoints, grid_on_cuda: bool = True):
    """Sample dense maps at keypoints (N,K,2) (x, y): bilinear points / local_points / colours (uint8),
    nearest conf / mask. Returns (points (N,K,3) f32, local (N,K,3) f32, conf (N,K,1) f32, mask (N,K,1) bool,
    colors (N,K,3) uint8 or None). Bit-exact with the reference's torch ops; `grid_on_cuda` says on which device the
    reference would have computed the normalised grid (where its keypoints live): CUDA (default extractor) or CPU."""
    lib = _lib.load()
    for t, nm in ((points, "points"), (local_points, "local_points"), (conf, "conf"), (keypoints, "keypoints")):
        _req(t, torch.float32, nm)
        assert t.is_contiguous()
    assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
    N, H, W = points.shape[:3]
    K = keypoints.shape[1]
    dev = points.device
    o_p = torch.empty(N, K, 3, device=dev); o_l = torch.empty(N, K, 3, device=dev)
    o_c = torch.empty(N, K, 1, device=dev); o_m = torch.empty(N, K, 1, device=dev, dtype=torch.uint8)
    o_col = None
    if images is not None:
        _req(images, torch.float32, "images")
        assert images.is_contiguous() and images.shape == (N, 3, H, W)
        o_col = torch.empty(N, K, 3, device=dev, dtype=torch.uint8)
    rc = lib.pi3b_keypoint_gather(_stream(), points.data_ptr(), local_points.data_ptr(), conf.data_ptr(),
                                  mask.data_ptr(), _ptr(images), keypoints.data_ptr(), o_p.data_ptr(), o_l.data_ptr(),
                                  o_c.data_ptr(), o_m.data_ptr(), _ptr(o_col), N, K, H, W, 1 if grid_on_cuda else 0)
    _lib.check(rc, "pi3b_keypoint_gather")
    return o_p, o_l, o_c, o_m.bool(), o_col


def sample_grid(local_points, conf, ys, xs, conf_thr: float = 0.1):
    """Nearest-downsampled local points [N, hs*ws, 3] and validity (sigmoid(conf) > thr) [N, hs*ws] uint8."""
    lib = _lib.load()
    _req(local_points, torch.float32, "local_points"); _req(conf, torch.float32, "conf")
    assert ys.dtype == torch.int32 and xs.dtype == torch.int32 and ys.is_cuda and xs.is_cuda
    N, H, W = local_points.shape[:3]
    hs, ws = ys.numel(), xs.numel()
    pts = torch.empty(N, hs * ws, 3, device=local_points.device)
    val = torch.empty(N, hs * ws, device=local_points.device, dtype=torch.uint8)
    rc = lib.pi3b_sample_grid(_stream(), local_points.data_ptr(), conf.data_ptr(), ys.data_ptr(), xs.data_ptr(),
                              pts.data_ptr(), val.data_ptr(), N, H, W, hs, ws, conf_thr)
    _lib.check(rc, "pi3b_sample_grid")
    return pts, val


def focal_shift_lm(pts: torch.Tensor, valid: torch.Tensor, uv: torch.Tensor, ftol: float = 1e-3) -> torch.Tensor:
    """Per-frame focal / z-shift solve on the device (scipy's 1-parameter LM restated). pts fp32 [N,S,3], valid uint8
    [N,S], uv fp32 [S,2]. Returns fp32 [N,4] = (shift, focal, nfev, #valid)."""
    lib = _lib.load()
    _req(pts, torch.float32, "pts"); _req(uv, torch.float32, "uv")
    assert pts.is_contiguous() and valid.is_contiguous() and uv.is_contiguous() and valid.dtype == torch.uint8
    N, S = pts.shape[:2]
    assert valid.shape == (N, S) and uv.shape == (S, 2) and valid.is_cuda and uv.is_cuda
    out = torch.empty(N, 4, device=pts.device)
    rc = lib.pi3b_focal_shift_lm(_stream(), pts.data_ptr(), valid.data_ptr(), uv.data_ptr(), N, S, ftol, out.data_ptr())
    _lib.check(rc, "pi3b_focal_shift_lm")
    return out


def median_ratio(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Lower median of a/b over mask (1-D strided views allowed). Returns a 2-vector (median, count) on device."""
    lib = _lib.load()
    _req(a, torch.float32, "a"); _req(b, torch.float32, "b")
    assert a.dim() == 1 and b.dim() == 1 and mask.dim() == 1 and mask.dtype == torch.uint8 and mask.is_contiguous()
    out = torch.empty(2, device=a.device)
    rc = lib.pi3b_median_ratio(_stream(), a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), mask.data_ptr(),
                               a.numel(), out.data_ptr())
    _lib.check(rc, "pi3b_median_ratio")
    return out


def umeyama_sim3(qry: torch.Tensor, ref: torch.Tensor, valid: Optional[torch.Tensor] = None):
    """Batched closed-form similarity mapping qry -> ref. qry, ref: (B, n, 3) fp32; valid (B, n) uint8 or None.
    Returns (T (B,4,4) = [sR|t], scale (B,), count (B,))."""
    lib = _lib.load()
    _req(qry, torch.float32, "qry"); _req(ref, torch.float32, "ref")
    assert qry.shape == ref.shape and qry.dim() == 3 and qry.is_contiguous() and ref.is_contiguous()
    if valid is not None:
        assert valid.dtype == torch.uint8 and valid.is_contiguous() and valid.shape == qry.shape[:2]
    out = torch.empty(qry.shape[0], 18, device=qry.device)
    rc = lib.pi3b_umeyama_sim3_batched(_stream(), qry.data_ptr(), ref.data_ptr(), _ptr(valid), qry.shape[0],
                                       qry.shape[1], out.data_ptr())
    _lib.check(rc, "pi3b_umeyama_sim3_batched")
    return out[:, :16].reshape(-1, 4, 4), out[:, 16], out[:, 17]


# ---------------------------------------------------------------------------------------------------------------
# MoGe metric-depth pass (row a15): dense-prediction pieces. Feature maps are channels-last fp32 [H*W, C].
# ---------------------------------------------------------------------------------------------------------------
def groupnorm_stats(x: torch.Tensor, groups: int, eps: float = 1e-5) -> torch.Tensor:
    """(mean, rstd) per group of one sample's map x fp32 [pixels, C] -> fp32 [groups, 2] on the device."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    ws = torch.empty(2 * groups, device=x.device, dtype=torch.float64)
    out = torch.empty(groups, 2, device=x.device, dtype=torch.float32)
    rc = lib.pi3b_groupnorm_stats(_stream(), x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], groups, eps,
                                  ws.data_ptr(), out.data_ptr())
    _lib.check(rc, "pi3b_groupnorm_stats")
    return out


def im2col3x3(x: torch.Tensor, H: int, W: int, C: int, out: torch.Tensor, blocked: bool = False,
              mean_rstd: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
              beta: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """bf16 [H*W, 9C] operand of a 3x3 replicate-padded conv, with GroupNorm + ReLU fused into the gather.
    blocked: x is the [H/2*W/2, 4C] output of a 2x2/stride-2 transposed-conv GEMM."""
    lib = _lib.load()
    _req(x, torch.float32, "x"); _req(out, torch.bfloat16, "out")
    assert x.dim() == 2 and x.stride(1) == 1 and out.stride(1) == 1 and out.shape == (H * W, 9 * C)
    assert x.shape == ((H // 2) * (W // 2), 4 * C) if blocked else x.shape == (H * W, C)
    groups = 0
    if mean_rstd is not None:
        _req(mean_rstd, torch.float32, "mean_rstd"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
        groups = mean_rstd.shape[0]
    rc = lib.pi3b_im2col3x3(_stream(), x.data_ptr(), x.stride(0), H, W, C, 1 if blocked else 0, _ptr(mean_rstd), groups,
                            _ptr(gamma), _ptr(beta), 1 if relu else 0, out.data_ptr(), out.stride(0))
    _lib.check(rc, "pi3b_im2col3x3")
    return out


def uv_affine(x: torch.Tensor, H: int, W: int, w: torch.Tensor, bias: Optional[torch.Tensor], u_tab: torch.Tensor,
              v_tab: torch.Tensor, accumulate: bool) -> torch.Tensor:
    """x[p, c] (+)= w[c,0]*u[p % W] + w[c,1]*v[p // W] + bias[c]  (1x1 conv on the view-plane uv map)."""
    lib = _lib.load()
    for t, nm in ((x, "x"), (w, "w"), (u_tab, "u_tab"), (v_tab, "v_tab")):
        _req(t, torch.float32, nm)
    assert x.shape[0] == H * W and w.shape == (x.shape[1], 2)
    assert u_tab.numel() == W and v_tab.numel() == H and w.stride(1) == 1
    rc = lib.pi3b_uv_affine(_stream(), x.data_ptr(), x.stride(0), H, W, x.shape[1], w.data_ptr(), w.stride(0),
                            _ptr(bias), u_tab.data_ptr(), v_tab.data_ptr(), 1 if accumulate else 0)
    _lib.check(rc, "pi3b_uv_affine")
    return x


def resample2d(src: torch.Tensor, src_layout: str, dst: torch.Tensor, dst_layout: str, tab_y, tab_x) -> torch.Tensor:
    """Separable resampling with tap tables (y0 int32 [Ho], wy fp32 [Ho, ty]) / (x0, wx). Layout 'chw' or 'hwc'."""
    lib = _lib.load()
    _req(src, torch.float32, "src"); _req(dst, torch.float32, "dst")
    assert src.is_contiguous() and dst.is_contiguous() and src.dim() == 3 and dst.dim() == 3

    def geom(t, layout):
        if layout == "chw":
            C, H, W = t.shape
            return C, H, W, H * W, W, 1
        H, W, C = t.shape
        return C, H, W, 1, W * C, C
    C, Hi, Wi, isc, isy, isx = geom(src, src_layout)
    C2, Ho, Wo, osc, osy, osx = geom(dst, dst_layout)
    (y0, wy), (x0, wx) = tab_y, tab_x
    assert C == C2 and y0.numel() == Ho and x0.numel() == Wo and wy.shape[0] == Ho and wx.shape[0] == Wo
    assert y0.dtype == torch.int32 and x0.dtype == torch.int32 and y0.is_cuda and wy.is_cuda and x0.is_cuda and wx.is_cuda
    rc = lib.pi3b_resample2d(_stream(), src.data_ptr(), isc, isy, isx, Hi, Wi, dst.data_ptr(), osc, osy, osx, C, Ho, Wo,
                             y0.data_ptr(), wy.data_ptr(), wy.shape[1], x0.data_ptr(), wx.data_ptr(), wx.shape[1])
    _lib.check(rc, "pi3b_resample2d")
    return dst


MOGE_REMAP = {"linear": 0, "sinh": 1, "exp": 2, "sinh_exp": 3}


def moge_remap(points: torch.Tensor, mask_logit: Optional[torch.Tensor], mask_prob: Optional[torch.Tensor], mode: str):
    lib = _lib.load()
    _req(points, torch.float32, "points")
    assert points.is_contiguous() and points.shape[-1] == 3
    n = points.numel() // 3
    if mask_logit is not None:
        _req(mask_logit, torch.float32, "mask_logit"); _req(mask_prob, torch.float32, "mask_prob")
        assert mask_logit.numel() == n and mask_prob.numel() == n and mask_logit.is_contiguous() and mask_prob.is_contiguous()
    rc = lib.pi3b_moge_remap(_stream(), points.data_ptr(), _ptr(mask_logit), _ptr(mask_prob), MOGE_REMAP[mode], n)
    _lib.check(rc, "pi3b_moge_remap")
    return points


def moge_depth(points: torch.Tensor, mask_prob: Optional[torch.Tensor], shift: float, log_scale: Optional[torch.Tensor],
               apply_mask: bool):
    """depth (fp32) and mask (uint8) of MoGeModel.infer from remapped points [H,W,3]."""
    lib = _lib.load()
    _req(points, torch.float32, "points")
    assert points.is_contiguous() and points.shape[-1] == 3
    n = points.numel() // 3
    depth = torch.empty(points.shape[:-1], device=points.device, dtype=torch.float32)
    mask = torch.empty(points.shape[:-1], device=points.device, dtype=torch.uint8)
    if mask_prob is not None:
        _req(mask_prob, torch.float32, "mask_prob")
        assert mask_prob.is_contiguous() and mask_prob.numel() == n
    if log_scale is not None:
        _req(log_scale, torch.float32, "log_scale")
    rc = lib.pi3b_moge_depth(_stream(), points.data_ptr(), _ptr(mask_prob), float(shift), _ptr(log_scale),
                             1 if apply_mask else 0, depth.data_ptr(), mask.data_ptr(), n)
    _lib.check(rc, "pi3b_moge_depth")
    return depth, mask


def resize_u8_pil(frames_u8: torch.Tensor, out: torch.Tensor, tab_y, tab_x) -> torch.Tensor:
    """uint8 [N,Hi,Wi,3] -> fp32 [N,3,Ho,Wo] in [0,1], bit-identical to PIL BILINEAR resize + ToTensor.
    tab_* = (min int32 [out], count int32 [out], coeffs int32 [out, ksize]) on the device
    (host.resample.pil_bilinear_coeffs)."""
    lib = _lib.load()
    _req(frames_u8, torch.uint8, "frames_u8"); _req(out, torch.float32, "out")
    assert frames_u8.is_contiguous() and out.is_contiguous() and frames_u8.dim() == 4 and frames_u8.shape[3] == 3
    N, Hi, Wi, _ = frames_u8.shape
    assert out.shape[0] == N and out.shape[1] == 3
    Ho, Wo = out.shape[2:]
    (ymin, ycnt, ky), (xmin, xcnt, kx) = tab_y, tab_x
    for t in (ymin, ycnt, ky, xmin, xcnt, kx):
        assert t.dtype == torch.int32 and t.is_cuda and t.is_contiguous()
    assert ymin.numel() == Ho and xmin.numel() == Wo and ky.shape[0] == Ho and kx.shape[0] == Wo
    rc = lib.pi3b_resize_u8_pil(_stream(), frames_u8.data_ptr(), N, Hi, Wi, out.data_ptr(), Ho, Wo, ymin.data_ptr(),
                                ycnt.data_ptr(), ky.data_ptr(), ky.shape[1], xmin.data_ptr(), xcnt.data_ptr(),
                                kx.data_ptr(), kx.shape[1])
    _lib.check(rc, "pi3b_resize_u8_pil")
    return out
