"""ctypes binding of libpi3b200.so — the C-ABI boundary declared in include/pi3b200.h.

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
PyTorch is only used by callers for device memory and streams; nothing torch-typed crosses this file.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p
from pathlib import Path

import os as _os

# PI3B_LIB lets kernel experiments point at an alternative build of the same ABI (never a different backend)
_LIB_PATH = Path(_os.environ.get("PI3B_LIB", Path(__file__).resolve().parent / "libpi3b200.so"))
_lib = None

# name -> (restype, argtypes); mirrors include/pi3b200.h one to one.
_P = c_void_p
SIGNATURES = {
    "pi3b_version": (c_int, []),
    "pi3b_last_error": (c_char_p, []),
    # stream, in_dtype, A, lda, W, ldw, M, N, K, bias, gamma, resid, ld_resid, out, ldo, out_f32, act, grp_in, grp_out, grp_off
    "pi3b_gemm": (c_int, [_P, c_int, _P, c_int64, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, c_int64, _P,
                          c_int64, c_int, c_int, c_int, c_int, c_int]),
    # stream, A, lda, W, ldw, M, num_heads, K, bias, out, ldo, qk_norm, eps, cs_table, table_rows, tokens_per_frame, n_special,
    # patch_w, q_scale
    "pi3b_gemm_qkv": (c_int, [_P, _P, c_int64, _P, c_int64, c_int, c_int, c_int, _P, _P, c_int64, _P, c_float, _P, c_int,
                              c_int, c_int, c_int, c_float]),
    # stream, q, k, v, ldq, ldk, ldv, q_col0, k_col0, v_col0, out, ldo, num_seqs, q_len, kv_len, num_heads, scale
    "pi3b_attn_fwd_hd64": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P, c_int64,
                                   c_int, c_int, c_int, c_int, c_float]),
    # stream, q, k, v, ldq, ldk, ldv, q_col0, k_col0, v_col0, out, ldo, lse, num_seqs, q_len, kv_len, num_heads, scale
    "pi3b_attn_fwd_hd64_lse": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P, c_int64, _P,
                                       c_int, c_int, c_int, c_int, c_float]),
    # ... as pi3b_attn_fwd_hd64_lse (lse nullable) ..., workspace, workspace_bytes
    "pi3b_attn_fwd_hd64_ws": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P, c_int64, _P,
                                      c_int, c_int, c_int, c_int, c_float, _P, c_int64]),
    # stream, n_parts, parts(host void*[n]), lses(host float*[n]), ld_part, out, ldo, lse_out, rows, num_heads
    "pi3b_attn_merge": (c_int, [_P, c_int, _P, _P, c_int64, _P, c_int64, _P, c_int, c_int]),
    # stream, x, ldx, gamma, beta, eps, out, ldo, out_f32, M, D, override_rows, period, n_override
    "pi3b_layernorm": (c_int, [_P, _P, c_int64, _P, _P, c_float, _P, c_int64, c_int, c_int, c_int, _P, c_int, c_int]),
    # stream, qkv, ld, M, heads, tokens_per_frame, n_special, patch_w, qn_w, qn_b, kn_w, kn_b, eps, cs_table, table_rows, do_rope,
    # q_scale
    "pi3b_qknorm_rope": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_float, _P,
                                 c_int, c_int, c_float]),
    # stream, imgs, out, N, H, W, ldk, mean3(host), std3(host)
    "pi3b_patch_pack": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    # stream, table, out, num_groups, rows_per_group, D
    "pi3b_broadcast_rows": (c_int, [_P, _P, _P, c_int, c_int, c_int]),
    # stream, feat_pts, ld_pts, feat_conf, ld_conf, poses, local_points, points, conf, N, H, W, tokens_per_frame, n_special
    "pi3b_unpatchify_points_conf": (c_int, [_P, _P, c_int64, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int,
                                            c_int]),
    # stream, x, ldx, out, ldo, M, C
    "pi3b_cast_f32_bf16": (c_int, [_P, _P, c_int64, _P, c_int64, c_int, c_int]),
    # stream, x, ldx, out, N, C, tokens_per_frame, n_special
    "pi3b_meanpool_tokens": (c_int, [_P, _P, c_int64, _P, c_int, c_int, c_int, c_int]),
    # stream, rot9, ld_r, trans, ld_t, poses, N
    "pi3b_pose_from_9d": (c_int, [_P, _P, c_int64, _P, c_int64, _P, c_int]),
    # stream, z, z_stride, conf, mask, N, H, W, conf_thr, rtol
    "pi3b_mask_depth_edge": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_float, c_float]),
    # stream, points, local_points, conf, mask, images, keypoints, out_points, out_local, out_conf, out_mask, out_colors, N, K, H, W, grid_math
    "pi3b_keypoint_gather": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int]),
    # stream, local_points, conf, ys, xs, out_pts, out_valid, N, H, W, hs, ws, conf_thr
    "pi3b_sample_grid": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float]),
    # stream, pts, valid, uv, N, S, ftol, out4
    "pi3b_focal_shift_lm": (c_int, [_P, _P, _P, _P, c_int, c_int, c_float, _P]),
    # stream, points, w2t, kmat, pairs, R, K, width, height, row_uv, row_kp, row_count, row_offset, obs_src, obs_kp, obs_view, obs_xy, capacity
    "pi3b_project_observations": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_double, c_double, _P, _P, _P, _P, _P, _P, _P,
                                          _P, c_int]),
    # stream, a, a_stride, b, b_stride, mask, n, out2
    "pi3b_median_ratio": (c_int, [_P, _P, c_int, _P, c_int, _P, c_int, _P]),
    # stream, qry, ref, valid, num_problems, n_per, out18
    "pi3b_umeyama_sim3_batched": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    # stream, x, ldx, rows, C, groups, eps, workspace(double[2*groups]), mean_rstd(float[2*groups])
    "pi3b_groupnorm_stats": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, c_float, _P, _P]),
    # stream, x, ldx, H, W, C, blocked, mean_rstd, groups, gamma, beta, relu, out, ldo
    "pi3b_im2col3x3": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int, _P, c_int64]),
    # stream, x, ldx, H, W, C, w, ldw, bias, u_tab, v_tab, accumulate
    "pi3b_uv_affine": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, c_int, _P, _P, _P, c_int]),
    # stream, in, in_sc, in_sy, in_sx, Hi, Wi, out, out_sc, out_sy, out_sx, C, Ho, Wo, y0, wy, taps_y, x0, wx, taps_x
    "pi3b_resample2d": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int, c_int, _P, c_int64, c_int64, c_int64, c_int,
                                c_int, c_int, _P, _P, c_int, _P, _P, c_int]),
    # stream, points, mask_logit, mask_prob, mode, n
    "pi3b_moge_remap": (c_int, [_P, _P, _P, _P, c_int, c_int64]),
    # stream, points, mask_prob, shift, log_scale, apply_mask, depth, mask_out, n
    "pi3b_moge_depth": (c_int, [_P, _P, _P, c_float, _P, c_int, _P, _P, c_int64]),
    # stream, in, N, Hi, Wi, out, Ho, Wo, ymin, ycnt, ky, ksize_y, xmin, xcnt, kx, ksize_x
    "pi3b_resize_u8_pil": (c_int, [_P, _P, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P, c_int]),
}


class Pi3bError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """Load libpi3b200.so (optionally compiling it first). Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from .build import build_lib
        build_lib()
    if not _LIB_PATH.exists():
        raise Pi3bError(
            f"{_LIB_PATH} not found: build it with `python -m pi3_slam_b200.build` "
            "(there is no CPU or PyTorch fallback for the pi3 chunk path)")
    try:
        import torch  # noqa: F401  (brings libcudart.so.12 into the process before dlopen)
    except Exception:  # pragma: no cover
        pass
    lib = ctypes.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


launch_count = 0   # number of C-ABI kernel launches issued by this process (bench.py reports it)


def check(rc: int, what: str) -> None:
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = load().pi3b_last_error()
        raise Pi3bError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
