"""Host-side launchers: torch tensors in, raw pointers + shapes through the C ABI, nothing else.

Each function mirrors one entry point of include/pi3b200.h. Tensors must already live on the CUDA
device; the kernels are enqueued on torch's current stream and never synchronise.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.Pi3bError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.Pi3bError(f"{name} must be {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, bias: Optional[torch.Tensor] = None,
         gamma: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         k: Optional[int] = None) -> torch.Tensor:
    """out[M,N] = [resid +] gamma * act(a[M,K] @ w[N,K]^T + bias).

    a, w: bf16 (tcgen05 kind::f16) or fp32 (read as TF32); row-major with unit inner stride, row pitch
    a multiple of 16 bytes. out: bf16 or fp32. bias/gamma: fp32 [N]. resid: fp32 (may alias out).
    `k` overrides the contraction length (for K-padded operands).
    """
    lib = _lib.load()
    assert a.dim() == 2 and w.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    assert a.dtype == w.dtype and a.dtype in (torch.bfloat16, torch.float32)
    _req(a, a.dtype, "a"); _req(w, a.dtype, "w")
    M = a.shape[0]
    N = w.shape[0]
    K = k if k is not None else a.shape[1]
    assert out.shape[0] == M and out.shape[1] == N
    assert out.dtype in (torch.bfloat16, torch.float32) and out.is_cuda
    for t, nm in ((bias, "bias"), (gamma, "gamma"), (resid, "resid")):
        if t is not None:
            _req(t, torch.float32, nm)
    rc = lib.pi3b_gemm(_stream(), 0 if a.dtype == torch.bfloat16 else 1, a.data_ptr(), a.stride(0), w.data_ptr(),
                       w.stride(0), M, N, K, _ptr(bias), _ptr(gamma), _ptr(resid),
                       resid.stride(0) if resid is not None else 0, out.data_ptr(), out.stride(0),
                       1 if out.dtype == torch.float32 else 0, act)
    _lib.check(rc, "pi3b_gemm")
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
    rc = lib.pi3b_attn_fwd_hd64(_stream(), qkv.data_ptr(), qkv.data_ptr(), qkv.data_ptr(), ld, ld, ld, 0, C, 2 * C,
                                out.data_ptr(), out.stride(0), num_seqs, seq_len, num_heads, scale)
    _lib.check(rc, "pi3b_attn_fwd_hd64")
    return out
