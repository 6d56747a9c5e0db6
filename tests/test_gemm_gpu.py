"""GPU parity: tcgen05 GEMM + fused epilogues (through the C ABI) vs a plain torch fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, gamma, resid, act):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    if gamma is not None:
        y = y * gamma
    if resid is not None:
        y = y + resid
    return y


CASES = [
    # M, N, K, in_dtype, out_f32, bias, gamma, resid, act
    (128, 256, 64, "bf16", True, False, False, False, 0),
    (256, 512, 1024, "bf16", False, True, False, False, 0),
    (2088, 3072, 1024, "bf16", False, True, False, False, 0),
    (2088, 4096, 1024, "bf16", False, True, False, False, 1),
    (2088, 1024, 4096, "bf16", True, True, True, True, 0),
    (20000, 1024, 1024, "bf16", True, True, True, True, 0),   # wide-tile path (>= 148 tiles)
    (20000, 4096, 1024, "bf16", False, True, False, False, 1),
    (1999, 1024, 592, "bf16", False, True, False, False, 0),   # patch embed: K tail (592 = 588 padded)
    (1999, 588, 1024, "fp32", True, True, False, False, 0),    # tf32 point head, ragged N
    (1999, 196, 1024, "fp32", True, True, False, False, 0),
    (1999, 512, 512, "fp32", True, True, False, True, 2),      # camera res-conv: relu + residual
    (37, 9, 512, "fp32", True, True, False, False, 0),
    (30000, 588, 1024, "fp32", True, True, False, False, 0),
]


@pytest.mark.parametrize("M,N,K,dt,out_f32,has_bias,has_gamma,has_resid,act", CASES)
def test_gemm_matches_torch(M, N, K, dt, out_f32, has_bias, has_gamma, has_resid, act):
    from pi3_slam_b200 import ops

    torch.manual_seed(M * 7 + N * 3 + K)
    dev = "cuda"
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    a = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
    w = (torch.randn(N, K, device=dev) * 0.05).to(dtype)
    bias = torch.randn(N, device=dev) if has_bias else None
    gamma = torch.rand(N, device=dev) + 0.5 if has_gamma else None
    resid = torch.randn(M, N, device=dev) if has_resid else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    if has_resid and out_f32:
        out.copy_(resid)           # in-place residual accumulate, as the engine uses it
        ops.gemm(a, w, out, bias=bias, gamma=gamma, resid=out, act=act)
    else:
        ops.gemm(a, w, out, bias=bias, gamma=gamma, resid=resid, act=act)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, gamma, resid, act)
    err = (out.float() - ref).norm() / ref.norm()
    # bf16 operands are exact in the reference too; the only error is accumulation order (+ bf16
    # output rounding, or tf32 operand truncation for the fp32 path).
    tol = 2e-3 if (not out_f32 or dt == "fp32") else 1e-5
    assert torch.isfinite(out.float()).all()
    assert err.item() < tol, f"rel err {err.item():.3e}"
    maxabs = (out.float() - ref).abs().max().item()
    assert maxabs < 0.05 * max(1.0, ref.abs().max().item())
