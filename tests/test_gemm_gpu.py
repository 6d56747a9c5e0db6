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
    # CTA-pair kernels, one case per epilogue instantiation and its fall-backs
    (19999, 1280, 512, "bf16", True, True, True, False, 0),    # streamlined fp32 epilogue without a residual, 5 column tiles
    (19999, 1024, 1024, "bf16", True, True, False, True, 2),   # streamlined: ReLU + residual
    (19999, 1152, 512, "bf16", True, True, True, True, 0),     # N % 256 != 0: general epilogue with a residual
    (19999, 1024, 512, "bf16", False, True, True, True, 0),    # bf16 output with a residual: general epilogue
    (19999, 1024, 512, "bf16", True, True, False, False, 1),   # GELU into fp32: general epilogue
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


def test_residual_into_a_separate_output_and_row_regrouping():
    """The streamlined residual epilogue with out != resid, and the row-regrouped residual GEMM (patch embedding:
    GEMM row m lands on row (m // grp_in) * grp_out + off + m % grp_in) at a size that runs the CTA-pair kernel."""
    from pi3_slam_b200 import ops

    torch.manual_seed(3)
    dev = "cuda"
    M, N, K = 20001, 1024, 512
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias, gamma = torch.randn(N, device=dev), torch.rand(N, device=dev) + 0.5
    resid = torch.randn(M, N, device=dev)
    keep = resid.clone()
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(a, w, out, bias=bias, gamma=gamma, resid=resid)
    ref = _ref(a, w, bias, gamma, keep, 0)
    assert torch.equal(resid, keep)
    assert ((out - ref).norm() / ref.norm()).item() < 1e-5

    P, T, off = 999, 1004, 5
    frames = 21
    a = (torch.randn(frames * P, K, device=dev) * 0.5).bfloat16()
    x = torch.randn(frames * T, N, device=dev)
    keep = x.clone()
    ops.gemm(a, w, x, bias=bias, resid=x, row_remap=(P, T, off))
    y = (a.float() @ w.float().t() + bias).view(frames, P, N)
    want = keep.view(frames, T, N).clone()
    want[:, off:] += y
    assert torch.equal(x.view(frames, T, N)[:, :off], keep.view(frames, T, N)[:, :off])   # special-token rows untouched
    assert ((x.view(frames, T, N) - want).norm() / want.norm()).item() < 1e-5


# M: narrow-tile path (< 148 wide tiles), wide single-CTA path (PI3B_GEMM_CLUSTER=0 in the subprocess case below) and the
# default wide 2-CTA-cluster path; ragged M tails; (norm, rope) combinations used by the model: decoder blocks (both),
# head decoders (rope only).
@pytest.mark.parametrize("M,T,pw", [(643 * 3, 643, 29), (1004 * 20, 1004, 37), (261 * 5 + 7, 261, 16)])
@pytest.mark.parametrize("norm,rope,q_scale", [(True, True, 1.0), (False, True, 1.0), (True, False, 1.0),
                                               (True, True, 0.125 * 1.4426950408889634), (False, True, 0.18033688)])
def test_fused_qkv_epilogue_is_bit_identical_to_gemm_plus_qknorm_rope(M, T, pw, norm, rope, q_scale):
    """pi3b_gemm_qkv == pi3b_gemm followed by pi3b_qknorm_rope, bit for bit (shared arithmetic, csrc/qk_math.cuh), with and
    without the softmax scale folded into the queries."""
    from pi3_slam_b200 import ops

    H, K = 16, 1024
    torch.manual_seed(M + T)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(3 * 64 * H, K, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(3 * 64 * H, device="cuda") * 0.1
    qn = (torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1)
    kn = (torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1)
    qkn = torch.cat([qn[0], qn[1], kn[0], kn[1]]).contiguous()
    max_pos = max(pw, (T - 5 + pw - 1) // pw)
    inv_freq = 1.0 / (100.0 ** (torch.arange(0, 32, 2).float() / 32))
    ang = torch.einsum("i,j->ij", torch.arange(max_pos + 1, dtype=torch.float32), inv_freq)
    cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().cuda()

    want = torch.empty(M, 3 * 64 * H, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, want, bias=bias)
    if q_scale != 1.0:
        unscaled = want.clone()
        ops.qknorm_rope(unscaled, H, T, 5, pw, cs if rope else None, qn if norm else None, kn if norm else None, 1e-5)
    ops.qknorm_rope(want, H, T, 5, pw, cs if rope else None, qn if norm else None, kn if norm else None, 1e-5, q_scale)
    got = torch.full_like(want, float("nan"))
    ops.gemm_qkv(a, w, got, bias, H, qkn if norm else None, 1e-5, cs if rope else None, T, 5, pw, q_scale)
    torch.cuda.synchronize()
    if q_scale != 1.0:
        # only the queries are scaled (k and v untouched), by q_scale up to the one bf16 rounding
        C = 64 * H
        assert torch.equal(got[:, C:], unscaled[:, C:])
        rel = (got[:, :C].float() - unscaled[:, :C].float() * q_scale).abs() / (unscaled[:, :C].float().abs() * q_scale + 1e-20)
        assert rel.max().item() <= 2.0 ** -7      # two bf16 roundings (x, then x * q_scale), half an ulp = 2^-8 relative each
    assert torch.isfinite(got.float()).all()
    neq = (got.view(torch.int16) != want.view(torch.int16)).sum().item()
    assert neq == 0, f"{neq} of {got.numel()} elements differ, max |d| {(got.float() - want.float()).abs().max().item():.3e}"


def _guarded(rows, cols, dtype, pad_rows=64):
    """A [rows, cols] view in the middle of a larger buffer filled with a sentinel: (view, check) — check() asserts that
    nothing outside the view was written (the epilogues store through transposed staging buffers, 8 rows per instruction:
    a row-bound slip on a ragged last tile would land here)."""
    sentinel = 12345.0
    big = torch.full((rows + 2 * pad_rows, cols), sentinel, device="cuda", dtype=dtype)
    view = big[pad_rows:pad_rows + rows]

    def check():
        assert bool((big[:pad_rows] == sentinel).all()) and bool((big[pad_rows + rows:] == sentinel).all())
    return view, check


@pytest.mark.parametrize("M", [19999, 20001, 2088 + 37])
def test_epilogue_stores_stay_inside_a_ragged_last_tile(M):
    """Every epilogue instantiation on an M that is not a multiple of 128 (nor of 8), output surrounded by canaries."""
    from pi3_slam_b200 import ops

    torch.manual_seed(M)
    K, H = 512, 16
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    bias1 = torch.randn(1024, device="cuda")
    w1 = (torch.randn(1024, K, device="cuda") * 0.05).bfloat16()
    # general epilogue, bf16 out (+ GELU)
    out, check = _guarded(M, 1024, torch.bfloat16)
    ops.gemm(a, w1, out, bias=bias1, act=ops.ACT_GELU)
    torch.cuda.synchronize(); check()
    ref = torch.nn.functional.gelu(a.float() @ w1.float().t() + bias1)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 5e-3
    # fp32 residual read-modify-write in place
    x, check = _guarded(M, 1024, torch.float32)
    x.copy_(torch.randn(M, 1024, device="cuda"))
    keep = x.clone()
    ops.gemm(a, w1, x, bias=bias1, gamma=torch.full((1024,), 0.5, device="cuda"), resid=x)
    torch.cuda.synchronize(); check()
    ref = keep + 0.5 * (a.float() @ w1.float().t() + bias1)
    assert ((x - ref).norm() / ref.norm()).item() < 1e-5
    # fused QKV epilogue
    T, pw = 261, 16
    wq = (torch.randn(3 * 64 * H, K, device="cuda") * 0.05).bfloat16()
    bq = torch.randn(3 * 64 * H, device="cuda") * 0.1
    qkn = torch.cat([torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1,
                     torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1]).contiguous()
    inv_freq = 1.0 / (100.0 ** (torch.arange(0, 32, 2).float() / 32))
    ang = torch.einsum("i,j->ij", torch.arange(pw + 2, dtype=torch.float32), inv_freq)
    cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().cuda()
    qkv, check = _guarded(M, 3 * 64 * H, torch.bfloat16)
    ops.gemm_qkv(a, wq, qkv, bq, H, qkn, 1e-5, cs, T, 5, pw, ops.Q_PRESCALE)
    torch.cuda.synchronize(); check()
    assert torch.isfinite(qkv.float()).all()
