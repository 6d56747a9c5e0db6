"""GPU parity: tcgen05 flash attention (hd=64) through the C ABI vs torch fp32 softmax attention."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, B, S, H):
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)  # (3,B,H,S,64)
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)      # fp32 math
    return o.permute(0, 2, 1, 3).reshape(B * S, H * 64)


# (num_seqs, seq_len, heads): frame-local shapes (T=261, 643, 1004), multiples of the tile sizes, ragged
# tails on both the Q and the KV side, and a chunk-global shape.
# Sequences up to 3000 keys run the four-tile kernel, longer ones the three-tile kernel: both sides of the dispatch are
# covered, and test_each_kernel_when_forced pins each kernel on both kinds of shape.
CASES = [(1, 128, 1), (1, 256, 2), (2, 384, 2), (3, 261, 16), (2, 643, 4), (2, 1004, 16), (1, 5, 1), (1, 129, 2),
         (1, 8 * 261, 16), (1, 3000, 2), (1, 3001, 2), (1, 4099, 3), (1, 9001, 2), (1, 8193, 1)]


@pytest.mark.parametrize("B,S,H", CASES)
def test_attention_matches_torch(B, S, H):
    from pi3_slam_b200 import ops

    torch.manual_seed(B * 1000 + S + H)
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda")
    qkv[:, : 2 * H * 64] *= 1.5        # some spread in the logits
    qkv = qkv.bfloat16()
    out = torch.full((B * S, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention_qkv(qkv, out, B, S, H)
    torch.cuda.synchronize()
    ref = _ref(qkv, B, S, H)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).norm() / ref.norm()
    assert err.item() < 1e-2, f"rel err {err.item():.3e}"   # bf16 P and bf16 output rounding
    assert (out.float() - ref).abs().max().item() < 0.06


def test_attention_large_scores_lazy_rescale():
    """Row max that keeps growing along KV exercises the lazy O-rescale path."""
    from pi3_slam_b200 import ops

    B, S, H = 1, 1024, 2
    torch.manual_seed(5)
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda")
    ramp = torch.linspace(0.2, 6.0, S, device="cuda")[:, None]
    qkv[:, H * 64: 2 * H * 64] *= ramp          # later keys have larger norms -> larger logits
    qkv = qkv.bfloat16()
    out = torch.empty(B * S, H * 64, device="cuda", dtype=torch.bfloat16)
    ops.attention_qkv(qkv, out, B, S, H)
    ref = _ref(qkv, B, S, H)
    err = (out.float() - ref).norm() / ref.norm()
    assert err.item() < 1e-2, f"rel err {err.item():.3e}"


def test_attention_local_queries_against_longer_kv():
    """q_len != kv_len (sequence-parallel layout): a slice of the queries against all keys/values equals the
    matching rows of the full attention."""
    from pi3_slam_b200 import ops

    S, H = 1100, 4
    torch.manual_seed(11)
    qkv = (torch.randn(S, 3 * H * 64, device="cuda") * 1.2).bfloat16()
    ref = _ref(qkv, 1, S, H)
    kv = qkv[:, H * 64:].contiguous()
    for lo, hi in ((0, 300), (300, 829), (829, 1100)):
        out = torch.empty(hi - lo, H * 64, device="cuda", dtype=torch.bfloat16)
        ops.attention_q_kv(qkv[lo:hi], kv, out, 1, hi - lo, S, H)
        err = (out.float() - ref[lo:hi]).norm() / ref[lo:hi].norm()
        assert err.item() < 1e-2, f"rows {lo}:{hi} rel err {err.item():.3e}"


@pytest.mark.parametrize("rev", [3, 4])
def test_each_kernel_when_forced(rev):
    """PI3B_ATTN_TILES pins one kernel for every shape (the library reads it once per process, hence the subprocess):
    each kernel must be correct on short ragged sequences and on a long one, whatever the default dispatch picks."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from pi3_slam_b200 import ops
worst = 0.0
for (B, S, H) in [(2, 643, 4), (1, 129, 2), (1, 8700, 1)]:
    torch.manual_seed(S)
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
    out = torch.full((B * S, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention_qkv(qkv, out, B, S, H)
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * S, H * 64)
    assert torch.isfinite(out.float()).all()
    worst = max(worst, float((out.float() - ref).norm() / ref.norm()))
print("WORST", worst)
""" % root
    env = dict(os.environ, PI3B_ATTN_TILES=str(rev))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = float(r.stdout.strip().split("WORST")[-1])
    assert worst < 1e-2, f"{rev}-tile kernel: rel err {worst:.3e}"


@pytest.mark.parametrize("S,H,splits", [(1100, 4, (0, 300, 829, 1100)), (9000, 2, (0, 4000, 9000)),
                                        (700, 16, (0, 64, 65, 700))])
def test_lse_and_merge_over_key_blocks(S, H, splits):
    """Attention over disjoint key/value blocks + pi3b_attn_merge == attention over all keys (the sequence-parallel
    chunk-global layer overlaps the remote blocks' NVLink transfer with the local block's attention this way)."""
    from pi3_slam_b200 import ops

    torch.manual_seed(S)
    qkv = (torch.randn(S, 3 * H * 64, device="cuda") * 1.3).bfloat16()
    ref = _ref(qkv, 1, S, H)
    parts, lses = [], []
    for lo, hi in zip(splits[:-1], splits[1:]):
        kv = qkv[lo:hi, H * 64:].contiguous()
        o = torch.empty(S, H * 64, device="cuda", dtype=torch.bfloat16)
        l = torch.empty(H, S, device="cuda", dtype=torch.float32)
        ops.attention_q_kv_lse(qkv, kv, o, l, 1, S, hi - lo, H)
        parts.append(o); lses.append(l)
    out = torch.empty(S, H * 64, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(H, S, device="cuda", dtype=torch.float32)
    ops.attention_merge(parts, lses, out, H, lse_out=lse)
    err = (out.float() - ref).norm() / ref.norm()
    assert err.item() < 1e-2, f"merged rel err {err.item():.3e}"
    # merged log-sum-exp equals the fp64 one (log2 domain, scale 1/8)
    q, k, _ = qkv.double().view(S, 3, H, 64).permute(1, 2, 0, 3)
    want = torch.logsumexp(q @ k.transpose(-1, -2) * 0.125, dim=-1) / 0.6931471805599453
    assert (lse.double() - want).abs().max().item() < 2e-2
    # a single part passes through unchanged
    one = torch.empty_like(out)
    ops.attention_merge(parts[:1], lses[:1], one, H)
    assert torch.equal(one, parts[0])


@pytest.mark.parametrize("outlier", [0.0, 60.0])
def test_bounded_reference_softmax_and_its_fallback(outlier):
    """Sequences longer than 3000 keys run the bounded-reference softmax (reference = |q| max|k| scale instead of the
    running maximum). outlier = 0: the bound is tight enough and the fixed reference is adopted. outlier = 60: one key
    with a 60x larger norm makes the Cauchy-Schwarz bound ~500 (log2 units) above the typical scores, so the warps must
    stay on the online path — both against an fp64 softmax on sampled rows."""
    from pi3_slam_b200 import ops

    S, H = 9100, 2
    torch.manual_seed(7)
    qkv = torch.randn(S, 3 * H * 64, device="cuda")
    if outlier:
        qkv[S - 3, H * 64: 2 * H * 64] *= outlier          # the outlier key sits in the LAST tile: the first tile cannot see it
    qkv = qkv.bfloat16()
    out = torch.full((S, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention_qkv(qkv, out, 1, S, H)
    rows = torch.arange(0, S, 13, device="cuda")
    q, k, v = qkv.double().view(S, 3, H, 64).permute(1, 2, 0, 3)
    ref = (torch.softmax(q[:, rows] @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).permute(1, 0, 2).reshape(len(rows), H * 64)
    assert torch.isfinite(out.float()).all()
    err = (out[rows].double() - ref).norm() / ref.norm()
    assert err.item() < 1e-2, f"rel err {err.item():.3e}"
    worst = ((out[rows].double() - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
    assert worst < 3e-2, worst


@pytest.mark.parametrize("q_gain,what", [(1.0, "zero reference"), (9.0, "bound above 100: fixed / online reference")])
def test_prescaled_queries_zero_reference_softmax(q_gain, what):
    """scale = ln 2 announces queries that already carry scale * log2(e) (pi3b_gemm_qkv's q_scale): the long-sequence kernel
    then exponentiates the accumulator as is when |q| max|k| <= 80, and must fall back to a subtracted reference when the
    bound is larger (q_gain = 9: |q| max|k| ~ 130). Both against an fp64 softmax on sampled rows."""
    import math
    from pi3_slam_b200 import ops

    S, H = 9100, 2
    C = H * 64
    torch.manual_seed(11)
    qkv = torch.randn(S, 3 * C, device="cuda")
    qkv[:, :C] *= ops.Q_PRESCALE * q_gain
    qkv = qkv.bfloat16()
    out = torch.full((S, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention_qkv(qkv, out, 1, S, H, scale=ops.ATTN_SCALE_PRESCALED)
    rows = torch.arange(0, S, 13, device="cuda")
    q, k, v = qkv.double().view(S, 3, H, 64).permute(1, 2, 0, 3)
    ref = (torch.softmax(q[:, rows] @ k.transpose(-1, -2) * math.log(2.0), dim=-1) @ v).permute(1, 0, 2).reshape(len(rows), C)
    assert torch.isfinite(out.float()).all()
    err = (out[rows].double() - ref).norm() / ref.norm()
    assert err.item() < 1e-2, f"{what}: rel err {err.item():.3e}"
    # and the log-sum-exp the sequence-parallel merge relies on (log2 domain)
    lse = torch.empty(H, S, device="cuda", dtype=torch.float32)
    kv = qkv[:, C:].contiguous()
    ops.attention_q_kv_lse(qkv, kv, out, lse, 1, S, S, H, scale=ops.ATTN_SCALE_PRESCALED)
    want = torch.logsumexp(q[:, rows] @ k.transpose(-1, -2) * math.log(2.0), dim=-1) / math.log(2.0)
    assert (lse[:, rows].double() - want).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,S,H", [(3, 261, 4), (1, 4099, 2)])
def test_attention_output_stays_inside_its_rows(B, S, H):
    """Ragged last query tiles of both kernels: nothing is written outside the [B*S, H*64] output (canary rows around it)."""
    from pi3_slam_b200 import ops

    torch.manual_seed(S)
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
    pad, sentinel = 256, 777.0
    big = torch.full((B * S + 2 * pad, H * 64), sentinel, device="cuda", dtype=torch.bfloat16)
    out = big[pad:pad + B * S]
    ops.attention_qkv(qkv, out, B, S, H)
    torch.cuda.synchronize()
    assert bool((big[:pad] == sentinel).all()) and bool((big[pad + B * S:] == sentinel).all())
    ref = _ref(qkv, B, S, H)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 1e-2
