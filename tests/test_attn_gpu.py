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
CASES = [(1, 128, 1), (1, 256, 2), (2, 384, 2), (3, 261, 16), (2, 643, 4), (2, 1004, 16), (1, 5, 1), (1, 129, 2),
         (1, 8 * 261, 16), (1, 4099, 3)]


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
