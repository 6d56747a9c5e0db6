// pi3b_gemm: C[M,N] = epilogue(A[M,K] * W[N,K]^T) on tcgen05 tensor cores (sm_100a).
//
// Replaces every nn.Linear on the pi3 chunk forward (reference: pi3/models/layers/attention.py:256-258,
// pi3/models/dinov2/layers/mlp.py:29-31, pi3/models/layers/transformer_head.py:24,46,68,
// pi3/models/layers/camera_head.py:21-23,40-46) and the patch-embed conv-as-GEMM
// (pi3/models/dinov2/layers/patch_embed.py:65,75).
//
// Design: persistent CTAs (one per SM), tile 128 x BLOCK_N x 128 B of K, TMA (SWIZZLE_128B) into a
// multi-stage smem ring, one elected thread issues tcgen05.mma into a double-buffered TMEM accumulator
// (bf16 wide shapes: CTA pairs, cta_group::2, 256 x 256 per pair; otherwise cta_group::1, M=128, N=BLOCK_N),
// eight epilogue warps drain TMEM with tcgen05.ld (one output row per thread) and apply the fused epilogue
// out = [resid +] gamma * act(acc + bias), so the epilogue of tile i overlaps the main loop of tile i+1.
// The epilogue kind is a template parameter (EPI, see gemm_kernel); every global store writes whole 32-byte
// sectors (rows are transposed through a per-warp staging buffer). Operands: bf16 (kind::f16) or fp32 read
// as TF32 (kind::tf32, used for the fp32 heads the reference runs with autocast disabled).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"
#include "qk_math.cuh"

namespace pi3b {

struct GemmParams {
  int M, N, K;
  const float* bias;   // [N] or null
  const float* gamma;  // [N] or null (LayerScale)
  const float* resid;  // fp32 [M, ld_resid] or null (may alias out)
  void* out;           // bf16 or fp32 [M, ldo]
  int ldo, ld_resid;
  int out_f32;  // 1: fp32 output, 0: bf16 output
  int act;      // 0 none, 1 GELU(erf), 2 ReLU
  int vec_ok;   // 1: out/resid/bias/gamma satisfy the 16-byte vector-path alignment
  // output-row remap: GEMM row m lands in row (m / grp_in) * grp_out + grp_off + m % grp_in of out/resid
  // (grp_in == 0: identity). Used to scatter patch rows between the per-frame special tokens.
  int grp_in, grp_out, grp_off;
  // fused QKV epilogue (pi3b_gemm_qkv): columns [0, qk_cols) are the q then the k heads (64 channels each, qk_cols / 2
  // per third), the rest (v) is stored as is. qk_mode bit 0: LayerNorm(64) with qk_norm = fp32 [4][64] (q weight, q bias,
  // k weight, k bias); bit 1: RoPE-2D with cs_table fp32 [pos][16] (cos, sin) and the token geometry below.
  int qk_mode, qk_cols;
  const float* qk_norm;
  const float2* cs_table;
  float qk_eps;
  int cs_rows;     // rows of cs_table (positions 0 .. cs_rows - 1)
  float q_scale;   // queries (columns [0, qk_cols / 2)) are multiplied by this before the final rounding (1 = none)
  int tokens_per_frame, n_special, patch_w;
};

constexpr int kBlockM = 128;
constexpr int kNumEpiWarps = 8;
constexpr int kGemmThreads = 64 + kNumEpiWarps * 32;

// RoPE table rows the fused QKV epilogue can keep in shared memory (144-byte row stride: the 16-byte reads of eight
// consecutive positions fall into distinct bank groups); longer tables are read from global memory
constexpr int kCsSmemRows = 60;
constexpr int kCsRowBytes = 144;

template <int BLOCK_N, int STAGES, bool TWOSM = false, bool QKEPI = false>
struct GemmSmem {
  static constexpr int A_BYTES = kBlockM * 128;
  static constexpr int B_BYTES = (TWOSM ? BLOCK_N / 2 : BLOCK_N) * 128;   // 2-SM: each CTA of the pair holds half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // per-epilogue-warp transpose staging: 32 rows x 16 fp32 (+4 pad so rows stay 16-byte aligned and the
  // 128-bit accesses of 8 consecutive rows fall into distinct bank groups)
  static constexpr int STG_STRIDE = 20;
  static constexpr int STG_BYTES_PER_WARP = 32 * STG_STRIDE * 4;
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  // per-epilogue-warp copy of the bias / gamma slice of the current tile (fetched while the main loop runs)
  static constexpr int BG_OFFSET = STG_OFFSET + kNumEpiWarps * STG_BYTES_PER_WARP;
  // (the fused QKV epilogue has no gamma: its half holds the RoPE table instead)
  static constexpr int BG_BYTES_PER_WARP = (QKEPI ? 1 : 2) * (BLOCK_N / 2) * 4;
  // q/k LayerNorm parameters of the fused QKV epilogue: [q w | q b | k w | k b] x 64 fp32
  static constexpr int QK_OFFSET = BG_OFFSET + kNumEpiWarps * BG_BYTES_PER_WARP;
  // the (cos, sin) table of the fused QKV epilogue: a row is 16 frequencies x 2 fp32 = 128 bytes, padded to 144
  static constexpr int CS_OFFSET = QK_OFFSET + 4 * 64 * 4;
  static constexpr int BAR_OFFSET = CS_OFFSET + (QKEPI ? kCsSmemRows * kCsRowBytes : 0);
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // +1024: manual alignment slack
  static_assert(TOTAL <= 227 * 1024, "shared-memory budget of one CTA");
};

// Exact (erf) GELU of nn.GELU() (mlp.py:29), 0.5 x (1 + erf(x / sqrt 2)), with erf from Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7): 15 instructions and 2 MUFU ops instead of libm erff's ~35 — the fc1 epilogue (128 values per
// thread per tile) was issue-bound on erff and longer than the K = 1024 main loop. Measured max |error| vs fp64:
// 4.7e-7 absolute, 2e-5 relative for |gelu| > 1e-2; the output is rounded to bf16 (4e-3) right after.
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * a * a));
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), fmaf(-p, e, 1.0f), hx);      // 0.5 x + 0.5 |x| erf(|x| / sqrt 2)
}


// Two GELUs at once with the polynomial on packed fp32x2 FMAs (FFMA2/FMUL2): same operations in the same order and rounding
// as gelu_erf, so the results are bit-identical; 12 packed + 4 ALU + 4 MUFU instructions per pair instead of 30.
__device__ __forceinline__ uint64_t gelu_pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t gelu_ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t gelu_fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#define GELU_K2(c) gelu_pack2(c, c)
__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) {
  const uint64_t x = gelu_pack2(x0, x1);
  const uint64_t a = gelu_fmul2(gelu_pack2(fabsf(x0), fabsf(x1)), GELU_K2(0.70710678118654752f));
  const uint64_t d = gelu_ffma2(GELU_K2(0.3275911f), a, GELU_K2(1.0f));
  float d0, d1, t0, t1, e0, e1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t = gelu_pack2(t0, t1);
  uint64_t q = gelu_ffma2(GELU_K2(1.061405429f), t, GELU_K2(-1.453152027f));
  q = gelu_ffma2(q, t, GELU_K2(1.421413741f));
  q = gelu_ffma2(q, t, GELU_K2(-0.284496736f));
  q = gelu_ffma2(q, t, GELU_K2(0.254829592f));
  q = gelu_fmul2(q, t);
  const uint64_t arg = gelu_fmul2(gelu_fmul2(GELU_K2(-1.4426950408889634f), a), a);
  float g0, g1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(g0), "=f"(g1) : "l"(arg));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const uint64_t hx = gelu_fmul2(GELU_K2(0.5f), x);
  float h0, h1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(hx));
  // fma(-p, e, 1): the negation is exact, so -(p) * e + 1 == fma(p, -e, 1)
  const uint64_t w = gelu_ffma2(q, gelu_pack2(-e0, -e1), GELU_K2(1.0f));
  const uint64_t r = gelu_ffma2(gelu_pack2(fabsf(h0), fabsf(h1)), w, hx);
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(r));
}
#undef GELU_K2

// bf16 epilogue store of one 32-row x 32-column chunk (64 bytes per row; packed[16] = this lane's row): the rows are
// transposed through the warp's staging buffer so that every store instruction writes whole 32-byte sectors (8 rows x 64
// contiguous bytes) instead of 32 rows x 16 bytes — half sectors that the L2 has to merge. Measured on the qkv GEMM
// (M = 100 400): 0.513 -> 0.442 ms = 1428 TFLOP/s, past cuBLAS' 0.463 ms (profiles/r02_gemm_epilogue_ab.log).
__device__ __forceinline__ void store_bf16_chunk_coalesced(uint32_t stg, const uint32_t* packed, const GemmParams& p,
                                                           int row0, int col0, int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 80 + j * 16), "r"(packed[4 * j]),
                 "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3]) : "memory");
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + (lane >> 2);
    const int mi = row0 + rr;
    uint32_t a, b, c, d;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "r"(stg + rr * 80 + (lane & 3) * 16));
    if (mi < p.M) {
      const int mo = p.grp_in ? (mi / p.grp_in) * p.grp_out + p.grp_off + mi % p.grp_in : mi;
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + size_t(mo) * p.ldo + col0 + (lane & 3) * 8) =
          make_uint4(a, b, c, d);
    }
  }
  __syncwarp();
}

// CL = 1: every CTA loads its own A (128 x K) and B (BLOCK_N x K) tiles.
// CL = 2: thread-block clusters of two CTAs that work on vertically adjacent output tiles (same n block, m blocks
//         2i and 2i+1). Each CTA loads its own A tile but only HALF of the shared B tile and TMA-multicasts that half
//         into both CTAs, so a CTA pulls 32 KB instead of 48 KB through L2 per k block. Reason (ncu, qkv-shaped GEMM):
//         the MMA warp spent 58 % of its time waiting on the TMA full barriers — with 192 KB of smem the ring holds
//         just enough bytes in flight for 96 B/clk/SM at the loaded L2 latency; at 64 B/clk/SM it does.
// TWOSM (CL = 2 only): the two CTAs of the cluster form a CTA PAIR and execute one tcgen05.mma.cta_group::2 of M = 256:
//         a 256 x BLOCK_N output tile per pair, each CTA loading its 128 rows of A and HALF of the B tile (nothing is
//         duplicated: 32 KB per CTA and k block through L2 and into shared memory instead of 48 KB, and the tensor core
//         reads half of B from its own shared memory, the other half from the peer's). The even CTA issues all MMAs and
//         owns the full and accumulator-empty barriers; see ptx.cuh. This is the configuration cuBLAS picks for these
//         shapes on B200 (nvjet 256x256_64x4 2cta, profiles/r02_gemm_full_metrics.txt).
// EPI selects the epilogue compiled into the instantiation (each keeps its own register budget and code size):
//   0 general (bias / GELU / ReLU / LayerScale / residual, bf16 or fp32 out, ragged N, row regrouping)
//   1 fused QKV: bias -> bf16 -> q/k LayerNorm(64) -> RoPE-2D (pi3b_gemm_qkv)
//   2 streamlined fp32 read-modify-write of the residual stream (whole chunks, no regrouping, no GELU)
template <int BLOCK_N, int STAGES, bool TF32, int CL, int EPI, bool TWOSM = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr bool QK = EPI == 1, FAST = EPI == 2;
  static_assert(!TWOSM || (CL == 2 && !TF32), "the CTA-pair kernel is the bf16 2-CTA-cluster variant");
  using S = GemmSmem<BLOCK_N, STAGES, TWOSM, EPI == 1>;
  constexpr int ELEM = TF32 ? 4 : 2;
  constexpr int BLOCK_K = 128 / ELEM;  // elements per 128-byte smem row
  constexpr int UMMA_K = 32 / ELEM;    // 16 (bf16) / 8 (tf32): 32 bytes of K per instruction
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_k = (p.K + BLOCK_K - 1) / BLOCK_K;
  // work units: CL = 1 -> output tiles; CL = 2 -> pairs of vertically adjacent tiles, one per CTA of the cluster
  // (both CTAs walk the same unit sequence, so their pipelines stay in lock step; an odd last m block gives the second
  // CTA an all-out-of-range tile: TMA zero-fills it and the epilogue stores nothing)
  const int cta_rank = CL == 2 ? int(cluster_ctarank()) : 0;
  const int num_tiles = ((num_m + CL - 1) / CL) * num_n;
  const int unit0 = int(blockIdx.x) / CL;
  const int unit_stride = int(gridDim.x) / CL;
  auto tile_m = [&](int unit) { return (unit / num_n) * CL + cta_rank; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      // the slot is reused only when every CTA that receives multicast data has consumed it (CTA pair: one multicast
      // commit of the pair's MMAs arrives once on each CTA)
      mbar_init(&empty_bar[i], TWOSM ? 1 : CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      // CTA pair: the epilogue warps of BOTH CTAs hand the accumulator stage back to the leader's barrier
      mbar_init(&tmem_empty[i], TWOSM ? 2 * kNumEpiWarps : kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (TWOSM) {
      tmem_alloc_2sm(tmem_slot, TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  if (QK && (p.qk_mode & 1) && threadIdx.x < 256)
    reinterpret_cast<float*>(smem + S::QK_OFFSET)[threadIdx.x] = __ldg(p.qk_norm + threadIdx.x);
  // the RoPE table into shared memory: every head of every row needs 2 x 128 bytes of it, and read through the L1/L2 path
  // right before use those loads were the epilogue warps' largest stall (long scoreboard 17 % of the kernel's samples)
  if (QK && (p.qk_mode & 2) && p.cs_rows <= kCsSmemRows) {
    for (int i = threadIdx.x; i < p.cs_rows * 8; i += kGemmThreads) {
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.cs_table) + i);
      *reinterpret_cast<float4*>(smem + S::CS_OFFSET + (i >> 3) * kCsRowBytes + (i & 7) * 16) = c4;
    }
  }
  tc_fence_before();
  if (CL == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything remote touches them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the tile schedule, one elected lane issues (operands stay uniform) =====
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      const int m_blk = tile_m(tile), n_blk = tile % num_n;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          if (TWOSM) {
            // CTA pair: both CTAs' bytes are credited to the leader's barrier, which therefore expects twice a stage
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
            tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * kBlockM, kEvictNormal);
            tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N + cta_rank * (BLOCK_N / 2), kEvictLast);
          } else {
          mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          // activations stream through once; the (small) weight panel is reused by every m block
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * kBlockM, kEvictNormal);
          if (CL == 2)  // my half of the B tile, delivered to both CTAs (the peer sends the other half)
            tma_load_2d_mcast(sb + cta_rank * (S::B_BYTES / 2), &tmB, &full_bar[stage], kb * BLOCK_K,
                              n_blk * BLOCK_N + cta_rank * (BLOCK_N / 2), uint16_t(3), kEvictLast);
          else
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N, kEvictLast);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && (!TWOSM || cta_rank == 0)) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma / commit =====
    constexpr uint32_t idesc = make_idesc(TF32 ? 2u : 1u, TWOSM ? 2 * kBlockM : kBlockM, BLOCK_N, 0, 0);
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem + S::A_BYTES), 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t soff = uint64_t((stage * S::STAGE_BYTES) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (TWOSM)
              umma_f16_ss_2sm(d_tmem, a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (kb | k) != 0);
            else if (TF32)
              umma_tf32_ss(d_tmem, a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (kb | k) != 0);
            else
              umma_f16_ss(d_tmem, a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (kb | k) != 0);
          }
          // frees the smem slot once these MMAs have read it (in both CTAs of a cluster: the peer multicasts into it /
          // the pair's MMA reads both CTAs' slots)
          if (TWOSM) umma_commit_2sm(&empty_bar[stage], uint16_t(3));
          else if (CL == 2) umma_commit_mcast(&empty_bar[stage], uint16_t(3));
          else umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) {  // accumulator complete -> epilogue (of both CTAs of a pair)
            if (TWOSM) umma_commit_2sm(&tmem_full[acc], uint16_t(3));
            else umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 2) {
    // ===== epilogue warps: TMEM -> registers -> fused epilogue -> HBM =====
    const int ew = warp - 2;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
    const int half = ew >> 2;                // which half of the BLOCK_N columns
    constexpr int COLS_PER_WARP = BLOCK_N / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      const int m_blk = tile_m(tile), n_blk = tile % num_n;
      const int m_in = m_blk * kBlockM + quarter * 32 + lane;
      const bool row_ok = m_in < p.M;
      const int m = p.grp_in ? (m_in / p.grp_in) * p.grp_out + p.grp_off + m_in % p.grp_in : m_in;
      // bias / gamma of this warp's columns go to smem BEFORE the accumulator is ready: their global-load latency
      // (26 % of the epilogue warps' stall samples when loaded per chunk) hides under the main loop
      float* bias_s = reinterpret_cast<float*>(smem + S::BG_OFFSET + ew * S::BG_BYTES_PER_WARP);
      float* gamma_s = bias_s + COLS_PER_WARP;
      if (p.vec_ok && lane * 4 < COLS_PER_WARP) {
        const int mycol = n_blk * BLOCK_N + half * COLS_PER_WARP + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (mycol + 3 < p.N) {
          if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + mycol));
          if (p.gamma) g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + mycol));
        }
        *reinterpret_cast<float4*>(bias_s + lane * 4) = b4;
        if (!QK) *reinterpret_cast<float4*>(gamma_s + lane * 4) = g4;   // (no gamma half in the fused QKV layout)
      }
      __syncwarp();
      // The accumulator buffer goes back to the MMA warp as soon as this warp has its last TMEM read in registers (not
      // after the stores): the next-but-one tile's main loop never waits for an epilogue's memory traffic.
      bool released = false;
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (TWOSM) mbar_arrive_cluster(&tmem_empty[acc], 0);   // the leader's MMA warp waits for both CTAs' epilogues
          else mbar_arrive(&tmem_empty[acc]);
        }
        released = true;
      };
      // EPI = 2: fp32 outputs of whole 32-column chunks without row regrouping (every residual-stream GEMM of the trunk;
      // gemm_dispatch checks the conditions): per-tile row pointers instead of per-row index arithmetic, the residual prefetched TWO
      // chunks ahead (2 x 4 KB in flight per warp, 64 KB per SM: one chunk ahead left the warps waiting on DRAM latency
      // for most of each chunk — long-scoreboard 40 % of the samples of the proj GEMM), stores issued as the sums appear.
      constexpr int NCH = COLS_PER_WARP / 32;
      const int sub_r = lane >> 2, sub_c = (lane & 3) * 4;
      const int row_first = m_blk * kBlockM + quarter * 32 + sub_r;
      const int col_first = n_blk * BLOCK_N + half * COLS_PER_WARP + sub_c;
      const float* rp = (FAST && p.resid) ? p.resid + size_t(row_first) * p.ld_resid + col_first : nullptr;
      const size_t rstep = size_t(8) * p.ld_resid;
      float4 rqA[8], rqB[8];
      auto loadq = [&](float4* q, int c) {
        if (rp == nullptr || c >= NCH) return;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          if (row_first + 8 * it < p.M) {
            q[it] = __ldcg(reinterpret_cast<const float4*>(rp + it * rstep + c * 32));
            q[4 + it] = __ldcg(reinterpret_cast<const float4*>(rp + it * rstep + c * 32 + 16));
          }
        }
      };
      loadq(rqA, 0);   // the residual does not depend on the accumulator: requested before the wait, under the main loop
      loadq(rqB, 1);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (FAST) {
        float* op = static_cast<float*>(p.out) + size_t(row_first) * p.ldo + col_first;
        const size_t ostep = size_t(8) * p.ldo;
        const uint32_t stg = smem_u32(smem + S::STG_OFFSET + ew * S::STG_BYTES_PER_WARP);
        auto chunk = [&](int c, const float4* q) {
          uint32_t r[32];
          tmem_ld_x32(tmem_base + (uint32_t(quarter * 32) << 16) + acc * BLOCK_N + half * COLS_PER_WARP + c * 32, r);
          tmem_ld_wait();
          if (c == NCH - 1) release_acc();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.gamma) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 g = *reinterpret_cast<const float4*>(gamma_s + c * 32 + j);
              v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
            }
          }
          // transpose through the warp's staging buffer (two halves of 16 columns): every global access below is 64
          // contiguous bytes per row, 8 rows per instruction
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (lane * S::STG_STRIDE + 4 * j) * 4),
                           "f"(v[h * 16 + 4 * j]), "f"(v[h * 16 + 4 * j + 1]), "f"(v[h * 16 + 4 * j + 2]),
                           "f"(v[h * 16 + 4 * j + 3]) : "memory");
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              float4 x = qk_lds128(stg + ((it * 8 + sub_r) * S::STG_STRIDE + sub_c) * 4);
              if (rp != nullptr) {
                const float4 qq = q[h * 4 + it];
                x.x += qq.x; x.y += qq.y; x.z += qq.z; x.w += qq.w;
              }
              if (row_first + 8 * it < p.M) *reinterpret_cast<float4*>(op + it * ostep + c * 32 + h * 16) = x;
            }
            __syncwarp();
          }
        };
#pragma unroll 1
        for (int c = 0; c < NCH; c += 2) {
          chunk(c, rqA);
          loadq(rqA, c + 2);
          chunk(c + 1, rqB);
          loadq(rqB, c + 3);
        }
      } else if (QK) {
        // ===== fused QKV epilogue (QK instantiations only: the plain kernels keep their register budget): bias -> bf16 -> [LayerNorm(64)] -> [RoPE-2D] -> bf16, one head (64 columns) at a time;
        // the thread owns its row, so the head's statistics and rotation pairs are all in its own registers. Same
        // arithmetic as qknorm_rope_kernel (qk_math.cuh): the result is bit-identical to GEMM + stand-alone pass, without
        // the 822 MB read-modify-write of the q and k thirds.
        int pos_y = 0, pos_x = 0;
        if (p.qk_mode & 2) qk_row_position(m_in, p.tokens_per_frame, p.n_special, p.patch_w, pos_y, pos_x);
        const uint32_t qk_saddr = smem_u32(smem + S::QK_OFFSET);
        const uint32_t cs_saddr = p.cs_rows <= kCsSmemRows ? smem_u32(smem + S::CS_OFFSET) : 0u;
        const uint32_t bias_saddr = smem_u32(bias_s);
#pragma unroll 1
        for (int hd = 0; hd < COLS_PER_WARP / 64; ++hd) {
          const int col_in_tile = half * COLS_PER_WARP + hd * 64;
          const int n0 = n_blk * BLOCK_N + col_in_tile;
          if (n0 >= p.N) break;  // warp-uniform
          uint32_t r[64];
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BLOCK_N + col_in_tile;
          tmem_ld_x32(taddr, r);
          tmem_ld_x32(taddr + 32, r + 32);
          tmem_ld_wait();
          if (hd == COLS_PER_WARP / 64 - 1) release_acc();
          float v[64];
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 b = qk_lds128(bias_saddr + (hd * 64 + j) * 4);
            v[j] = __fadd_rn(__uint_as_float(r[j]), b.x);
            v[j + 1] = __fadd_rn(__uint_as_float(r[j + 1]), b.y);
            v[j + 2] = __fadd_rn(__uint_as_float(r[j + 2]), b.z);
            v[j + 3] = __fadd_rn(__uint_as_float(r[j + 3]), b.w);
            qk_round_bf16_x2(v[j], v[j + 1]);     // what the plain epilogue would have stored: bf16(acc + bias)
            qk_round_bf16_x2(v[j + 2], v[j + 3]);
          }
          if (n0 < p.qk_cols) {
            const uint32_t norm_saddr = (p.qk_mode & 1) ? qk_saddr + (n0 >= (p.qk_cols >> 1) ? 512u : 0u) : 0u;
            qk_head_inplace(v, norm_saddr, p.qk_eps, (p.qk_mode & 2) ? p.cs_table : nullptr, pos_y, pos_x, cs_saddr);
            if (p.q_scale != 1.0f && n0 < (p.qk_cols >> 1)) {
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] = qk_scale_out(v[j], p.q_scale);
            }
          }
          {
            const uint32_t stg = smem_u32(smem + S::STG_OFFSET + ew * S::STG_BYTES_PER_WARP);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {   // the head's 128 bytes per row in two 64-byte halves
              uint32_t packed[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[32 * hh + 2 * j], v[32 * hh + 2 * j + 1]);
              store_bf16_chunk_coalesced(stg, packed, p, m_blk * kBlockM + quarter * 32, n0 + 32 * hh, lane);
            }
          }
        }
      } else
#pragma unroll 1
      for (int c = 0; c < (FAST ? 0 : COLS_PER_WARP / 32); ++c) {
        const int col_in_tile = half * COLS_PER_WARP + c * 32;
        const int n0 = n_blk * BLOCK_N + col_in_tile;
        if (n0 >= p.N) break;  // warp-uniform
        const int nvalid = min(32, p.N - n0);
        uint32_t r[32];
        tmem_ld_x32(tmem_base + (uint32_t(quarter * 32) << 16) + acc * BLOCK_N + col_in_tile, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (nvalid == 32 && p.vec_ok) {
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_erf_x2(v[j], v[j + 1]);
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.gamma) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 g = *reinterpret_cast<const float4*>(gamma_s + c * 32 + j);
              v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
            }
          }
          if (p.resid && row_ok) {
            // (row-regrouped or bf16-output residual GEMMs only — the patch embedding, once per forward; the residual-stream
            // GEMMs of the trunk take the streamlined path above)
            const float4* rr = reinterpret_cast<const float4*>(p.resid + size_t(m) * p.ld_resid + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q = __ldcg(rr + j);
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          }
          if (p.out_f32 && row_ok) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + size_t(m) * p.ldo + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (!p.out_f32) {
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            store_bf16_chunk_coalesced(smem_u32(smem + S::STG_OFFSET + ew * S::STG_BYTES_PER_WARP), packed, p,
                                       m_blk * kBlockM + quarter * 32, n0, lane);
          }
        } else if (row_ok) {
          // ragged N tail (N = 588 / 196 / 9 / 3 heads): scalar path
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              float x = v[j];
              if (p.bias) x += __ldg(p.bias + n0 + j);
              if (p.act == 1) x = gelu_erf(x);
              else if (p.act == 2) x = fmaxf(x, 0.f);
              if (p.gamma) x *= __ldg(p.gamma + n0 + j);
              if (p.resid) x += p.resid[size_t(m) * p.ld_resid + n0 + j];
              if (p.out_f32) static_cast<float*>(p.out)[size_t(m) * p.ldo + n0 + j] = x;
              else static_cast<__nv_bfloat16*>(p.out)[size_t(m) * p.ldo + n0 + j] = __float2bfloat16(x);
            }
          }
        }
      }
      if (!released) release_acc();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  if (CL == 2) cluster_sync_all();  // nobody leaves while the peer may still multicast into / arrive on this CTA
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (TWOSM) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int STAGES, bool TF32, int CL, int EPI = 0, bool TWOSM = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                       cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N, STAGES, TWOSM, EPI == 1>;
  auto kern = gemm_kernel<BLOCK_N, STAGES, TF32, CL, EPI, TWOSM>;
  static DeviceOnce configured;   // one per template instantiation; the attribute is per device
  const int dev = current_device();
  if (configured.first_call_on(dev)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return check_launch("gemm: cudaFuncSetAttribute");
    configured.done(dev);
  }
  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int units = ((num_m + CL - 1) / CL) * ((p.N + BLOCK_N - 1) / BLOCK_N);
  int grid = units * CL < num_sms ? units * CL : num_sms;
  grid -= grid % CL;
  if (CL == 1) {
    kern<<<grid, kGemmThreads, S::TOTAL, stream>>>(tmA, tmB, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p) != cudaSuccess) return check_launch("gemm cluster launch");
  }
  return check_launch("gemm launch");
}

}  // namespace pi3b

using namespace pi3b;

// shared by pi3b_gemm and pi3b_gemm_qkv: tile choice, tensor maps, launch
static int gemm_dispatch(void* stream, int in_dtype, const void* A, int64_t lda, const void* W, int64_t ldw, GemmParams p) {
  const int M = p.M, N = p.N, K = p.K;
  const int esz = in_dtype == 0 ? 2 : 4;
  const int num_sms = num_sms_of(current_device());
  const int block_k = 128 / esz;
  // Use the wide tile when it still fills the machine; otherwise the narrow tile for more CTAs.
  const int tiles256 = ((M + 127) / 128) * ((N + 255) / 256);
  const bool wide = tiles256 >= num_sms && N > 128;
  // wide tiles run as 2-CTA clusters sharing the B tile through TMA multicast (PI3B_GEMM_CLUSTER=0 disables, for A/B runs)
  static const bool cluster_ok = []() { const char* e = getenv("PI3B_GEMM_CLUSTER"); return !(e && e[0] == '0'); }();
  const bool cluster = wide && cluster_ok && M > 128;
  // bf16 wide tiles: CTA pairs (cta_group::2, 256 x 256 per pair); PI3B_GEMM_2SM=0 falls back to the multicast clusters
  static const bool twosm_ok = []() { const char* e = getenv("PI3B_GEMM_2SM"); return !(e && e[0] == '0'); }();
  const bool twosm = cluster && twosm_ok && in_dtype == 0;

  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, in_dtype, (uint64_t)M, (uint64_t)K, (uint64_t)lda * esz, 128, block_k);
  if (rc) return rc;
  // B box: the whole 256-row tile (single CTA), or the half each CTA of a multicast cluster / CTA pair loads
  rc = make_tmap_2d(&tmB, W, in_dtype, (uint64_t)N, (uint64_t)K, (uint64_t)ldw * esz, (wide && !cluster) ? 256 : 128, block_k);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.qk_mode != 0) {   // fused QKV epilogue (bf16 operands only)
    if (twosm) return launch_gemm<256, 6, false, 2, 1, true>(tmA, tmB, p, num_sms, st);
    if (cluster) return launch_gemm<256, 4, false, 2, 1>(tmA, tmB, p, num_sms, st);
    return wide ? launch_gemm<256, 4, false, 1, 1>(tmA, tmB, p, num_sms, st)
                : launch_gemm<128, 6, false, 1, 1>(tmA, tmB, p, num_sms, st);
  }
  if (in_dtype == 0) {
    // fp32 outputs (the residual stream's read-modify-write, LayerScale fused) of whole column chunks: streamlined epilogue
    if (twosm && p.out_f32 && p.vec_ok && p.grp_in == 0 && p.act != 1 && N % 256 == 0)
      return launch_gemm<256, 6, false, 2, 2, true>(tmA, tmB, p, num_sms, st);
    if (twosm) return launch_gemm<256, 6, false, 2, 0, true>(tmA, tmB, p, num_sms, st);
    if (cluster) return launch_gemm<256, 4, false, 2>(tmA, tmB, p, num_sms, st);
    return wide ? launch_gemm<256, 4, false, 1>(tmA, tmB, p, num_sms, st) : launch_gemm<128, 6, false, 1>(tmA, tmB, p, num_sms, st);
  } else {
    if (cluster) return launch_gemm<256, 4, true, 2>(tmA, tmB, p, num_sms, st);
    return wide ? launch_gemm<256, 4, true, 1>(tmA, tmB, p, num_sms, st) : launch_gemm<128, 6, true, 1>(tmA, tmB, p, num_sms, st);
  }
}

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// See include/pi3b200.h for the contract.
extern "C" int pi3b_gemm(void* stream, int in_dtype, const void* A, int64_t lda, const void* W, int64_t ldw, int M,
                         int N, int K, const float* bias, const float* gamma, const float* resid, int64_t ld_resid,
                         void* out, int64_t ldo, int out_f32, int act, int grp_in, int grp_out, int grp_off) {
  PI3B_REQUIRE(A && W && out, "pi3b_gemm: null pointer");
  PI3B_REQUIRE(M > 0 && N > 0 && K > 0, "pi3b_gemm: bad shape M=%d N=%d K=%d", M, N, K);
  PI3B_REQUIRE(in_dtype == 0 || in_dtype == 1, "pi3b_gemm: in_dtype must be 0 (bf16) or 1 (fp32->tf32)");
  PI3B_REQUIRE(act >= 0 && act <= 2, "pi3b_gemm: act must be 0/1/2");
  PI3B_REQUIRE(grp_in == 0 || (grp_in > 0 && grp_out >= grp_in + grp_off && grp_off >= 0), "pi3b_gemm: bad row remap");
  const int esz = in_dtype == 0 ? 2 : 4;
  PI3B_REQUIRE((lda * esz) % 16 == 0 && (ldw * esz) % 16 == 0, "pi3b_gemm: lda/ldw must give 16-byte row pitch");
  PI3B_REQUIRE(lda >= K && ldw >= K, "pi3b_gemm: lda/ldw < K");
  PI3B_REQUIRE(ldo >= N && (!resid || ld_resid >= N), "pi3b_gemm: ldo/ld_resid < N");
  const int vec_ok = (ldo % (out_f32 ? 4 : 8) == 0) && al16(out) && (!resid || (ld_resid % 4 == 0 && al16(resid))) &&
                     (!bias || al16(bias)) && (!gamma || al16(gamma));
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.gamma = gamma; p.resid = resid;
  p.out = out; p.ldo = (int)ldo; p.ld_resid = (int)ld_resid;
  p.out_f32 = out_f32; p.act = act; p.vec_ok = vec_ok;
  p.grp_in = grp_in; p.grp_out = grp_out; p.grp_off = grp_off;
  return gemm_dispatch(stream, in_dtype, A, lda, W, ldw, p);
}

extern "C" int pi3b_gemm_qkv(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int num_heads,
                             int K, const float* bias, void* out, int64_t ldo, const float* qk_norm, float eps,
                             const float* cs_table, int table_rows, int tokens_per_frame, int n_special, int patch_w,
                             float q_scale) {
  PI3B_REQUIRE(A && W && out, "pi3b_gemm_qkv: null pointer");
  PI3B_REQUIRE(q_scale > 0.f, "pi3b_gemm_qkv: q_scale must be positive (1 = none)");
  PI3B_REQUIRE(M > 0 && num_heads > 0 && K > 0, "pi3b_gemm_qkv: bad shape");
  PI3B_REQUIRE(qk_norm || cs_table, "pi3b_gemm_qkv: neither q/k norm nor RoPE requested (use pi3b_gemm)");
  const int N = 3 * 64 * num_heads;
  PI3B_REQUIRE((lda * 2) % 16 == 0 && (ldw * 2) % 16 == 0 && lda >= K && ldw >= K, "pi3b_gemm_qkv: bad lda/ldw");
  PI3B_REQUIRE(ldo >= N && ldo % 8 == 0 && al16(out) && (!bias || al16(bias)) && (!qk_norm || al16(qk_norm)),
               "pi3b_gemm_qkv: out / bias / qk_norm must be 16-byte aligned, ldo a multiple of 8");
  if (cs_table) {
    PI3B_REQUIRE(al16(cs_table) && tokens_per_frame > n_special && n_special >= 0 && patch_w > 0,
                 "pi3b_gemm_qkv: bad RoPE geometry");
    const int ph = (tokens_per_frame - n_special + patch_w - 1) / patch_w;
    PI3B_REQUIRE(table_rows > (ph > patch_w ? ph : patch_w), "pi3b_gemm_qkv: cos/sin table too short");
  }
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias;
  p.out = out; p.ldo = (int)ldo;
  p.vec_ok = 1;
  p.qk_mode = (qk_norm ? 1 : 0) | (cs_table ? 2 : 0);
  p.qk_cols = 2 * 64 * num_heads;
  p.qk_norm = qk_norm;
  p.cs_table = reinterpret_cast<const float2*>(cs_table);
  p.qk_eps = eps;
  p.cs_rows = table_rows;
  p.q_scale = q_scale;
  p.tokens_per_frame = tokens_per_frame > 0 ? tokens_per_frame : 1;
  p.n_special = n_special;
  p.patch_w = patch_w > 0 ? patch_w : 1;
  return gemm_dispatch(stream, 0, A, lda, W, ldw, p);
}
