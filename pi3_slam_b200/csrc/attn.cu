// pi3b_attn_fwd_hd64: softmax(Q K^T / 8) V for head_dim 64, bf16 in / bf16 out, no mask.
//
// Replaces F.scaled_dot_product_attention on the pi3 chunk forward (reference:
// pi3/models/layers/attention.py:102-104 and :336-338) for both regimes of Pi3.decode
// (pi3/models/pi3.py:156-166): many frame-local sequences (S = T ~ 1k) and one chunk-global
// sequence (S = N*T ~ 100k), which is ~75 % of the FLOPs of a chunk.
//
// Two kernels share one protocol (FA4-style, hand-written tcgen05):
//   * a CTA owns several 128-row Q tiles of one (sequence, head); warp 0 streams 64-row K/V tiles with TMA into an
//     8-stage smem ring; issuer warps (one elected lane) issue S = Q K^T (M128 N64 K64, both operands K-major in smem)
//     and O += P V (A = P from TMEM, B = V MN-major in smem); S, P and O live in TMEM;
//   * softmax warps own one query row per thread (tcgen05.ld 32x32b), so row max / row sum need no shuffles; online
//     softmax in the exp2 domain with lazy rescaling of O (only when the row max grows by more than 2^8), P written
//     back to TMEM as packed bf16;
//   * at head_dim 64 a 128x64 score tile costs 256 tensor clocks (QK^T + PV) but 512 MUFU clocks (16 ex2/clk/SM), so
//     the softmax side is the roofline. A fixed share of the column pairs therefore computes its exponentials on the
//     FMA pipe: Cody-Waite split + a minimax polynomial evaluated on PACKED fp32x2 (FFMA2 / FADD2), i.e. half the
//     issue slots per element of a scalar polynomial; its accuracy (degree 2: 2^-9.2) is matched to the bf16 rounding
//     of P (2^-9) that follows.
//
// attn_fwd_kernel_t3 (kv_len > 3000: the chunk-global layers, MoGe's 3.6k-token encoder): THREE Q tiles per CTA so that three softmax warps share
// each scheduler; S 3x64, O 3x64, P 3x32 TMEM columns (480 of 512); ONE ISSUER WARP PER TILE (with one issuer and grouped
// waits the tiles ran in lockstep, their exponential phases coincided on the MUFU pipe: 645 vs 888 TFLOP/s,
// profiles/r01_attn_rev8_ab.log); a softmax warp hands S back as soon as the scores are in registers (s_free) so QK(j+1)
// runs under this step's exponentials; setmaxnreg 48/152 (budgets must add up to the launch allocation).
//
// attn_fwd_kernel_t4 (kv_len <= 3000, frame-local layers): FOUR Q tiles, P written in place over the first 32 columns
// of its own S tile (S/P 4x64 + O 4x64 = 512 columns), four softmax warps per scheduler, two issuer warps serving two
// tiles each in a fixed order: PV(j) then QK(j+1) back to back (the in-order tensor pipe makes the new scores overwrite P
// only after PV has read it). The softmax step is two passes over 32-column halves (row max, then exponentials, S re-read
// from TMEM) so that it fits the 96 registers a 640-thread CTA leaves per thread. +9 % at S = 1004, -1.5 % at S = 100400.
//
// Optional second output `lse` (log2-domain log-sum-exp per query row and head) lets the caller merge attention over
// disjoint key/value sets (pi3b_attn_merge): the sequence-parallel chunk-global layer runs the local K/V block while the
// remote blocks are still in flight over NVLink.
//
// History of measured variants (round 1: profiles/r01_attn_*.log; round 2: profiles/r02_attn_*.log) is in DESIGN.md.
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>

#include <algorithm>

namespace pi3b {

// ---- exponential mix (compile-time; A/B'd through -D on the same box) -----------------------------------------------
// Of every 8 column pairs, PI3B_POLY_PAIRS compute BOTH exponentials with the packed polynomial (share = PAIRS/8).
#ifndef PI3B_POLY_PAIRS
#define PI3B_POLY_PAIRS 3
#endif
#ifndef PI3B_POLY_PAIRS_ZERO   // the same share for the zero-reference steps of the three-tile kernel
#define PI3B_POLY_PAIRS_ZERO 4
#endif
#ifndef PI3B_POLY_DEG
#define PI3B_POLY_DEG 2
#endif



constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kKvBytes = 64 * 128;     // 64 rows x 64 bf16

struct AttnParams {
  int num_seqs, q_len, kv_len, num_heads;  // per sequence: q_len query rows attend to kv_len key/value rows
  int q_col0, k_col0, v_col0;              // element column of head 0 inside each operand's rows
  __nv_bfloat16* out;
  int ldo;
  float scale_log2;  // softmax scale * log2(e)
  int unit_scale;    // scale_log2 == 1 exactly: q was pre-multiplied by scale * log2(e) (pi3b_gemm_qkv's q_scale)
  float* lse;        // optional fp32 [num_heads][num_seqs * q_len]: max + log2(sum) of the scaled log2-domain scores
  // optional fp32 [num_seqs][num_heads]: max over the sequence's keys of |k|^2 (kmax2_kernel). Enables the bounded-
  // reference softmax of the three-tile kernel: |q . k| <= |q| max|k| is an a-priori upper bound of every score of a row
  const float* kmax2;
  // work decomposition: one block per (query block of ROWS rows, sequence, head), query block fastest so that
  // neighbouring CTAs share one head's K/V in L2
  int blocks_per_seq;
};

struct Work {
  int seq, head, q0, q_valid;
};
struct FalseTag { static constexpr bool value = false; };
struct TrueTag { static constexpr bool value = true; };
template <int ROWS>
__device__ __forceinline__ Work decode_work(const AttnParams& p) {
  Work w;
  int item = blockIdx.x;
  const int qb = item % p.blocks_per_seq;
  item /= p.blocks_per_seq;
  w.seq = item % p.num_seqs;
  w.head = item / p.num_seqs;
  w.q0 = qb * ROWS;
  w.q_valid = min(ROWS, p.q_len - w.q0);
  return w;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// packed fp32x2 helpers (FFMA2 / FADD2 on sm_100): two lanes of fp32 per 64-bit register pair
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two exp2 at once on the FMA pipe: clamp (ALU pipe), round-to-integer via the 1.5*2^23 trick, minimax polynomial of 2^f
// on [-0.5, 0.5] (degree 2: max rel. error 1.73e-3 = 2^-9.2; degree 3: 7.5e-5), exponent spliced in with one integer
// shift-add per element. 7 (8) packed/scalar instructions for two elements next to 2 MUFU + nothing for the ex2 path.
// CLAMP = false: the caller guarantees x >= -126 (zero-reference steps: every score is within [-80, 80]), the two FMNMX
// per pair are dropped.
template <bool CLAMP = true>
__device__ __forceinline__ void poly_exp2_x2(uint64_t x, float& p0, float& p1) {
  uint64_t xc = x;
  if (CLAMP) {
    float x0, x1;
    unpack2(x, x0, x1);
    xc = pack2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  }
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);
  const uint64_t t = fadd2(xc, magic);
  const uint64_t r = fadd2(t, pack2(-12582912.0f, -12582912.0f));
  const uint64_t f = ffma2(r, pack2(-1.0f, -1.0f), xc);
#if PI3B_POLY_DEG == 2
  uint64_t q = ffma2(pack2(0.2384289323f, 0.2384289323f), f, pack2(0.7034480056f, 0.7034480056f));
  q = ffma2(q, f, pack2(1.0004431424f, 1.0004431424f));
#else
  uint64_t q = ffma2(pack2(0.0551716686f, 0.0551716686f), f, pack2(0.2426111222f, 0.2426111222f));
  q = ffma2(q, f, pack2(0.6932609856f, 0.6932609856f));
  q = ffma2(q, f, pack2(0.9999280735f, 0.9999280735f));
#endif
  float q0, q1, t0, t1;
  unpack2(q, q0, q1);
  unpack2(t, t0, t1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// ---- pieces shared by both kernels ----------------------------------------------------------------------------------

// exponentials of NCOL scores held in registers: a[c] (fp32 bit patterns) -> packed bf16 P in out[0 .. NCOL/2), row sum
// accumulated in the packed pair `lsum`. x = score * scale_log2 - m_used.
// RAW = true: the scores are already in their final units (log2 domain, reference 0): no scale-and-subtract at all.
template <int NCOL, bool RAW = false, bool CLAMP = true, int PAIRS = PI3B_POLY_PAIRS>
__device__ __forceinline__ void exp_block(const uint32_t* a, uint32_t* out, uint64_t sc2, uint64_t negm2, uint64_t& lsum) {
#pragma unroll
  for (int c = 0; c < NCOL; c += 2) {
    const uint64_t xin = pack2(__uint_as_float(a[c]), __uint_as_float(a[c + 1]));
    const uint64_t x = RAW ? xin : ffma2(xin, sc2, negm2);
    float p0, p1;
    // pairs are spread evenly over the row so that ptxas can interleave the FMA-pipe chains with the MUFU stream
    const bool poly = (((c >> 1) % 8) * PAIRS) % 8 < PAIRS;
    if (poly) {
      poly_exp2_x2<CLAMP>(x, p0, p1);
    } else {
      float x0, x1;
      unpack2(x, x0, x1);
      p0 = fast_exp2(x0);
      p1 = fast_exp2(x1);
    }
    lsum = fadd2(lsum, pack2(p0, p1));
    out[c >> 1] = pack_bf16x2(p0, p1);   // out may alias a: element c >> 1 <= c has already been consumed
  }
}

template <int NCOL>
__device__ __forceinline__ float row_max(const uint32_t* a) {
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < NCOL; c += 8) {
    mx0 = max3(mx0, __uint_as_float(a[c]), __uint_as_float(a[c + 1]));
    mx1 = max3(mx1, __uint_as_float(a[c + 2]), __uint_as_float(a[c + 3]));
    mx2 = max3(mx2, __uint_as_float(a[c + 4]), __uint_as_float(a[c + 5]));
    mx3 = max3(mx3, __uint_as_float(a[c + 6]), __uint_as_float(a[c + 7]));
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// key/value columns past the end of the sequence -> -inf (col0 = index of a[0] within the 64-column step)
template <int NCOL>
__device__ __forceinline__ void mask_tail(uint32_t* a, int col0, int kv_valid) {
#pragma unroll
  for (int c = 0; c < NCOL; ++c)
    if (col0 + c >= kv_valid) a[c] = 0xff800000u;
}

// the row max grew by more than 2^8 somewhere in this warp: rescale the running sum and O (64 fp32 TMEM columns)
__device__ __forceinline__ void rescale_row(bool grow, float m_tile, float& m_used, uint64_t& lsum, uint32_t tm_o) {
  float alpha = 1.0f;
  if (grow) {
    alpha = fast_exp2(m_used - m_tile);
    m_used = m_tile;
  }
  float la, lb;
  unpack2(lsum, la, lb);
  lsum = pack2(la * alpha, lb * alpha);
  uint32_t o[32];
#pragma unroll
  for (int hcol = 0; hcol < 64; hcol += 32) {
    tmem_ld_x32(tm_o + hcol, o);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
    tmem_st_x32(tm_o + hcol, o);
  }
  tmem_st_wait();
}

// epilogue of one query row: O / l -> bf16 -> HBM, optional log-sum-exp
__device__ __forceinline__ void store_row(const AttnParams& p, uint32_t tm_o, uint64_t lsum, float m_used, int head,
                                          int64_t grow, bool row_ok) {
  float la, lb;
  unpack2(lsum, la, lb);
  const float l = la + lb;
  const float inv_l = 1.0f / l;
  __nv_bfloat16* orow = p.out + grow * p.ldo + head * 64;
#pragma unroll
  for (int hcol = 0; hcol < 64; hcol += 32) {
    uint32_t o[32];
    tmem_ld_x32(tm_o + hcol, o);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + hcol + c) = v;
      }
    }
  }
  if (p.lse != nullptr && row_ok) p.lse[int64_t(head) * p.num_seqs * p.q_len + grow] = m_used + log2f(l);
}

// warp 0 of both kernels: Q tiles once, then the K/V ring (whole warp walks the loop, one elected lane issues)
template <int STAGES>
__device__ __forceinline__ void tma_producer(uint8_t* smem_q, uint8_t* smem_k, uint8_t* smem_v, uint64_t* q_full,
                                             uint64_t* k_full, uint64_t* v_full, uint64_t* kv_empty,
                                             const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                             const AttnParams& p, const Work& w, int num_t, int num_kv) {
  const int row_base = w.seq * p.q_len;
  const int kv_base = w.seq * p.kv_len;
  if (elect_one()) {
    mbar_expect_tx(q_full, num_t * kTileBytes);
    for (int t = 0; t < num_t; ++t)
      tma_load_2d(smem_q + t * kTileBytes, tmQ, q_full, p.q_col0 + w.head * 64, row_base + w.q0 + t * 128, kEvictFirst);
  }
  __syncwarp();
  for (int j = 0; j < num_kv; ++j) {
    const int s = j % STAGES;
    const uint32_t ph = (j / STAGES) & 1;
    mbar_wait(&kv_empty[s], ph ^ 1);
    if (elect_one()) {
      mbar_expect_tx(&k_full[s], kKvBytes);
      tma_load_2d(smem_k + s * kKvBytes, tmK, &k_full[s], p.k_col0 + w.head * 64, kv_base + j * 64, kEvictLast);
      mbar_expect_tx(&v_full[s], kKvBytes);
      tma_load_2d(smem_v + s * kKvBytes, tmV, &v_full[s], p.v_col0 + w.head * 64, kv_base + j * 64, kEvictLast);
    }
    __syncwarp();
  }
}

// (dropping this fence after tcgen05.wait::ld / wait::st measured no gain: profiles/r02_attn_poly_ab.log, "lean")
__device__ __forceinline__ void softmax_release_fence() { tc_fence_before(); }

// =====================================================================================================================
// Bounded-reference softmax (three-tile kernel, long sequences). Online softmax spends, per 64-key step and row, a
// 33-instruction FMNMX3 chain, the lazy-rescale vote and ~250 clocks of serial latency on tracking the running maximum.
// Softmax is invariant to the reference, so ANY upper bound of the row's scores that is not absurdly loose will do:
// Cauchy-Schwarz gives m_ref = |q| * max_j |k_j| * scale (exact in fp32 up to rounding, margin added). max_j |k_j|^2 per
// (sequence, head) comes from one streaming pass over K (kmax2_kernel, 206 MB at 100 400 keys, ~40 us); |q|^2 from the
// row's 64 values already in shared memory. Safety: the first step still runs the online code and yields m_lo <= true
// maximum; the fixed reference is only adopted when m_ref - m_lo <= 100 for every row of the warp (then the dominant
// terms are >= 2^-100, far from fp32 / bf16 underflow); otherwise the warp simply stays on the online path.
// =====================================================================================================================
__global__ void __launch_bounds__(256) kmax2_kernel(const __nv_bfloat16* __restrict__ k, int64_t ldk, int k_col0,
                                                    int kv_len, int num_heads, float* __restrict__ kmax2) {
  __shared__ int smax[64];
  const int seq = blockIdx.y;
  if (threadIdx.x < 64) smax[threadIdx.x] = 0;
  __syncthreads();
  const int sub = threadIdx.x & 7;                       // 8 lanes x 16 bytes = the 64 channels of one (row, head)
  const int64_t units = int64_t(kv_len) * num_heads;
  // block-uniform trip count (the shuffles below need every lane of the warp): out-of-range groups contribute 0
  for (int64_t u0 = int64_t(blockIdx.x) * 32; u0 < units; u0 += int64_t(gridDim.x) * 32) {
    const int64_t u = u0 + (threadIdx.x >> 3);
    const bool in_range = u < units;
    const int row = in_range ? int(u / num_heads) : 0, head = in_range ? int(u % num_heads) : 0;
    float q = 0.f;
    if (in_range) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(k + (int64_t(seq) * kv_len + row) * ldk + k_col0 + head * 64 + sub * 8));
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float lo = __uint_as_float(w4[i] << 16), hi = __uint_as_float(w4[i] & 0xffff0000u);
        q = fmaf(lo, lo, fmaf(hi, hi, q));
      }
    }
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 4);
    if (sub == 0 && in_range) atomicMax(&smax[head], __float_as_int(q));   // non-negative floats order like their bit patterns
  }
  __syncthreads();
  if (threadIdx.x < num_heads) atomicMax(reinterpret_cast<int*>(kmax2) + seq * num_heads + threadIdx.x, smax[threadIdx.x]);
}

// |q_row|^2 of this thread's query row from the SWIZZLE_128B Q tile in shared memory (128 bytes per row; the order of
// the 16-byte chunks does not matter for a sum of squares; lanes start at different chunks to spread the banks)
__device__ __forceinline__ float q_row_sqnorm(uint32_t q_tile_saddr, int row) {
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t a, b, cc, d;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(cc), "=r"(d) : "r"(q_tile_saddr + row * 128 + (((c + row) & 7) << 4)));
    const uint32_t w4[4] = {a, b, cc, d};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo = __uint_as_float(w4[i] << 16), hi = __uint_as_float(w4[i] & 0xffff0000u);
      q = fmaf(lo, lo, fmaf(hi, hi, q));
    }
  }
  return q;
}

// =====================================================================================================================
// Three-tile kernel (long sequences)
// =====================================================================================================================
constexpr int kA3Threads = 512;
constexpr int kA3Tiles = 3;
constexpr int kA3Stages = 8;
struct Attn3Smem {
  static constexpr int Q_OFF = 0;                                   // 3 tiles of 128 rows
  static constexpr int K_OFF = kA3Tiles * kTileBytes;
  static constexpr int V_OFF = K_OFF + kA3Stages * kKvBytes;
  static constexpr int BAR_OFF = V_OFF + kA3Stages * kKvBytes;
  static constexpr int NUM_BARS = 1 + 3 * kA3Stages + 4 * kA3Tiles;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
};
constexpr uint32_t kT3S = 0;     // S tile t at t*64        (64 fp32 columns)
constexpr uint32_t kT3O = 192;   // O tile t at 192 + t*64  (64 fp32 columns)
constexpr uint32_t kT3P = 384;   // P tile t at 384 + t*32  (64 bf16 = 32 packed columns)

__global__ void __launch_bounds__(kA3Threads, 1)
attn_fwd_kernel_t3(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Attn3Smem::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + kA3Stages;
  uint64_t* kv_empty = v_full + kA3Stages;
  uint64_t* s_full = kv_empty + kA3Stages;
  uint64_t* s_free = s_full + kA3Tiles;
  uint64_t* p_full = s_free + kA3Tiles;
  uint64_t* pv_done = p_full + kA3Tiles;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + kA3Tiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const Work w = decode_work<128 * kA3Tiles>(p);
  const int num_t = (w.q_valid + 127) / 128;
  const int num_kv = (p.kv_len + 63) / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kA3Stages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], num_t);   // one tcgen05.commit per tile issuer
    }
    for (int t = 0; t < kA3Tiles; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);         // one arrival per softmax warp
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    reg_dealloc<48>();
    if (warp == 0) {
      tma_producer<kA3Stages>(smem + Attn3Smem::Q_OFF, smem + Attn3Smem::K_OFF, smem + Attn3Smem::V_OFF, q_full, k_full,
                              v_full, kv_empty, &tmQ, &tmK, &tmV, p, w, num_t, num_kv);
    } else if (warp - 1 < num_t) {
      // ===== one MMA issuer warp per Q tile (warps 1..3): the tiles never wait for each other, so their softmax
      // warps are free to drift out of phase and interleave their exponential phases on the shared MUFU pipe =====
      const int t = warp - 1;
      constexpr uint32_t idesc_qk = make_idesc(1, 128, 64, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(1, 128, 64, 0, 1);  // B (=V) is MN-major
      const uint64_t q_desc = make_smem_desc_sw128(smem_u32(smem + Attn3Smem::Q_OFF + t * kTileBytes), 16, 1024);
      const uint64_t k_desc = make_smem_desc_sw128(smem_u32(smem + Attn3Smem::K_OFF), 16, 1024);
      const uint64_t v_desc = make_smem_desc_sw128(smem_u32(smem + Attn3Smem::V_OFF), 1024, 1024);
      const uint32_t d_s = tmem_base + kT3S + t * 64;
      const uint32_t d_o = tmem_base + kT3O + t * 64;
      const uint32_t a_p = tmem_base + kT3P + t * 32;

      auto issue_qk = [&](int s) {  // S_t = Q_t K(stage s)^T   (M128 N64 K64)
        const uint64_t b0 = k_desc + uint64_t((s * kKvBytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d_s, q_desc + 2 * k, b0 + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0);

      for (int j = 0; j < num_kv; ++j) {
        const int s = j % kA3Stages;
        const uint32_t ph = (j / kA3Stages) & 1;
        if (j + 1 < num_kv) {
          const int sn = (j + 1) % kA3Stages;
          mbar_wait(&k_full[sn], ((j + 1) / kA3Stages) & 1);
          mbar_wait(&s_free[t], j & 1);   // the softmax warps hold S_t(j) in registers: the buffer can be refilled
          tc_fence_after();
          issue_qk(sn);
        }
        mbar_wait(&v_full[s], ph);
        mbar_wait(&p_full[t], j & 1);     // P_t(j) is in TMEM
        tc_fence_after();
        const uint64_t b0 = v_desc + uint64_t((s * kKvBytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(d_o, a_p + k * 8, b0 + uint64_t(k * (2048 >> 4)), idesc_pv, (j | k) != 0);
          umma_commit(&pv_done[t]);
          umma_commit(&kv_empty[s]);   // this tile's QK(j) and PV(j) have consumed stage s (count = num_t)
        }
        __syncwarp();
      }
    }
  } else {
    // ===== softmax groups: one thread per query row, 64 columns per step =====
    reg_alloc<152>();
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    if (t < num_t) {
      const uint32_t lane_sel = uint32_t(quarter * 32) << 16;
      const uint32_t tm_s = tmem_base + lane_sel + kT3S + t * 64;
      const uint32_t tm_o = tmem_base + lane_sel + kT3O + t * 64;
      const uint32_t tm_p = tmem_base + lane_sel + kT3P + t * 32;
      const int row_in_cta = t * 128 + quarter * 32 + lane;
      const float sc = p.scale_log2;
      float m_used = -INFINITY;  // running max (scaled, log2 domain) the stored exponents are relative to
      const uint64_t sc2 = pack2(sc, sc);
      uint64_t lsum = pack2(0.f, 0.f);
      // A try_wait costs ~240 clocks even when the barrier has long completed (clock64 trace, profiles/
      // r01_attn_clock_trace.log). Both per-step barriers are therefore polled EARLY, while there is math to issue, and
      // only waited on if the early poll failed.
      bool s_ready = false;
      // One softmax step. MASK = true only for the last step of the sequence (kv_len % 64 != 0): inside the loop ptxas
      // if-converted the tail mask into 64 ISETP + 64 SEL executed on EVERY step (27 % of the loop's issue slots in the
      // round-1 kernel), so the masked variant is a separate copy of the body that runs once.
      auto step = [&](const int j, auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        if (!s_ready) mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t sr[64];
        tmem_ld_x32(tm_s, sr);
        tmem_ld_x32(tm_s + 32, sr + 32);
        tmem_ld_wait();
        // the scores are in registers: hand the S buffer back so QK(j+1) runs under this step's exponentials
        softmax_release_fence();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        // early poll: PV(j-1) was issued a whole step ago, its barrier is normally complete by now
        bool pv_ready = (j == 0) || mbar_try_wait(&pv_done[t], (j - 1) & 1);

        if (MASK) mask_tail<64>(sr, 0, p.kv_len - j * 64);
        const float m_tile = row_max<64>(sr) * sc;

        if (j == 0) {
          m_used = m_tile;
        } else {
          const bool grow = m_tile > m_used + 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            // O_t must be quiescent: PV(j-1) has finished
            if (!pv_ready) mbar_wait(&pv_done[t], (j - 1) & 1);
            pv_ready = true;
            tc_fence_after();
            rescale_row(grow, m_tile, m_used, lsum, tm_o);
          }
        }

        const uint64_t negm2 = pack2(-m_used, -m_used);
        exp_block<64>(sr, sr, sc2, negm2, lsum);
        // early poll of the next S tile: QK(j+1) was issued when s_free(j) arrived, ~a whole math phase ago
        s_ready = !MASK && mbar_try_wait(&s_full[t], (j + 1) & 1);
        // P_t was last read by PV(j-1): that MMA must have completed before the buffer is overwritten
        if (!pv_ready) mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
        tmem_st_x32(tm_p, sr);
        tmem_st_wait();
        softmax_release_fence();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      };
      // The same step with a FIXED reference m_used (bounded-reference softmax, see above): no row maximum, no rescale vote.
      auto step_fixed = [&](const int j, const uint64_t negm2, auto mask_tag, auto raw_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
        constexpr bool RAW = decltype(raw_tag)::value;
        if (!s_ready) mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t sr[64];
        tmem_ld_x32(tm_s, sr);
        tmem_ld_x32(tm_s + 32, sr + 32);
        tmem_ld_wait();
        softmax_release_fence();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        if (MASK) mask_tail<64>(sr, 0, p.kv_len - j * 64);
        // (moving the two polls into the middle of the exponentials measured no gain: profiles/r02_attn_poly_ab.log "mid2")
        const bool pv_ready = mbar_try_wait(&pv_done[t], (j - 1) & 1);
        // masked columns are -inf: the last step clamps. Without clamp and scale-and-subtract the polynomial costs 7
        // instructions per pair: the zero-reference steps take half of the pairs (PAIRS_ZERO = 4 of 8) on the FMA pipe
        exp_block<64, RAW, !(RAW && !MASK), RAW ? PI3B_POLY_PAIRS_ZERO : PI3B_POLY_PAIRS>(sr, sr, sc2, negm2, lsum);
        s_ready = !MASK && mbar_try_wait(&s_full[t], (j + 1) & 1);
        if (!pv_ready) mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
        tmem_st_x32(tm_p, sr);
        tmem_st_wait();
        softmax_release_fence();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      };

      // loop bounds are compared against p.kv_len (constant bank) so that no register is held for the trip count
      int j = 0;
      bool fixed = false, zero = false;
      if (p.kmax2 != nullptr && p.kv_len > 64) {
        // a-priori bound of this row's scores (log2 domain), with a margin for the fp32 rounding of norms and products
        mbar_wait(q_full, 0);
        const float q2 = q_row_sqnorm(smem_u32(smem + Attn3Smem::Q_OFF + t * kTileBytes), quarter * 32 + lane);
        const float m_ref = sqrtf(q2 * __ldg(p.kmax2 + w.seq * p.num_heads + w.head)) * sc * 1.001f + 0.0625f;
        step(0, FalseTag{});                        // online first step: m_used = maximum of the first 64 scores <= true maximum
        j = 1;
        fixed = __all_sync(0xffffffffu, m_ref - m_used <= 100.0f);
        // Unit scale (the caller folded scale * log2(e) into q: p.scale_log2 == 1) and every score of the warp's rows
        // within [-80, 80] of zero: zero itself is a valid reference — exp2 is applied to the accumulator as it comes
        // out of TMEM, the scale-and-subtract FFMA2 per pair of scores disappears (32 of ~290 instructions per step and row).
        // Range: P <= 2^80, row sum <= 2^80 x 100 400 keys = 2^97, O <= 2^97 max|v| — overflow needs |v| > 2^30, and
        // the terms that matter (>= 2^-80 by m_used >= -80) are far above the bf16 / fp32 denormal range.
        zero = p.unit_scale && __all_sync(0xffffffffu, m_ref <= 80.0f && m_used >= -80.0f);
        if (fixed || zero) {                        // move the reference (up to the bound, or to zero): O(0) and the row sum are rescaled once
          mbar_wait(&pv_done[t], 0);
          tc_fence_after();
          rescale_row(true, zero ? 0.0f : m_ref, m_used, lsum, tm_o);
        }
      }
      if (zero) {
        for (; (j + 1) * 64 < p.kv_len; ++j) step_fixed(j, 0ull, FalseTag{}, TrueTag{});
        step_fixed(j, 0ull, TrueTag{}, TrueTag{});
      } else if (fixed) {
        const uint64_t negm2 = pack2(-m_used, -m_used);
        for (; (j + 1) * 64 < p.kv_len; ++j) step_fixed(j, negm2, FalseTag{}, FalseTag{});
        step_fixed(j, negm2, TrueTag{}, FalseTag{});
      } else {
        for (; (j + 1) * 64 < p.kv_len; ++j) step(j, FalseTag{});
        step(j, TrueTag{});   // last step: columns past kv_len (if any) are masked to -inf
      }

      mbar_wait(&pv_done[t], j & 1);
      tc_fence_after();
      store_row(p, tm_o, lsum, m_used, w.head, int64_t(w.seq) * p.q_len + w.q0 + row_in_cta, row_in_cta < w.q_valid);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// Four-tile kernel (short sequences)
// =====================================================================================================================
constexpr int kA4Tiles = 4;
constexpr int kA4Threads = 32 * (4 + 4 * kA4Tiles);   // warp 0 TMA, warps 1-2 issuers (2 tiles each), warps 4-19 softmax
constexpr int kA4Stages = 8;
struct Attn4Smem {
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = kA4Tiles * kTileBytes;
  static constexpr int V_OFF = K_OFF + kA4Stages * kKvBytes;
  static constexpr int BAR_OFF = V_OFF + kA4Stages * kKvBytes;
  static constexpr int NUM_BARS = 1 + 3 * kA4Stages + 3 * kA4Tiles;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
};
constexpr uint32_t kT4S = 0;     // S tile t at t*64 (64 fp32 columns); P tile t = its first 32 columns (64 packed bf16)
constexpr uint32_t kT4O = 256;   // O tile t at 256 + t*64

__global__ void __launch_bounds__(kA4Threads, 1)
attn_fwd_kernel_t4(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Attn4Smem::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + kA4Stages;
  uint64_t* kv_empty = v_full + kA4Stages;
  uint64_t* s_full = kv_empty + kA4Stages;
  uint64_t* p_full = s_full + kA4Tiles;
  uint64_t* pv_done = p_full + kA4Tiles;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + kA4Tiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const Work w = decode_work<128 * kA4Tiles>(p);
  const int num_t = (w.q_valid + 127) / 128;
  const int num_kv = (p.kv_len + 63) / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kA4Stages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], num_t);   // one tcgen05.commit per tile
    }
    for (int t = 0; t < kA4Tiles; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    tma_producer<kA4Stages>(smem + Attn4Smem::Q_OFF, smem + Attn4Smem::K_OFF, smem + Attn4Smem::V_OFF, q_full, k_full,
                            v_full, kv_empty, &tmQ, &tmK, &tmV, p, w, num_t, num_kv);
  } else if (warp < 3) {
    // ===== MMA issuers: warp 1 serves tiles 0,1 and warp 2 tiles 2,3, each in a fixed order with per-tile waits =====
    constexpr uint32_t idesc_qk = make_idesc(1, 128, 64, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc(1, 128, 64, 0, 1);  // B (=V) is MN-major
    const uint64_t k_desc = make_smem_desc_sw128(smem_u32(smem + Attn4Smem::K_OFF), 16, 1024);
    const uint64_t v_desc = make_smem_desc_sw128(smem_u32(smem + Attn4Smem::V_OFF), 1024, 1024);
    const int t0 = (warp - 1) * 2;
    const int n_mine = min(2, num_t - t0);          // tiles this warp serves (<= 0: none)
    if (n_mine > 0) {
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int i = 0; i < n_mine; ++i) {
        const int t = t0 + i;
        const uint64_t q_desc = make_smem_desc_sw128(smem_u32(smem + Attn4Smem::Q_OFF + t * kTileBytes), 16, 1024);
        const uint32_t d_s = tmem_base + kT4S + t * 64;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d_s, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      }
      for (int j = 0; j < num_kv; ++j) {
        const int s = j % kA4Stages;
        const int sn = (j + 1) % kA4Stages;
        const bool more = j + 1 < num_kv;
        mbar_wait(&v_full[s], (j / kA4Stages) & 1);
        if (more) mbar_wait(&k_full[sn], ((j + 1) / kA4Stages) & 1);
        const uint64_t vb = v_desc + uint64_t((s * kKvBytes) >> 4);
        const uint64_t kb = k_desc + uint64_t((sn * kKvBytes) >> 4);
        for (int i = 0; i < n_mine; ++i) {
          const int t = t0 + i;
          mbar_wait(&p_full[t], j & 1);     // P_t(j) is in TMEM (over S_t(j), which is fully consumed)
          tc_fence_after();
          const uint64_t q_desc = make_smem_desc_sw128(smem_u32(smem + Attn4Smem::Q_OFF + t * kTileBytes), 16, 1024);
          const uint32_t d_s = tmem_base + kT4S + t * 64;   // S_t, and P_t in its first 32 columns
          const uint32_t d_o = tmem_base + kT4O + t * 64;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ts(d_o, d_s + k * 8, vb + uint64_t(k * (2048 >> 4)), idesc_pv, (j | k) != 0);
            umma_commit(&pv_done[t]);
            if (more) {
              // in-order tensor pipe: these writes to S_t follow PV(j)'s reads of P_t
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_ss(d_s, q_desc + 2 * k, kb + 2 * k, idesc_qk, k != 0);
              umma_commit(&s_full[t]);
            }
            umma_commit(&kv_empty[s]);     // this tile's QK(j) and PV(j) have consumed stage s
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // idle
  } else {
    // ===== softmax groups: one thread per query row, 64 columns per step =====
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    if (t < num_t) {
      const uint32_t lane_sel = uint32_t(quarter * 32) << 16;
      const uint32_t tm_s = tmem_base + lane_sel + kT4S + t * 64;
      const uint32_t tm_o = tmem_base + lane_sel + kT4O + t * 64;
      const int row_in_cta = t * 128 + quarter * 32 + lane;
      const float sc = p.scale_log2;
      float m_used = -INFINITY;
      const uint64_t sc2 = pack2(sc, sc);
      uint64_t lsum = pack2(0.f, 0.f);
      auto step = [&](const int j, auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;   // last step only (see the three-tile kernel)
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        const int kv_valid = p.kv_len - j * 64;
        // pass 1: row max (both halves in flight at once: only the max chains are live next to the 64 scores)
        float m_tile;
        {
          uint32_t a[64];
          tmem_ld_x32(tm_s, a);
          tmem_ld_x32(tm_s + 32, a + 32);
          tmem_ld_wait();
          if (MASK) mask_tail<64>(a, 0, kv_valid);
          m_tile = row_max<64>(a) * sc;
        }

        if (j == 0) {
          m_used = m_tile;
        } else {
          const bool grow = m_tile > m_used + 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            // O_t must be quiescent: PV(j-1) has finished (S(j) being ready already implies it: QK(j) was issued behind
            // PV(j-1) on the in-order pipe; the wait makes the dependency explicit)
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
            rescale_row(grow, m_tile, m_used, lsum, tm_o);
          }
        }

        const uint64_t negm2 = pack2(-m_used, -m_used);
        // pass 2: exponentials; S is re-read from TMEM, P (packed bf16) goes over the first 32 columns of S
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t a[32];
          tmem_ld_x32(tm_s + h * 32, a);
          tmem_ld_wait();
          if (MASK) mask_tail<32>(a, h * 32, kv_valid);
          exp_block<32>(a, a, sc2, negm2, lsum);
          // P half h goes to columns [16h, 16h+16); the other half's scores (columns 32..63) are still intact
          tmem_st_x16(tm_s + h * 16, a);
        }
        tmem_st_wait();
        softmax_release_fence();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      };
      int j = 0;
      for (; (j + 1) * 64 < p.kv_len; ++j) step(j, FalseTag{});
      step(j, TrueTag{});

      mbar_wait(&pv_done[t], j & 1);
      tc_fence_after();
      store_row(p, tm_o, lsum, m_used, w.head, int64_t(w.seq) * p.q_len + w.q0 + row_in_cta, row_in_cta < w.q_valid);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// Merge of attention results over disjoint key/value sets: out = sum_i w_i out_i, w_i = 2^(lse_i - max) / sum
// =====================================================================================================================
struct MergeParams {
  const __nv_bfloat16* part[4];
  const float* lse[4];
  int n_parts;
  int64_t ld_part;
  __nv_bfloat16* out;
  int64_t ldo;
  float* lse_out;     // optional
  int rows, num_heads;
};

// one thread = 8 channels of one (row, head); fully coalesced 16-byte accesses
__global__ void __launch_bounds__(256) attn_merge_kernel(const MergeParams p) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int per_row = p.num_heads * 8;
  if (idx >= int64_t(p.rows) * per_row) return;
  const int64_t row = idx / per_row;
  const int c8 = int(idx % per_row);
  const int head = c8 >> 3;
  float l[4];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    l[i] = i < p.n_parts ? __ldg(p.lse[i] + int64_t(head) * p.rows + row) : -INFINITY;
    mx = fmaxf(mx, l[i]);
  }
  float wsum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    l[i] = i < p.n_parts ? exp2f(l[i] - mx) : 0.f;
    wsum += l[i];
  }
  const float inv = 1.0f / wsum;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < p.n_parts) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.part[i] + row * p.ld_part + c8 * 8));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
      const float wgt = l[i] * inv;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        acc[2 * k] = fmaf(wgt, f.x, acc[2 * k]);
        acc[2 * k + 1] = fmaf(wgt, f.y, acc[2 * k + 1]);
      }
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]);
  o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]);
  o.w = pack_bf16x2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(p.out + row * p.ldo + c8 * 8) = o;
  if (p.lse_out != nullptr && (c8 & 7) == 0) p.lse_out[int64_t(head) * p.rows + row] = mx + log2f(wsum);
}

}  // namespace pi3b

using namespace pi3b;

constexpr int kLongSeqKeys = 3000;   // above: three-tile kernel with the bounded-reference softmax

static int attn_launch(void* stream, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                       int q_col0, int k_col0, int v_col0, void* out, int64_t ldo, float* lse, int num_seqs, int q_len,
                       int kv_len, int num_heads, float scale, void* workspace = nullptr, int64_t workspace_bytes = 0) {
  PI3B_REQUIRE(q && k && v && out, "pi3b_attn_fwd_hd64: null pointer");
  PI3B_REQUIRE(num_seqs > 0 && q_len > 0 && kv_len > 0 && num_heads > 0, "pi3b_attn_fwd_hd64: bad shape");
  PI3B_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "pi3b_attn_fwd_hd64: row pitches must be multiples of 8 elements");
  PI3B_REQUIRE(q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0, "pi3b_attn_fwd_hd64: column offsets %% 8");
  PI3B_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "pi3b_attn_fwd_hd64: out must be 16-byte aligned");
  const uint64_t rows = uint64_t(num_seqs) * q_len;
  const uint64_t kv_rows = uint64_t(num_seqs) * kv_len;
  PI3B_REQUIRE(rows < (1ull << 31) && kv_rows < (1ull << 31), "pi3b_attn_fwd_hd64: too many rows");
  // Dispatch by key/value length, from a same-box sweep at constant work (profiles/r02_attn_dispatch_threshold.json, ms for
  // three-tile + bounded reference / four-tile): S=1004 0.635 / 0.591, 2048 0.504 / 0.502, 3571 0.462 / 0.519, 4096 0.445 /
  // 0.503, 8192 0.545 / 0.636, 64300 17.2 / 19.0 — the four-tile kernel for the frame-local layers (T ~ 1k), the
  // three-tile kernel from ~3k keys up. PI3B_ATTN_TILES=3|4 forces one kernel (A/B runs and the both-kernels parity tests).
  static const int forced = []() {
    const char* e = getenv("PI3B_ATTN_TILES");
    const int r = e ? atoi(e) : 0;
    return (r == 3 || r == 4) ? r : 0;
  }();
  const int tiles = forced ? forced : (kv_len <= kLongSeqKeys ? 4 : 3);
  const int rows_per_cta = 128 * tiles;

  AttnParams p;
  p.num_seqs = num_seqs; p.q_len = q_len; p.kv_len = kv_len; p.num_heads = num_heads;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = (int)ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  // scale = ln 2 announces queries that already carry scale * log2(e) (the fused QKV epilogue's q_scale)
  static const bool zero_ok = []() { const char* e = getenv("PI3B_ATTN_ZEROREF"); return !(e && e[0] == '0'); }();
  p.unit_scale = zero_ok && fabsf(p.scale_log2 - 1.0f) < 1e-6f;
  if (p.unit_scale) p.scale_log2 = 1.0f;
  p.lse = lse;
  p.kmax2 = nullptr;

  const int dev = current_device();
  // (Handing the rows of the last, partly filled wave out as one-tile CTAs was measured and dropped: no gain at 28.3 waves,
  // profiles/r02_attn_poly_ab.log "@tail0", and the one-tile CTAs of all 16 heads run together, re-streaming every head's
  // K/V from DRAM: +0.4 GB per launch.)
  p.blocks_per_seq = (q_len + rows_per_cta - 1) / rows_per_cta;
  const int64_t items = int64_t(p.blocks_per_seq) * num_seqs * num_heads;
  PI3B_REQUIRE(items < (1ll << 31), "pi3b_attn_fwd_hd64: too many work items");

  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, q, 0, rows, uint64_t(q_col0) + 64ull * num_heads, uint64_t(ldq) * 2, 128, 64))) return rc;
  if ((rc = make_tmap_2d(&tmK, k, 0, kv_rows, uint64_t(k_col0) + 64ull * num_heads, uint64_t(ldk) * 2, 64, 64))) return rc;
  if ((rc = make_tmap_2d(&tmV, v, 0, kv_rows, uint64_t(v_col0) + 64ull * num_heads, uint64_t(ldv) * 2, 64, 64))) return rc;

  // the opt-in shared-memory size is a per-device function attribute
  static DeviceOnce configured;
  if (configured.first_call_on(dev)) {
    if (cudaFuncSetAttribute(attn_fwd_kernel_t3, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Smem::TOTAL) != cudaSuccess)
      return check_launch("attn: cudaFuncSetAttribute (t3)");
    if (cudaFuncSetAttribute(attn_fwd_kernel_t4, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn4Smem::TOTAL) != cudaSuccess)
      return check_launch("attn: cudaFuncSetAttribute (t4)");
    configured.done(dev);
  }
  // bounded-reference softmax (three-tile kernel): needs max |k|^2 per (sequence, head) in the caller's workspace.
  // (Measured and not kept for the four-tile kernel: at S = 1004 the pre-pass costs more than the dropped row-maximum
  // pass saves, 0.617 vs 0.595 ms; at S = 100 400 it ties with the three-tile kernel — profiles/r02_attn_poly_ab.log "t4b".)
  static const bool bounded_ok = []() { const char* e = getenv("PI3B_ATTN_BOUNDED"); return !(e && e[0] == '0'); }();
  if (tiles == 3 && bounded_ok && workspace != nullptr && num_heads <= 64 && kv_len > 64 &&
      workspace_bytes >= int64_t(num_seqs) * num_heads * 4 && (reinterpret_cast<uintptr_t>(workspace) & 3) == 0) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(workspace, 0, size_t(num_seqs) * num_heads * 4, st) != cudaSuccess) return check_launch("attn: memset");
    const int64_t units = int64_t(kv_len) * num_heads;
    const unsigned gx = unsigned(std::min<int64_t>((units + 31) / 32, int64_t(num_sms_of(dev)) * 8));
    kmax2_kernel<<<dim3(gx, num_seqs), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(k), ldk, k_col0, kv_len, num_heads,
                                                     static_cast<float*>(workspace));
    if (int rc2 = check_launch("kmax2 launch")) return rc2;
    p.kmax2 = static_cast<const float*>(workspace);
  }
  if (tiles == 4)
    attn_fwd_kernel_t4<<<(unsigned)items, kA4Threads, Attn4Smem::TOTAL, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  else
    attn_fwd_kernel_t3<<<(unsigned)items, kA3Threads, Attn3Smem::TOTAL, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  return check_launch("attn launch");
}

extern "C" int pi3b_attn_fwd_hd64(void* stream, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                                  int64_t ldv, int q_col0, int k_col0, int v_col0, void* out, int64_t ldo,
                                  int num_seqs, int q_len, int kv_len, int num_heads, float scale) {
  return attn_launch(stream, q, k, v, ldq, ldk, ldv, q_col0, k_col0, v_col0, out, ldo, nullptr, num_seqs, q_len, kv_len,
                     num_heads, scale);
}

extern "C" int pi3b_attn_fwd_hd64_lse(void* stream, const void* q, const void* k, const void* v, int64_t ldq,
                                      int64_t ldk, int64_t ldv, int q_col0, int k_col0, int v_col0, void* out,
                                      int64_t ldo, float* lse, int num_seqs, int q_len, int kv_len, int num_heads,
                                      float scale) {
  PI3B_REQUIRE(lse != nullptr, "pi3b_attn_fwd_hd64_lse: lse is null");
  return attn_launch(stream, q, k, v, ldq, ldk, ldv, q_col0, k_col0, v_col0, out, ldo, lse, num_seqs, q_len, kv_len,
                     num_heads, scale);
}

extern "C" int pi3b_attn_merge(void* stream, int n_parts, const void* const* parts, const float* const* lses,
                               int64_t ld_part, void* out, int64_t ldo, float* lse_out, int rows, int num_heads) {
  PI3B_REQUIRE(n_parts >= 1 && n_parts <= 4, "pi3b_attn_merge: 1..4 partial results");
  PI3B_REQUIRE(parts && lses && out && rows > 0 && num_heads > 0, "pi3b_attn_merge: bad argument");
  PI3B_REQUIRE(ld_part % 8 == 0 && ldo % 8 == 0, "pi3b_attn_merge: row pitches must be multiples of 8 elements");
  MergeParams p;
  for (int i = 0; i < 4; ++i) {
    p.part[i] = i < n_parts ? static_cast<const __nv_bfloat16*>(parts[i]) : nullptr;
    p.lse[i] = i < n_parts ? lses[i] : nullptr;
    if (i < n_parts) {
      PI3B_REQUIRE(p.part[i] && p.lse[i], "pi3b_attn_merge: null partial result");
      PI3B_REQUIRE((reinterpret_cast<uintptr_t>(p.part[i]) & 15) == 0, "pi3b_attn_merge: parts must be 16-byte aligned");
    }
  }
  PI3B_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "pi3b_attn_merge: out must be 16-byte aligned");
  p.n_parts = n_parts;
  p.ld_part = ld_part;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.lse_out = lse_out;
  p.rows = rows;
  p.num_heads = num_heads;
  const int64_t threads = int64_t(rows) * num_heads * 8;
  attn_merge_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("attn merge launch");
}

extern "C" int pi3b_attn_fwd_hd64_ws(void* stream, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                                     int64_t ldv, int q_col0, int k_col0, int v_col0, void* out, int64_t ldo, float* lse,
                                     int num_seqs, int q_len, int kv_len, int num_heads, float scale, void* workspace,
                                     int64_t workspace_bytes) {
  return attn_launch(stream, q, k, v, ldq, ldk, ldv, q_col0, k_col0, v_col0, out, ldo, lse, num_seqs, q_len, kv_len,
                     num_heads, scale, workspace, workspace_bytes);
}
