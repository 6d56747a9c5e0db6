// pi3b_attn_fwd_hd64: softmax(Q K^T / 8) V for head_dim 64, bf16 in / bf16 out, no mask.
//
// Replaces F.scaled_dot_product_attention on the pi3 chunk forward (reference:
// pi3/models/layers/attention.py:102-104 and :336-338) for both regimes of Pi3.decode
// (pi3/models/pi3.py:156-166): many frame-local sequences (S = T ~ 1k) and one chunk-global
// sequence (S = N*T ~ 100k), which is ~75 % of the FLOPs of a chunk.
//
// Design (FA4-style, hand-written tcgen05), revision 5:
//   * one CTA = one (sequence, head, 256 query rows): two 128-row Q tiles ping-pong so that the
//     tensor pipe works on one tile while the other tile's softmax runs on the SIMT/MUFU pipes;
//   * warp 0 streams K/V tiles (128 rows x 64) with TMA into a 4-stage smem ring; warp 1 (one thread)
//     issues S = Q K^T (M128 N128 K64, both operands K-major in smem) and O += P V (A = P from TMEM,
//     B = V MN-major in smem); S, P and O all live in TMEM (2x128 + 2x64 + 2x64 = 512 columns);
//   * warps 4-7 / 8-11 are the two softmax groups: one thread per query row (tcgen05.ld 32x32b), so
//     row max / row sum need no shuffles; online softmax in the exp2 domain with lazy rescaling of O
//     (only when the row max grows by more than 2^8), P written back to TMEM as packed bf16; packed
//     fp32x2 FMA/ADD and 3-input max halve the issue slots next to the 128 MUFU.EX2 per row-step;
//   * at head_dim 64 the softmax side (MUFU: 16 ex2/clk/SM; TMEM read: 64 B/clk/SM) is the bound, so the
//     tensor pipe must never be the reason a softmax warp waits: the whole MMA warp walks the schedule with
//     uniform operands and one elected lane issues (no per-instruction waterfall loops), and a softmax
//     group releases its S buffer as soon as the scores are in registers (s_free), so QK(j+1) is computed
//     while the group is still busy with the exponentials of tile j;
//   * ragged tails: KV columns past the sequence end are masked to -inf, Q rows past it are not stored.
//
// Measured anatomy of one softmax step (clock64 trace of CTA 0, -DPI3B_TRACE + scripts/trace_attn.py; 128 KV columns,
// both Q tiles, ~2850 clocks): barrier wait + tcgen05 fence for S ~250, tcgen05.ld of the row ~250, max + exponentials
// ~1650, wait for PV(j-1) ~230, tcgen05.st + fence + arrive ~200. The exponential phase is not MUFU-bound (its MUFU work
// is ~900 clocks): ptxas places each MUFU.EX2's consumer 5-6 instructions behind it, warps issue in order, and with two
// softmax warps per scheduler only ~1 MUFU per 14 clocks is issued. Variants measured on the same box and NOT kept
// (all within 800-870 TFLOP/s at S = 100400): 16 softmax warps with column-split rows (MUFU saturates, but every warp
// then sits in the ~1100-clock latency phase at the same time), one MMA issuer per Q tile with a half-step stagger,
// consumers lagging one block behind the exponentials (ptxas re-interleaves them, also across bar.warp.sync), 64-column
// steps with double / triple buffered S, intra-warp S prefetch, strict MUFU ping-pong via named barriers. Without any
// exponentials the kernel reaches 1000 TFLOP/s, so the remaining gap to the tensor roofline is synchronisation latency,
// not MUFU throughput or TMEM bandwidth. The two tiles do NOT run in lockstep: QK0(0) / QK1(0) are issued back to back,
// which leaves tile 1 about 1000 clocks (a third of a step) behind tile 0 for the whole kernel, and an extra start delay
// of 500 / 900 / 1400 clocks for tile 1 changes nothing (845-853 TFLOP/s, profiles/r01_attn_stagger_ab.log): the
// ~1650-clock exponential phase is the warp's own in-order stream (~550 instructions, each MUFU pair consumed ~6
// instructions later), not contention between the two groups. What would hide it is more resident softmax warps per
// scheduler (4 Q tiles x 64-column steps with P aliased onto S fit TMEM exactly: 4 x (64 + 64) columns) — next round.
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace pi3b {

constexpr int kAttnThreads = 384;
constexpr int kKvStages = 4;
#ifndef PI3B_POLY_EVERY
#define PI3B_POLY_EVERY 4
#endif
constexpr int kPolyEvery = PI3B_POLY_EVERY;  // 1 of every 2*kPolyEvery exponentials runs on the FMA pipe (A/B on B200: 0 -> 815, 2 -> 828, 3 -> 862, 4 -> 866 TFLOP/s)  // every kPolyEvery-th pair computes one of its two exp2 on the FMA pipe (0 = off)
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16

struct AttnSmem {
  static constexpr int Q_OFF = 0;                                // 2 tiles
  static constexpr int K_OFF = 2 * kTileBytes;                   // kKvStages tiles
  static constexpr int V_OFF = K_OFF + kKvStages * kTileBytes;   // kKvStages tiles
  static constexpr int BAR_OFF = V_OFF + kKvStages * kTileBytes;
  // barriers: q_full, k_full[S], v_full[S], kv_empty[S], s_full[2], s_free[2], p_full[2], pv_done[2]
  static constexpr int NUM_BARS = 1 + 3 * kKvStages + 8;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
};

struct AttnParams {
  int num_seqs, q_len, kv_len, num_heads;  // per sequence: q_len query rows attend to kv_len key/value rows
  int q_col0, k_col0, v_col0;  // element column of head 0 inside each operand's rows
  __nv_bfloat16* out;
  int ldo;
  float scale_log2;  // softmax scale * log2(e)
};

#ifdef PI3B_TRACE
// debug-only timeline: CTA 0 writes clock64 stamps [role][iteration][event] into a global buffer
__device__ unsigned long long* g_trace = nullptr;
#define TRACE(role, it, ev)                                                                           \
  do {                                                                                                \
    if (blockIdx.x == 0 && lane == 0 && (it) < 64 && g_trace) g_trace[((role)*64 + (it)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define TRACE(role, it, ev) do {} while (0)
#endif

// TMEM column map
constexpr uint32_t kTmS = 0;     // S tile t at kTmS + t*128   (128 fp32 columns)
constexpr uint32_t kTmO = 256;   // O tile t at kTmO + t*64    (64 fp32 columns)
constexpr uint32_t kTmP = 384;   // P tile t at kTmP + t*64    (128 bf16 = 64 packed columns)

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, rel. error < 1e-4, well below the
// bf16 rounding of P): used for a fixed share of the elements to take load off the MUFU pipe.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: rounds x to the nearest integer n in the low mantissa bits
  const float f = x - (t - 12582912.0f);       // f in [-0.5, 0.5]
  float p = fmaf(0.0555041086f, f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // p * 2^n
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

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + kKvStages;
  uint64_t* kv_empty = v_full + kKvStages;
  uint64_t* s_full = kv_empty + kKvStages;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_full = s_free + 2;
  uint64_t* pv_done = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item: query-pair fastest, then sequence, then head (neighbouring CTAs share one head's K/V in L2)
  const int num_qpairs = (p.q_len + 255) / 256;
  int item = blockIdx.x;
  const int qp = item % num_qpairs;
  item /= num_qpairs;
  const int seq = item % p.num_seqs;
  const int head = item / p.num_seqs;
  const int row_base = seq * p.q_len;            // first query / output row of this sequence
  const int kv_base = seq * p.kv_len;            // first key / value row of this sequence
  const int q0 = qp * 256;                       // first query row (within the sequence) of this CTA
  const int q_valid = min(256, p.q_len - q0);    // valid query rows in this CTA
  const int num_t = q_valid > 128 ? 2 : 1;       // second Q tile entirely past the end -> skip it
  const int num_kv = (p.kv_len + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);  // one arrival per softmax warp
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

  // softmax threads hold a whole 128-wide score row in registers; the other warpgroup donates its budget
  if (warp < 4) {
    reg_dealloc<48>();
    if (warp == 0) {
      // ===== TMA producer (whole warp runs the loop, one elected lane issues: keeps operands uniform) =====
      if (elect_one()) {
        mbar_expect_tx(q_full, num_t * kTileBytes);
        for (int t = 0; t < num_t; ++t)
          tma_load_2d(smem + AttnSmem::Q_OFF + t * kTileBytes, &tmQ, q_full, p.q_col0 + head * 64,
                      row_base + q0 + t * 128, kEvictFirst);
      }
      __syncwarp();
      for (int j = 0; j < num_kv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[s], kTileBytes);
          tma_load_2d(smem + AttnSmem::K_OFF + s * kTileBytes, &tmK, &k_full[s], p.k_col0 + head * 64,
                      kv_base + j * 128, kEvictLast);
          mbar_expect_tx(&v_full[s], kTileBytes);
          tma_load_2d(smem + AttnSmem::V_OFF + s * kTileBytes, &tmV, &v_full[s], p.v_col0 + head * 64,
                      kv_base + j * 128, kEvictLast);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===== MMA issuer: the whole warp walks the schedule (uniform control flow and operands, so the
      // descriptors live in uniform registers), one elected lane issues tcgen05.mma / commit =====
      constexpr uint32_t idesc_qk = make_idesc(1, 128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(1, 128, 64, 0, 1);  // B (=V) is MN-major
      // descriptor = base + (byte offset >> 4): the 14-bit start-address field never carries out
      const uint64_t q_desc = make_smem_desc_sw128(smem_u32(smem + AttnSmem::Q_OFF), 16, 1024);
      const uint64_t k_desc = make_smem_desc_sw128(smem_u32(smem + AttnSmem::K_OFF), 16, 1024);
      const uint64_t v_desc = make_smem_desc_sw128(smem_u32(smem + AttnSmem::V_OFF), 1024, 1024);

      auto issue_qk = [&](int t, int s) {  // S_t = Q_t K(stage s)^T
        const uint32_t d = tmem_base + kTmS + t * 128;
        const uint64_t a0 = q_desc + uint64_t((t * kTileBytes) >> 4);
        const uint64_t b0 = k_desc + uint64_t((s * kTileBytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d, a0 + 2 * k, b0 + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < num_t; ++t) issue_qk(t, 0);

      for (int j = 0; j < num_kv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        if (j + 1 < num_kv) {
          const int sn = (j + 1) % kKvStages;
          mbar_wait(&k_full[sn], ((j + 1) / kKvStages) & 1);
          for (int t = 0; t < num_t; ++t) {
            mbar_wait(&s_free[t], j & 1);  // softmax t holds S_t(j) in registers: the buffer can be refilled
            tc_fence_after();
            TRACE(0, j, t);          // ev 0/1: s_free[t] observed
            issue_qk(t, sn);
            TRACE(0, j, 2 + t);      // ev 2/3: QK(t, j+1) issued
          }
        }
        mbar_wait(&v_full[s], ph);
        for (int t = 0; t < num_t; ++t) {
          mbar_wait(&p_full[t], j & 1);  // P_t(j) is in TMEM
          tc_fence_after();
          TRACE(0, j, 4 + t);        // ev 4/5: p_full[t] observed
          const uint32_t d = tmem_base + kTmO + t * 64;
          const uint32_t a = tmem_base + kTmP + t * 64;
          const uint64_t b0 = v_desc + uint64_t((s * kTileBytes) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16_ts(d, a + k * 8, b0 + uint64_t(k * (2048 >> 4)), idesc_pv, (j | k) != 0);
            umma_commit(&pv_done[t]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&kv_empty[s]);  // QK(j) and PV(j) of both tiles have consumed stage s
        __syncwarp();
        TRACE(0, j, 6);              // ev 6: iteration issued
      }
    }
  } else {
    // ===== softmax groups: one thread per query row =====
    reg_alloc<224>();
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    if (t < num_t) {
      const uint32_t lane_sel = uint32_t(quarter * 32) << 16;
      const uint32_t tm_s = tmem_base + lane_sel + kTmS + t * 128;
      const uint32_t tm_o = tmem_base + lane_sel + kTmO + t * 64;
      const uint32_t tm_p = tmem_base + lane_sel + kTmP + t * 64;
      const int row_in_cta = t * 128 + quarter * 32 + lane;
      const float sc = p.scale_log2;
      float m_used = -INFINITY;  // running max (scaled, log2 domain) the stored exponents are relative to

      const uint64_t sc2 = pack2(sc, sc);
      uint64_t lsum = pack2(0.f, 0.f);
      // A try_wait costs ~240 clocks even when the barrier has long completed (measured with the clock64 trace:
      // 2 waits = 17 % of a softmax step). Both per-step barriers are therefore polled EARLY, while there is math
      // to issue, and only waited on if the early poll failed.
      bool s_ready = false;
      for (int j = 0; j < num_kv; ++j) {
        if (quarter == 0) TRACE(1 + t, j, 0);   // ev 0: step start
        if (!s_ready) mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        if (quarter == 0) TRACE(1 + t, j, 1);   // ev 1: S(j) ready
        uint32_t sr[128];
        tmem_ld_x32(tm_s, sr);
#ifdef PI3B_DBG_NOLD
#pragma unroll
        for (int c = 32; c < 128; ++c) sr[c] = sr[c & 31] + c;   // roofline experiment: 1/4 of the TMEM reads
#else
        tmem_ld_x32(tm_s + 32, sr + 32);
        tmem_ld_x32(tm_s + 64, sr + 64);
        tmem_ld_x32(tm_s + 96, sr + 96);
#endif
        tmem_ld_wait();
        // the scores are in registers: hand the S buffer back so QK(j+1) runs under this step's exponentials
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        if (quarter == 0) TRACE(1 + t, j, 2);   // ev 2: S in registers, s_free arrived
        // early poll: PV(j-1) was issued a whole step ago, its barrier is normally complete by now
        bool pv_ready = (j == 0) || mbar_try_wait(&pv_done[t], (j - 1) & 1);

        const int kv_valid = p.kv_len - j * 128;
        if (kv_valid < 128) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= kv_valid) sr[c] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
          mx0 = max3(mx0, __uint_as_float(sr[c]), __uint_as_float(sr[c + 1]));
          mx1 = max3(mx1, __uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3]));
          mx2 = max3(mx2, __uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5]));
          mx3 = max3(mx3, __uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7]));
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;

        if (j == 0) {
          m_used = m_tile;
        } else {
          const bool grow = m_tile > m_used + 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            // O_t must be quiescent: PV(j-1) has finished
            if (!pv_ready) mbar_wait(&pv_done[t], (j - 1) & 1);
            pv_ready = true;
            tc_fence_after();
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
        }

        const uint64_t negm2 = pack2(-m_used, -m_used);
#pragma unroll
        for (int c = 0; c < 128; c += 2) {
          const uint64_t x = ffma2(pack2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), sc2, negm2);
          float x0, x1;
          unpack2(x, x0, x1);
#ifdef PI3B_DBG_NOEXP
          const float p0 = x0;
#else
          const float p0 = (kPolyEvery > 0 && (c % (2 * (kPolyEvery > 0 ? kPolyEvery : 1))) == 0) ? poly_exp2(x0) : fast_exp2(x0);
#endif
#ifdef PI3B_DBG_NOEXP
          const float p1 = x1;                                   // roofline experiment: no MUFU work
#else
          const float p1 = fast_exp2(x1);
#endif
          lsum = fadd2(lsum, pack2(p0, p1));
          sr[c >> 1] = pack_bf16x2(p0, p1);
        }
        if (quarter == 0) TRACE(1 + t, j, 3);   // ev 3: exponentials done
        // early poll of the next S tile: QK(j+1) was issued when s_free(j) arrived, ~a whole math phase ago
        s_ready = (j + 1 < num_kv) && mbar_try_wait(&s_full[t], (j + 1) & 1);
        // P_t was last read by PV(j-1): that MMA must have completed before the buffer is overwritten
        if (!pv_ready) mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
        if (quarter == 0) TRACE(1 + t, j, 4);   // ev 4: PV(j-1) done
        tmem_st_x32(tm_p, sr);
        tmem_st_x32(tm_p + 32, sr + 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (quarter == 0) TRACE(1 + t, j, 5);   // ev 5: P stored, p_full arrived
      }

      // epilogue: O / l -> bf16 -> HBM
      mbar_wait(&pv_done[t], (num_kv - 1) & 1);
      tc_fence_after();
      float la, lb;
      unpack2(lsum, la, lb);
      const float inv_l = 1.0f / (la + lb);
      const bool row_ok = row_in_cta < q_valid;
      __nv_bfloat16* orow = p.out + size_t(row_base + q0 + row_in_cta) * p.ldo + head * 64;
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
// Revision 8 (default): THREE 128-row Q tiles per CTA and 64-column KV steps, so that three softmax warps share each
// scheduler instead of two. The per-step chain of one softmax warp is dominated by fixed latencies (barrier + tcgen05
// fences ~450, tcgen05.ld/st ~350, loop ~270 clocks) around an in-order exponential stream; with two resident warps per
// scheduler nothing hides them. S (3 x 64 fp32 columns), O (3 x 64) and P (3 x 32) fit TMEM as 480 of 512 columns,
// K/V tiles are 64 rows in an 8-stage ring. Same barrier protocol as revision 5, but ONE MMA ISSUER WARP PER TILE: with a
// single issuer doing grouped waits the three tiles ran in lockstep, their exponential phases coincided on the MUFU
// pipe and the kernel fell to 645 TFLOP/s; decoupled, the tiles drift apart and it reaches 880-890 (revision 5 on the
// same box: 855-868), profiles/r01_attn_rev8_ab.log. ptxas sizes each role's code to its setmaxnreg budget, and the
// budgets must add up to the CTA's launch allocation (threads x launch registers) or the last setmaxnreg.inc blocks
// forever. The four-tile layout (P in place over S) is revision 9 below; its first form (single-pass softmax spilling
// 2 KB/thread at 640 threads, issuer warps spinning on two tiles each) ran 14x slower.
// =====================================================================================================================
constexpr int kA3Threads = 512;
constexpr int kA3Tiles = 3;
constexpr int kA3Stages = 8;
constexpr int kA3KvBytes = 64 * 128;   // 64 rows x 64 bf16
struct Attn3Smem {
  static constexpr int Q_OFF = 0;                                   // 3 tiles of 128 rows
  static constexpr int K_OFF = kA3Tiles * kTileBytes;
  static constexpr int V_OFF = K_OFF + kA3Stages * kA3KvBytes;
  static constexpr int BAR_OFF = V_OFF + kA3Stages * kA3KvBytes;
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

  constexpr int kRows = 128 * kA3Tiles;
  const int num_qblocks = (p.q_len + kRows - 1) / kRows;
  int item = blockIdx.x;
  const int qb = item % num_qblocks;
  item /= num_qblocks;
  const int seq = item % p.num_seqs;
  const int head = item / p.num_seqs;
  const int row_base = seq * p.q_len;
  const int kv_base = seq * p.kv_len;
  const int q0 = qb * kRows;
  const int q_valid = min(kRows, p.q_len - q0);
  const int num_t = (q_valid + 127) / 128;
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
      mbar_init(&s_free[t], 4);
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
      // ===== TMA producer =====
      if (elect_one()) {
        mbar_expect_tx(q_full, num_t * kTileBytes);
        for (int t = 0; t < num_t; ++t)
          tma_load_2d(smem + Attn3Smem::Q_OFF + t * kTileBytes, &tmQ, q_full, p.q_col0 + head * 64,
                      row_base + q0 + t * 128, kEvictFirst);
      }
      __syncwarp();
      for (int j = 0; j < num_kv; ++j) {
        const int s = j % kA3Stages;
        const uint32_t ph = (j / kA3Stages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&k_full[s], kA3KvBytes);
          tma_load_2d(smem + Attn3Smem::K_OFF + s * kA3KvBytes, &tmK, &k_full[s], p.k_col0 + head * 64,
                      kv_base + j * 64, kEvictLast);
          mbar_expect_tx(&v_full[s], kA3KvBytes);
          tma_load_2d(smem + Attn3Smem::V_OFF + s * kA3KvBytes, &tmV, &v_full[s], p.v_col0 + head * 64,
                      kv_base + j * 64, kEvictLast);
        }
        __syncwarp();
      }
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
        const uint64_t b0 = k_desc + uint64_t((s * kA3KvBytes) >> 4);
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
          mbar_wait(&s_free[t], j & 1);
          tc_fence_after();
          issue_qk(sn);
        }
        mbar_wait(&v_full[s], ph);
        mbar_wait(&p_full[t], j & 1);
        tc_fence_after();
        const uint64_t b0 = v_desc + uint64_t((s * kA3KvBytes) >> 4);
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
      float m_used = -INFINITY;
      const uint64_t sc2 = pack2(sc, sc);
      uint64_t lsum = pack2(0.f, 0.f);
      bool s_ready = false;
      for (int j = 0; j < num_kv; ++j) {
        if (!s_ready) mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t sr[64];
        tmem_ld_x32(tm_s, sr);
        tmem_ld_x32(tm_s + 32, sr + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        bool pv_ready = (j == 0) || mbar_try_wait(&pv_done[t], (j - 1) & 1);

        const int kv_valid = p.kv_len - j * 64;
        if (kv_valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= kv_valid) sr[c] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          mx0 = max3(mx0, __uint_as_float(sr[c]), __uint_as_float(sr[c + 1]));
          mx1 = max3(mx1, __uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3]));
          mx2 = max3(mx2, __uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5]));
          mx3 = max3(mx3, __uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7]));
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;

        if (j == 0) {
          m_used = m_tile;
        } else {
          const bool grow = m_tile > m_used + 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            if (!pv_ready) mbar_wait(&pv_done[t], (j - 1) & 1);
            pv_ready = true;
            tc_fence_after();
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
        }

        const uint64_t negm2 = pack2(-m_used, -m_used);
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          const uint64_t x = ffma2(pack2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), sc2, negm2);
          float x0, x1;
          unpack2(x, x0, x1);
          const float p0 = (kPolyEvery > 0 && (c % (2 * (kPolyEvery > 0 ? kPolyEvery : 1))) == 0) ? poly_exp2(x0) : fast_exp2(x0);
          const float p1 = fast_exp2(x1);
          lsum = fadd2(lsum, pack2(p0, p1));
          sr[c >> 1] = pack_bf16x2(p0, p1);
        }
        s_ready = (j + 1 < num_kv) && mbar_try_wait(&s_full[t], (j + 1) & 1);
        if (!pv_ready) mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
        tmem_st_x32(tm_p, sr);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }

      mbar_wait(&pv_done[t], (num_kv - 1) & 1);
      tc_fence_after();
      float la, lb;
      unpack2(lsum, la, lb);
      const float inv_l = 1.0f / (la + lb);
      const bool row_ok = row_in_cta < q_valid;
      __nv_bfloat16* orow = p.out + size_t(row_base + q0 + row_in_cta) * p.ldo + head * 64;
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
// Revision 9 (default for kv_len <= 8192): FOUR 128-row Q tiles per CTA, 64-column KV steps, P written in place over
// the first 32 columns of its own S tile (S/P 4 x 64 + O 4 x 64 = 512 TMEM columns), four softmax warps per scheduler.
// Two issuer warps serve two tiles each, in a fixed order with per-tile blocking waits: PV(j) and then QK(j+1) back to
// back (the tensor pipe executes them in order, so the new scores overwrite P only after PV has read it); no s_free /
// per-step pv_done handshakes remain on the softmax side. The softmax step is two passes over 32-column halves (row max,
// then exponentials, S re-read from TMEM) so that it fits the 96 registers a 640-thread CTA leaves per thread.
// =====================================================================================================================
constexpr int kA4Tiles = 4;
constexpr int kA4Threads = 32 * (4 + 4 * kA4Tiles);   // warp 0 TMA, warps 1-2 issuers (2 tiles each), warps 4-19 softmax
constexpr int kA4Stages = 8;
struct Attn4Smem {
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = kA4Tiles * kTileBytes;
  static constexpr int V_OFF = K_OFF + kA4Stages * kA3KvBytes;
  static constexpr int BAR_OFF = V_OFF + kA4Stages * kA3KvBytes;
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

  constexpr int kRows = 128 * kA4Tiles;
  const int num_qblocks = (p.q_len + kRows - 1) / kRows;
  int item = blockIdx.x;
  const int qb = item % num_qblocks;
  item /= num_qblocks;
  const int seq = item % p.num_seqs;
  const int head = item / p.num_seqs;
  const int row_base = seq * p.q_len;
  const int kv_base = seq * p.kv_len;
  const int q0 = qb * kRows;
  const int q_valid = min(kRows, p.q_len - q0);
  const int num_t = (q_valid + 127) / 128;
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
      mbar_init(&kv_empty[i], num_t);   // one tcgen05.commit per tile issuer
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
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(q_full, num_t * kTileBytes);
      for (int t = 0; t < num_t; ++t)
        tma_load_2d(smem + Attn4Smem::Q_OFF + t * kTileBytes, &tmQ, q_full, p.q_col0 + head * 64,
                    row_base + q0 + t * 128, kEvictFirst);
    }
    __syncwarp();
    for (int j = 0; j < num_kv; ++j) {
      const int s = j % kA4Stages;
      const uint32_t ph = (j / kA4Stages) & 1;
      mbar_wait(&kv_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&k_full[s], kA3KvBytes);
        tma_load_2d(smem + Attn4Smem::K_OFF + s * kA3KvBytes, &tmK, &k_full[s], p.k_col0 + head * 64,
                    kv_base + j * 64, kEvictLast);
        mbar_expect_tx(&v_full[s], kA3KvBytes);
        tma_load_2d(smem + Attn4Smem::V_OFF + s * kA3KvBytes, &tmV, &v_full[s], p.v_col0 + head * 64,
                    kv_base + j * 64, kEvictLast);
      }
      __syncwarp();
    }
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
        const uint64_t vb = v_desc + uint64_t((s * kA3KvBytes) >> 4);
        const uint64_t kb = k_desc + uint64_t((sn * kA3KvBytes) >> 4);
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
      for (int j = 0; j < num_kv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        const int kv_valid = p.kv_len - j * 64;
        // pass 1: row max (both halves in flight at once: only the max chains are live next to the 64 scores)
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        {
          uint32_t a[64];
          tmem_ld_x32(tm_s, a);
          tmem_ld_x32(tm_s + 32, a + 32);
          tmem_ld_wait();
          if (kv_valid < 64) {
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (c >= kv_valid) a[c] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int c = 0; c < 64; c += 8) {
            mx0 = max3(mx0, __uint_as_float(a[c]), __uint_as_float(a[c + 1]));
            mx1 = max3(mx1, __uint_as_float(a[c + 2]), __uint_as_float(a[c + 3]));
            mx2 = max3(mx2, __uint_as_float(a[c + 4]), __uint_as_float(a[c + 5]));
            mx3 = max3(mx3, __uint_as_float(a[c + 6]), __uint_as_float(a[c + 7]));
          }
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;

        if (j == 0) {
          m_used = m_tile;
        } else {
          const bool grow = m_tile > m_used + 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            // O_t must be quiescent: PV(j-1) has finished (S(j) being ready already implies it: QK(j) was issued behind
            // PV(j-1) on the in-order pipe; the wait makes the dependency explicit)
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
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
        }

        const uint64_t negm2 = pack2(-m_used, -m_used);
        // pass 2: exponentials; S is re-read from TMEM, P (packed bf16) goes over the first 32 columns of S
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t a[32];
          tmem_ld_x32(tm_s + h * 32, a);
          tmem_ld_wait();
          if (kv_valid < 64) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (h * 32 + c >= kv_valid) a[c] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const uint64_t x = ffma2(pack2(__uint_as_float(a[c]), __uint_as_float(a[c + 1])), sc2, negm2);
            float x0, x1;
            unpack2(x, x0, x1);
            const float p0 = (kPolyEvery > 0 && (c % (2 * (kPolyEvery > 0 ? kPolyEvery : 1))) == 0) ? poly_exp2(x0) : fast_exp2(x0);
            const float p1 = fast_exp2(x1);
            lsum = fadd2(lsum, pack2(p0, p1));
            a[c >> 1] = pack_bf16x2(p0, p1);
          }
          // P half h goes to columns [16h, 16h+16); the other half's scores (columns 32..63) are still intact
          tmem_st_x16(tm_s + h * 16, a);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }

      mbar_wait(&pv_done[t], (num_kv - 1) & 1);
      tc_fence_after();
      float la, lb;
      unpack2(lsum, la, lb);
      const float inv_l = 1.0f / (la + lb);
      const bool row_ok = row_in_cta < q_valid;
      __nv_bfloat16* orow = p.out + size_t(row_base + q0 + row_in_cta) * p.ldo + head * 64;
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
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pi3b

using namespace pi3b;

#ifdef PI3B_TRACE
extern "C" int pi3b_debug_set_trace(void* ptr) {
  return cudaMemcpyToSymbol(g_trace, &ptr, sizeof(ptr)) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int pi3b_attn_fwd_hd64(void* stream, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                                  int64_t ldv, int q_col0, int k_col0, int v_col0, void* out, int64_t ldo,
                                  int num_seqs, int q_len, int kv_len, int num_heads, float scale) {
  PI3B_REQUIRE(q && k && v && out, "pi3b_attn_fwd_hd64: null pointer");
  PI3B_REQUIRE(num_seqs > 0 && q_len > 0 && kv_len > 0 && num_heads > 0, "pi3b_attn_fwd_hd64: bad shape");
  PI3B_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "pi3b_attn_fwd_hd64: row pitches must be multiples of 8 elements");
  PI3B_REQUIRE(q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0, "pi3b_attn_fwd_hd64: column offsets %% 8");
  PI3B_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "pi3b_attn_fwd_hd64: out must be 16-byte aligned");
  const uint64_t rows = uint64_t(num_seqs) * q_len;
  const uint64_t kv_rows = uint64_t(num_seqs) * kv_len;
  PI3B_REQUIRE(rows < (1ull << 31) && kv_rows < (1ull << 31), "pi3b_attn_fwd_hd64: too many rows");
  // Default: revision 9 (four tiles) for short sequences — frame-local attention, +9 % at S = 1004 — and revision 8
  // (three tiles) for long ones (chunk-global attention, +1.5 % at S = 100400); same-box A/B in
  // profiles/r01_attn_rev8_ab.log. PI3B_ATTN_REV=5|8|9 forces one kernel (A/B runs).
  static const int forced = []() {
    const char* e = getenv("PI3B_ATTN_REV");
    const int r = e ? atoi(e) : 0;
    return (r == 5 || r == 8 || r == 9) ? r : 0;
  }();
  const int rev = forced ? forced : (kv_len <= 8192 ? 9 : 8);
  const int rows_per_cta = rev == 9 ? 512 : (rev == 8 ? 384 : 256);
  const int64_t items = int64_t((q_len + rows_per_cta - 1) / rows_per_cta) * num_seqs * num_heads;
  PI3B_REQUIRE(items < (1ll << 31), "pi3b_attn_fwd_hd64: too many work items");

  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_2d(&tmQ, q, 0, rows, uint64_t(q_col0) + 64ull * num_heads, uint64_t(ldq) * 2, 128, 64))) return rc;
  const uint32_t kv_box = rev >= 8 ? 64 : 128;
  if ((rc = make_tmap_2d(&tmK, k, 0, kv_rows, uint64_t(k_col0) + 64ull * num_heads, uint64_t(ldk) * 2, kv_box, 64))) return rc;
  if ((rc = make_tmap_2d(&tmV, v, 0, kv_rows, uint64_t(v_col0) + 64ull * num_heads, uint64_t(ldv) * 2, kv_box, 64))) return rc;

  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL) !=
        cudaSuccess)
      return check_launch("attn: cudaFuncSetAttribute");
    if (cudaFuncSetAttribute(attn_fwd_kernel_t3, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Smem::TOTAL) !=
        cudaSuccess)
      return check_launch("attn: cudaFuncSetAttribute (t3)");
    if (cudaFuncSetAttribute(attn_fwd_kernel_t4, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn4Smem::TOTAL) !=
        cudaSuccess)
      return check_launch("attn: cudaFuncSetAttribute (t4)");
    configured = true;
  }
  AttnParams p;
  p.num_seqs = num_seqs; p.q_len = q_len; p.kv_len = kv_len; p.num_heads = num_heads;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = (int)ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (rev == 9)
    attn_fwd_kernel_t4<<<(unsigned)items, kA4Threads, Attn4Smem::TOTAL, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  else if (rev == 8)
    attn_fwd_kernel_t3<<<(unsigned)items, kA3Threads, Attn3Smem::TOTAL, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  else
    attn_fwd_kernel<<<(unsigned)items, kAttnThreads, AttnSmem::TOTAL, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  return check_launch("attn launch");
}
