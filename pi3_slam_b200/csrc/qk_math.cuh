// q/k LayerNorm(64) + RoPE-2D arithmetic shared by the stand-alone pass (ops.cu: qknorm_rope_kernel, 8 lanes per head) and
// the fused QKV-GEMM epilogue (gemm.cu, one thread per row). Reference: pi3/models/layers/attention.py:330-334 (q_norm /
// k_norm then rope), pi3/models/layers/pos_embed.py:112-159 (RoPE2D), pi3/models/curope/kernels.cu:17-82 (the native
// kernel the reference ships). The operation ORDER is fixed here with explicit round-to-nearest intrinsics (no compiler
// FMA contraction), so the two kernels produce bit-identical results and the fused epilogue can be tested for equality:
//   * the 64 values of a head are summed as 8 sequential partial sums of 8 consecutive channels, combined as
//     ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)) (the xor-1/2/4 butterfly of the 8-lane version);
//   * variance = the same tree over fma(d, d, acc), d = x - mean; rstd = rsqrtf(var / 64 + eps);
//   * y = fma((x - mean) * rstd, w, b), rounded to bf16 (the reference casts the normed q/k to bf16 before RoPE);
//   * RoPE pair (a, b) = channels (j, j + 16) of a 32-channel half, angle pos * base^(-j/16):
//     a' = fma(a, cos, -(b * sin)), b' = fma(b, cos, a * sin); pos = 0 (special tokens) is the identity;
//   * queries only, optional: x q_scale (one fp32 multiply) before the final rounding to bf16.
#pragma once
#include <cuda_bf16.h>

namespace pi3b {

__device__ __forceinline__ float qk_round_bf16(float y) { return __bfloat162float(__float2bfloat16(y)); }

__device__ __forceinline__ float qk_sum8(const float* x) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = __fadd_rn(s, x[i]);
  return s;
}
__device__ __forceinline__ float qk_sq8(const float* x, float mean) {
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = __fsub_rn(x[i], mean);
    q = __fmaf_rn(d, d, q);
  }
  return q;
}
__device__ __forceinline__ float qk_tree8(const float* p) {
  return __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), __fadd_rn(p[2], p[3])),
                   __fadd_rn(__fadd_rn(p[4], p[5]), __fadd_rn(p[6], p[7])));
}
__device__ __forceinline__ float qk_mean(float sum) { return __fmul_rn(sum, 1.0f / 64.0f); }
__device__ __forceinline__ float qk_rstd(float sq, float eps) { return rsqrtf(__fadd_rn(__fmul_rn(sq, 1.0f / 64.0f), eps)); }
__device__ __forceinline__ float qk_norm1(float x, float mean, float rstd, float w, float b) {
  return qk_round_bf16(__fmaf_rn(__fmul_rn(__fsub_rn(x, mean), rstd), w, b));
}
// (a, b) = (first, second) member of a rotation pair
__device__ __forceinline__ float qk_rope_first(float a, float b, float c, float s) { return __fmaf_rn(a, c, -__fmul_rn(b, s)); }
__device__ __forceinline__ float qk_rope_second(float a, float b, float c, float s) { return __fmaf_rn(b, c, __fmul_rn(a, s)); }

// optional last step for the queries: multiply by q_scale (the softmax scale x log2(e) folded into q, so that the attention
// kernel can feed its accumulator straight to exp2) before the one rounding to bf16
__device__ __forceinline__ float qk_scale_out(float x, float s) { return __fmul_rn(x, s); }

// 16-byte load from shared memory by 32-bit shared address (pointers derived from the manually aligned dynamic smem base
// are generic to the compiler, which would emit slower generic loads)
__device__ __forceinline__ float4 qk_lds128(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}
// two values rounded to bf16 at once (one F2FP pack + two ALU unpacks instead of two XU-pipe F2F conversions + shifts);
// same round-to-nearest-even as qk_round_bf16
__device__ __forceinline__ void qk_round_bf16_x2(float& a, float& b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t u = *reinterpret_cast<uint32_t*>(&h);
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}

// One whole head held by one thread: v[64] (fp32 values that are exact bf16 numbers) -> LayerNorm (if norm_saddr != 0: the
// shared-memory address of 64 weights followed by 64 biases) -> RoPE with positions (pos_y, pos_x) (if cs != nullptr;
// cs = fp32 [pos][16] (cos, sin); cs_saddr != 0: the same table in shared memory, rows cs_row_bytes apart). Results are
// left in v (unrounded after RoPE: the caller packs to bf16).
__device__ __forceinline__ void qk_head_inplace(float* v, uint32_t norm_saddr, float eps, const float2* cs, int pos_y,
                                                int pos_x, uint32_t cs_saddr = 0, uint32_t cs_row_bytes = 144) {
  if (norm_saddr != 0) {
    float part[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) part[g] = qk_sum8(v + 8 * g);
    const float mean = qk_mean(qk_tree8(part));
#pragma unroll
    for (int g = 0; g < 8; ++g) part[g] = qk_sq8(v + 8 * g, mean);
    const float rstd = qk_rstd(qk_tree8(part), eps);
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 w4 = qk_lds128(norm_saddr + i * 4);
      const float4 b4 = qk_lds128(norm_saddr + 256 + i * 4);
      float y0 = __fmaf_rn(__fmul_rn(__fsub_rn(v[i], mean), rstd), w4.x, b4.x);
      float y1 = __fmaf_rn(__fmul_rn(__fsub_rn(v[i + 1], mean), rstd), w4.y, b4.y);
      float y2 = __fmaf_rn(__fmul_rn(__fsub_rn(v[i + 2], mean), rstd), w4.z, b4.z);
      float y3 = __fmaf_rn(__fmul_rn(__fsub_rn(v[i + 3], mean), rstd), w4.w, b4.w);
      qk_round_bf16_x2(y0, y1);
      qk_round_bf16_x2(y2, y3);
      v[i] = y0; v[i + 1] = y1; v[i + 2] = y2; v[i + 3] = y3;
    }
  }
  if (cs != nullptr) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pos = h == 0 ? pos_y : pos_x;
      if (pos != 0) {
        const float4* t = reinterpret_cast<const float4*>(cs + size_t(pos) * 16);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          // (cos, sin) of frequencies j and j + 1
          const float4 c = cs_saddr ? qk_lds128(cs_saddr + uint32_t(pos) * cs_row_bytes + (j >> 1) * 16) : __ldg(t + (j >> 1));
          const float a0 = v[32 * h + j], b0 = v[32 * h + j + 16];
          const float a1 = v[32 * h + j + 1], b1 = v[32 * h + j + 17];
          v[32 * h + j] = qk_rope_first(a0, b0, c.x, c.y);
          v[32 * h + j + 16] = qk_rope_second(a0, b0, c.x, c.y);
          v[32 * h + j + 1] = qk_rope_first(a1, b1, c.z, c.w);
          v[32 * h + j + 17] = qk_rope_second(a1, b1, c.z, c.w);
        }
      }
    }
  }
}

// token position of a row: (0, 0) for the special tokens, otherwise (y, x) + 1 on the patch grid
__device__ __forceinline__ void qk_row_position(int row, int tokens_per_frame, int n_special, int patch_w, int& pos_y,
                                                int& pos_x) {
  const int t = int(uint32_t(row) % uint32_t(tokens_per_frame));
  pos_y = 0;
  pos_x = 0;
  if (t >= n_special) {
    const uint32_t pidx = uint32_t(t - n_special);
    const uint32_t py = pidx / uint32_t(patch_w);
    pos_y = int(py) + 1;
    pos_x = int(pidx - py * uint32_t(patch_w)) + 1;
  }
}

}  // namespace pi3b
