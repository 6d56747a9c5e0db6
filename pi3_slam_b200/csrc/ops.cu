// HBM-bound kernels of the pi3 chunk path: LayerNorm, QK-norm + 2-D RoPE, patch packing, token-table
// broadcast, head unpatchify (+exp, xy*z, pose transform), token mean-pool, 9D->SO(3) pose, validity mask.
// All are coalesced / 16-byte vectorised streaming kernels; none reshapes its work into a GEMM.
#include "host_common.h"
#include "qk_math.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace pi3b {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------
// LayerNorm over D = 128 * NV (1024 for the pi3 trunk; 384 / 768 for the MoGe ViT-S / ViT-B encoders)
// (reference: nn.LayerNorm(eps=1e-6) in dinov2/layers/block.py:88-113 and layers/block.py:310-335; encoder
// final norm vision_transformer.py:271). One warp per row, the row lives in registers (NV x float4 per
// lane), two-pass statistics in fp32 like torch.
// Optional row override: rows with (row % period) < n_override are replaced by override[row % period]
// (used to drop the encoder's cls/register tokens and install the decoder's register tokens,
// pi3.py:140-144, in the same pass).
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t ldx,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, void* out,
                                                        int64_t ldo, int out_f32, int M,
                                                        const float* __restrict__ override_rows, int period,
                                                        int n_override) {
  constexpr int D = NV * 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float4 v[NV];
  if (override_rows && (row % period) < n_override) {
    const float4* src = reinterpret_cast<const float4*>(override_rows + size_t(row % period) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = src[i * 32 + lane];
  } else {
    const float4* src = reinterpret_cast<const float4*>(x + size_t(row) * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = src[i * 32 + lane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / float(D));
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / float(D)) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
      v[i].x = (v[i].x - mean) * rstd * g.x + b.x;
      v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
      v[i].z = (v[i].z - mean) * rstd * g.z + b.z;
      v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
    }
  }
  if (out_f32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out) + size_t(row) * ldo);
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i * 32 + lane] = v[i];
  } else {
    uint2* dst = reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + size_t(row) * ldo);
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i * 32 + lane] = make_uint2(bf16x2(v[i].x, v[i].y), bf16x2(v[i].z, v[i].w));
  }
}

// ------------------------------------------------------------------------------------------------
// QK LayerNorm(64) + RoPE-2D, in place on the q and k thirds of the packed qkv buffer.
// Reference: attention.py:330-334 (q_norm/k_norm -> .to(bf16) -> rope), pos_embed.py:120-159 and the
// native spec curope/kernels.cu:17-82: per head, dims [0,32) rotate by the y position, [32,64) by x;
// pairs (j, j+16), theta_j = pos * base^(-j/16). Register tokens have pos (0,0) -> identity
// (pi3.py:146-154). 8 lanes share one (row, head) so every access is a fully coalesced 16-byte vector;
// the rotation partner sits 2 lanes away (shfl_xor 2). cs_table[pos][j] = (cos, sin) in fp32.
// ------------------------------------------------------------------------------------------------
constexpr int kQkRows = 4;  // rows per thread: 4 independent 16-byte loads in flight (the 1-row version was latency-bound at 2.7 TB/s)

__global__ void __launch_bounds__(256) qknorm_rope_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ld, int M,
                                                          int num_heads, int tokens_per_frame, int n_special,
                                                          int patch_w, const float* __restrict__ qn_w,
                                                          const float* __restrict__ qn_b,
                                                          const float* __restrict__ kn_w,
                                                          const float* __restrict__ kn_b, float eps,
                                                          const float2* __restrict__ cs_table, int do_rope,
                                                          float q_scale) {
  // 32-bit index math only: the first version spent most of its instructions in 64-bit divisions
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int sub = int(gid & 7u);
  const uint32_t unit = gid >> 3;  // (row group, which, head)
  const uint32_t heads2 = 2u * uint32_t(num_heads);
  const uint32_t grp = unit / heads2;
  const int row0 = int(grp) * kQkRows;
  if (row0 >= M) return;  // whole 8-lane groups exit together
  const int hh = int(unit - grp * heads2);
  const int which = hh >= num_heads ? 1 : 0;  // 0 = q, 1 = k
  const int head = hh - which * num_heads;
  __nv_bfloat16* base = qkv + which * num_heads * 64 + head * 64 + sub * 8;
  uint4 raw[kQkRows];
#pragma unroll
  for (int r = 0; r < kQkRows; ++r)
    if (row0 + r < M) raw[r] = *reinterpret_cast<const uint4*>(base + int64_t(row0 + r) * ld);
  const float* nw = which ? kn_w : qn_w;
  const float* nb = which ? kn_b : qn_b;
  float w8[8], b8[8];
  if (nw) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { w8[i] = __ldg(nw + sub * 8 + i); b8[i] = __ldg(nb + sub * 8 + i); }
  }
  const int sub4 = sub & 3;
  const int j0 = (sub4 & 1) * 8;
  const bool is_v = sub4 >= 2;  // second member of the (j, j+16) pair
#pragma unroll
  for (int r = 0; r < kQkRows; ++r) {
    const int row = row0 + r;
    const bool ok = row < M;  // uniform inside an 8-lane group; shuffles below stay convergent for the full warp
    float x[8] = {bf_lo(raw[r].x), bf_hi(raw[r].x), bf_lo(raw[r].y), bf_hi(raw[r].y),
                  bf_lo(raw[r].z), bf_hi(raw[r].z), bf_lo(raw[r].w), bf_hi(raw[r].w)};
    // arithmetic order shared with the fused QKV-GEMM epilogue (qk_math.cuh): bit-identical results
    if (nw) {
      float s = qk_sum8(x);
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
      s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
      const float mean = qk_mean(s);
      float q = qk_sq8(x, mean);
      q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 1));
      q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 2));
      q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 4));
      const float rstd = qk_rstd(q, eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = qk_norm1(x[i], mean, rstd, w8[i], b8[i]);
    }
    if (do_rope) {
      const int t = int(uint32_t(row) % uint32_t(tokens_per_frame));
      int pos = 0;
      if (ok && t >= n_special) {
        const uint32_t pidx = uint32_t(t - n_special);
        const uint32_t py = pidx / uint32_t(patch_w);
        pos = int(sub < 4 ? py : pidx - py * uint32_t(patch_w)) + 1;
      }
      float partner[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) partner[i] = __shfl_xor_sync(0xffffffffu, x[i], 2);
      if (pos != 0) {
        const float4* cs = reinterpret_cast<const float4*>(cs_table + size_t(pos) * 16 + j0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 c = __ldg(cs + i);  // (cos, sin) of two consecutive frequencies
          x[2 * i] = is_v ? qk_rope_second(partner[2 * i], x[2 * i], c.x, c.y) : qk_rope_first(x[2 * i], partner[2 * i], c.x, c.y);
          x[2 * i + 1] = is_v ? qk_rope_second(partner[2 * i + 1], x[2 * i + 1], c.z, c.w)
                              : qk_rope_first(x[2 * i + 1], partner[2 * i + 1], c.z, c.w);
        }
      }
    }
    if (which == 0 && q_scale != 1.0f) {   // queries leave pre-multiplied (softmax scale folded in), rounded to bf16 once
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = qk_scale_out(x[i], q_scale);
    }
    if (ok) {
      uint4 o;
      o.x = bf16x2(x[0], x[1]); o.y = bf16x2(x[2], x[3]); o.z = bf16x2(x[4], x[5]); o.w = bf16x2(x[6], x[7]);
      *reinterpret_cast<uint4*>(base + int64_t(row) * ld) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Patch packing: imgs fp32 [N,3,H,W] in [0,1] -> A bf16 [N*P, ldk] with column c*196 + i*14 + j,
// normalised by the ImageNet mean/std (pi3.py:174) — i.e. the im2col of the 14x14/s14 patch-embed conv
// (patch_embed.py:65) so that the conv becomes one GEMM. Columns [588, ldk) are zero padding.
// One thread per (patch, channel, row-in-patch): 14 contiguous pixels in, 14 contiguous bf16 out.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_pack_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                         int N, int H, int W, int ldk, float m0, float m1, float m2,
                                                         float is0, float is1, float is2) {
  const int ph = H / 14, pw = W / 14;
  const int64_t total = int64_t(N) * ph * pw * 42;
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int ci = int(gid % 42);  // c*14 + i
  const int64_t patch = gid / 42;
  const int c = ci / 14, i = ci % 14;
  const int px = int(patch % pw);
  const int py = int((patch / pw) % ph);
  const int n = int(patch / (int64_t(pw) * ph));
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
  const float istd = c == 0 ? is0 : (c == 1 ? is1 : is2);
  const float* src = img + ((int64_t(n) * 3 + c) * H + (py * 14 + i)) * W + px * 14;
  __nv_bfloat16* dst = out + patch * ldk + c * 196 + i * 14;
#pragma unroll
  for (int j = 0; j < 14; ++j) dst[j] = __float2bfloat16((src[j] - mean) / istd);
  if (ci == 0) {
    for (int k = 588; k < ldk; ++k) out[patch * ldk + k] = __float2bfloat16(0.f);
  }
}

// out[n, t, :] = table[t, :]  (fp32, D = 1024): installs cls+pos / register / interpolated pos-embed rows
// (vision_transformer.py:215-234) before the patch-embed GEMM accumulates onto the patch rows.
__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float4* __restrict__ table, float4* __restrict__ out,
                                                             int64_t total_vec, int64_t period_vec) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total_vec) return;
  out[gid] = __ldg(table + gid % period_vec);
}

// ------------------------------------------------------------------------------------------------
// Head unpatchify (transformer_head.py:76-81 pixel_shuffle(14) + pi3.py:194-209):
//   out[n, 14y+i, 14x+j, ci] = feat[n, n_special + y*w + x, ci*196 + 14 i + j]
//   z = exp(z); local = (x z, y z, z); points = R local + t; conf = raw logit.
// One thread per pixel; writes are contiguous (12 / 12 / 4 bytes per pixel).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpatchify_kernel(const float* __restrict__ fpts, int64_t ld_pts,
                                                         const float* __restrict__ fconf, int64_t ld_conf,
                                                         const float* __restrict__ poses, float* __restrict__ local_pts,
                                                         float* __restrict__ points, float* __restrict__ conf, int N,
                                                         int H, int W, int tokens_per_frame, int n_special) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(N) * H * W;
  if (gid >= total) return;
  const int xpix = int(gid % W);
  const int ypix = int((gid / W) % H);
  const int n = int(gid / (int64_t(W) * H));
  const int pw = W / 14;
  const int token = n_special + (ypix / 14) * pw + xpix / 14;
  const int inner = (ypix % 14) * 14 + xpix % 14;
  const int64_t trow = int64_t(n) * tokens_per_frame + token;
  const float* fp = fpts + trow * ld_pts + inner;
  const float fx = fp[0], fy = fp[196], fz = fp[392];
  const float z = expf(fz);
  const float lx = fx * z, ly = fy * z;
  const float* P = poses + n * 16;
  const float wx = P[0] * lx + P[1] * ly + P[2] * z + P[3];
  const float wy = P[4] * lx + P[5] * ly + P[6] * z + P[7];
  const float wz = P[8] * lx + P[9] * ly + P[10] * z + P[11];
  float* lo = local_pts + gid * 3;
  lo[0] = lx; lo[1] = ly; lo[2] = z;
  float* po = points + gid * 3;
  po[0] = wx; po[1] = wy; po[2] = wz;
  conf[gid] = fconf[trow * ld_conf + inner];
}

// fp32 -> bf16 cast of a [M, C] matrix with independent row pitches (decoder taps -> concat buffer,
// pi3.py:168-171; fp32 residual -> bf16 GEMM operand, as autocast does for nn.Linear inputs).
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, int64_t ldx,
                                                        __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t M,
                                                        int C4) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= M * C4) return;
  const int64_t row = gid / C4;
  const int c = int(gid % C4);
  const float4 v = *reinterpret_cast<const float4*>(x + row * ldx + c * 4);
  *reinterpret_cast<uint2*>(out + row * ldo + c * 4) = make_uint2(bf16x2(v.x, v.y), bf16x2(v.z, v.w));
}

// Mean over the patch tokens of each frame (camera_head.py:55): x fp32 [N*T, C] -> out [N, C].
__global__ void __launch_bounds__(256) meanpool_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out,
                                                       int C, int tokens_per_frame, int n_special) {
  __shared__ float red[8][33];
  const int n = blockIdx.y;
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  float s = 0.f;
  if (col < C)
    for (int t = n_special + ry; t < tokens_per_frame; t += 8) s += x[(int64_t(n) * tokens_per_frame + t) * ldx + col];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && col < C) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
    out[int64_t(n) * C + col] = tot / float(tokens_per_frame - n_special);
  }
}

// ------------------------------------------------------------------------------------------------
// 9D -> SO(3) (camera_head.py:66-93): rows of m normalised, R = V diag(1,1,det(V U^T)) U^T for
// m_n^T = U S V^T, i.e. the rotation nearest to m_n. Using the two leading singular pairs (w_i, z_i) of
// B = m_n (B z_i = s_i w_i):  R = [w1 w2 w1 x w2] [z1 z2 z1 x z2]^T  — the sign ambiguities of the third
// pair cancel, so no explicit det fix-up is needed. Jacobi eigen-solve of B^T B in fp64, one thread per pose.
// ------------------------------------------------------------------------------------------------
__device__ void jacobi_eig3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// Nearest rotation to B (3x3, row-major) via its two leading singular pairs. Returns R row-major.
__device__ void nearest_rotation(const double B[3][3], double R[3][3]) {
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = B[0][i] * B[0][j] + B[1][i] * B[1][j] + B[2][i] * B[2][j];
  jacobi_eig3(A, V);
  int i0 = 0;
  if (A[1][1] > A[i0][i0]) i0 = 1;
  if (A[2][2] > A[i0][i0]) i0 = 2;
  int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
  if (A[i2][i2] > A[i1][i1]) { const int tmp = i1; i1 = i2; i2 = tmp; }
  double z1[3], z2[3], w1[3], w2[3];
  for (int k = 0; k < 3; ++k) { z1[k] = V[k][i0]; z2[k] = V[k][i1]; }
  double n1 = 0, n2 = 0;
  for (int r = 0; r < 3; ++r) {
    w1[r] = B[r][0] * z1[0] + B[r][1] * z1[1] + B[r][2] * z1[2];
    w2[r] = B[r][0] * z2[0] + B[r][1] * z2[1] + B[r][2] * z2[2];
    n1 += w1[r] * w1[r];
  }
  n1 = sqrt(n1);
  for (int r = 0; r < 3; ++r) w1[r] /= n1;
  // Gram-Schmidt keeps w2 orthogonal to w1 when s2 is small
  double d = w1[0] * w2[0] + w1[1] * w2[1] + w1[2] * w2[2];
  for (int r = 0; r < 3; ++r) { w2[r] -= d * w1[r]; n2 += w2[r] * w2[r]; }
  n2 = sqrt(n2);
  for (int r = 0; r < 3; ++r) w2[r] /= n2;
  const double w3[3] = {w1[1] * w2[2] - w1[2] * w2[1], w1[2] * w2[0] - w1[0] * w2[2], w1[0] * w2[1] - w1[1] * w2[0]};
  const double z3[3] = {z1[1] * z2[2] - z1[2] * z2[1], z1[2] * z2[0] - z1[0] * z2[2], z1[0] * z2[1] - z1[1] * z2[0]};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r][c] = w1[r] * z1[c] + w2[r] * z2[c] + w3[r] * z3[c];
}

__global__ void pose_from_9d_kernel(const float* __restrict__ rot9, int64_t ld_r, const float* __restrict__ trans,
                                    int64_t ld_t, float* __restrict__ poses, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double B[3][3], R[3][3];
  for (int r = 0; r < 3; ++r) {
    const float a = rot9[n * ld_r + r * 3], b = rot9[n * ld_r + r * 3 + 1], c = rot9[n * ld_r + r * 3 + 2];
    const float nrm = fmaxf(sqrtf(a * a + b * b + c * c), 1e-12f);  // F.normalize(p=2, dim=-1, eps=1e-12)
    B[r][0] = a / nrm; B[r][1] = b / nrm; B[r][2] = c / nrm;
  }
  nearest_rotation(B, R);
  float* P = poses + n * 16;
  for (int r = 0; r < 3; ++r) {
    P[r * 4 + 0] = float(R[r][0]); P[r * 4 + 1] = float(R[r][1]); P[r * 4 + 2] = float(R[r][2]);
    P[r * 4 + 3] = trans[n * ld_t + r];
  }
  P[12] = 0.f; P[13] = 0.f; P[14] = 0.f; P[15] = 1.f;
}

// ------------------------------------------------------------------------------------------------
// Validity mask (offline_chunk_creator.py:114-119 + pi3/utils/geometry.py:347-375):
//   mask = sigmoid(conf) > thr  &&  !((max3x3(z) - min3x3(z)) / z > rtol), -inf padding at the border.
// z is read with a stride (the z channel of local_points). One thread per pixel; the 3x3 window re-reads hit L1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_kernel(const float* __restrict__ z, int z_stride, const float* __restrict__ conf,
                                                   uint8_t* __restrict__ mask, int N, int H, int W, float conf_thr,
                                                   float rtol) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(N) * H * W;
  if (gid >= total) return;
  const int x = int(gid % W);
  const int y = int((gid / W) % H);
  const int64_t base = gid - int64_t(y) * W - x;
  float mx = -INFINITY, mn = INFINITY;
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      const float v = z[(base + int64_t(yy) * W + xx) * z_stride];
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
  }
  const float zc = z[gid * z_stride];
  const float diff = mx + (-mn);
  float rel = diff / zc;
  if (isnan(rel)) rel = 0.f;                       // nan_to_num_
  else if (isinf(rel)) rel = rel > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  const bool edge = rel > rtol;
  const float sg = 1.0f / (1.0f + expf(-conf[gid]));
  mask[gid] = (sg > conf_thr && !edge) ? 1 : 0;
}

}  // namespace pi3b

using namespace pi3b;

static inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" int pi3b_layernorm(void* stream, const float* x, int64_t ldx, const float* gamma, const float* beta,
                              float eps, void* out, int64_t ldo, int out_f32, int M, int D,
                              const float* override_rows, int period, int n_override) {
  PI3B_REQUIRE(x && gamma && beta && out, "pi3b_layernorm: null pointer");
  PI3B_REQUIRE(D == 1024 || D == 768 || D == 384, "pi3b_layernorm: D must be 384, 768 or 1024 (got %d)", D);
  PI3B_REQUIRE(M > 0 && ldx % 4 == 0 && ldo % 4 == 0, "pi3b_layernorm: bad M/ld");
  PI3B_REQUIRE(!override_rows || (period > 0 && n_override >= 0 && n_override <= period), "pi3b_layernorm: bad override");
  auto kern = D == 1024 ? layernorm_kernel<8> : (D == 768 ? layernorm_kernel<6> : layernorm_kernel<3>);
  kern<<<blocks_for(M, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, gamma, beta, eps, out, ldo, out_f32, M, override_rows, period > 0 ? period : 1, n_override);
  return check_launch("layernorm launch");
}

extern "C" int pi3b_qknorm_rope(void* stream, void* qkv, int64_t ld, int M, int num_heads, int tokens_per_frame,
                                int n_special, int patch_w, const float* qn_w, const float* qn_b, const float* kn_w,
                                const float* kn_b, float eps, const float* cs_table, int table_rows, int do_rope,
                                float q_scale) {
  PI3B_REQUIRE(qkv && M > 0 && num_heads > 0, "pi3b_qknorm_rope: bad args");
  PI3B_REQUIRE(q_scale > 0.f, "pi3b_qknorm_rope: q_scale must be positive (1 = none)");
  PI3B_REQUIRE(ld % 8 == 0 && ld >= 3 * 64 * num_heads, "pi3b_qknorm_rope: bad ld");
  PI3B_REQUIRE((qn_w == nullptr) == (kn_w == nullptr) && (qn_w == nullptr) == (qn_b == nullptr) &&
                   (kn_w == nullptr) == (kn_b == nullptr), "pi3b_qknorm_rope: norm params must be all set or all null");
  if (do_rope) {
    PI3B_REQUIRE(cs_table && tokens_per_frame > n_special && patch_w > 0, "pi3b_qknorm_rope: rope needs table/geometry");
    const int ph = (tokens_per_frame - n_special + patch_w - 1) / patch_w;
    PI3B_REQUIRE(table_rows > (ph > patch_w ? ph : patch_w), "pi3b_qknorm_rope: cos/sin table too short");
  }
  const int64_t threads = int64_t((M + pi3b::kQkRows - 1) / pi3b::kQkRows) * 2 * num_heads * 8;
  PI3B_REQUIRE(threads < (1ll << 31), "pi3b_qknorm_rope: too many rows");
  qknorm_rope_kernel<<<blocks_for(threads, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(qkv), ld, M, num_heads, tokens_per_frame, n_special, patch_w, qn_w, qn_b, kn_w, kn_b,
      eps, reinterpret_cast<const float2*>(cs_table), do_rope, q_scale);
  return check_launch("qknorm_rope launch");
}

extern "C" int pi3b_patch_pack(void* stream, const float* imgs, void* out, int N, int H, int W, int ldk,
                               const float* mean3, const float* std3) {
  PI3B_REQUIRE(imgs && out && mean3 && std3, "pi3b_patch_pack: null pointer");
  PI3B_REQUIRE(N > 0 && H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0, "pi3b_patch_pack: H, W must be multiples of 14");
  PI3B_REQUIRE(ldk >= 588 && ldk % 8 == 0, "pi3b_patch_pack: ldk must be >= 588 and a multiple of 8");
  const int64_t total = int64_t(N) * (H / 14) * (W / 14) * 42;
  patch_pack_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      imgs, static_cast<__nv_bfloat16*>(out), N, H, W, ldk, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  return check_launch("patch_pack launch");
}

extern "C" int pi3b_broadcast_rows(void* stream, const float* table, float* out, int num_groups, int rows_per_group,
                                   int D) {
  PI3B_REQUIRE(table && out && num_groups > 0 && rows_per_group > 0 && D % 4 == 0, "pi3b_broadcast_rows: bad args");
  const int64_t period_vec = int64_t(rows_per_group) * D / 4;
  const int64_t total_vec = period_vec * num_groups;
  broadcast_rows_kernel<<<blocks_for(total_vec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(table), reinterpret_cast<float4*>(out), total_vec, period_vec);
  return check_launch("broadcast_rows launch");
}

extern "C" int pi3b_unpatchify_points_conf(void* stream, const float* feat_pts, int64_t ld_pts, const float* feat_conf,
                                           int64_t ld_conf, const float* poses, float* local_points, float* points,
                                           float* conf, int N, int H, int W, int tokens_per_frame, int n_special) {
  PI3B_REQUIRE(feat_pts && feat_conf && poses && local_points && points && conf, "pi3b_unpatchify: null pointer");
  PI3B_REQUIRE(H % 14 == 0 && W % 14 == 0 && tokens_per_frame == n_special + (H / 14) * (W / 14),
               "pi3b_unpatchify: geometry mismatch");
  PI3B_REQUIRE(ld_pts >= 588 && ld_conf >= 196, "pi3b_unpatchify: bad ld");
  const int64_t total = int64_t(N) * H * W;
  unpatchify_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      feat_pts, ld_pts, feat_conf, ld_conf, poses, local_points, points, conf, N, H, W, tokens_per_frame, n_special);
  return check_launch("unpatchify launch");
}

extern "C" int pi3b_cast_f32_bf16(void* stream, const float* x, int64_t ldx, void* out, int64_t ldo, int M, int C) {
  PI3B_REQUIRE(x && out && M > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "pi3b_cast_f32_bf16: bad args");
  const int64_t total = int64_t(M) * (C / 4);
  cast_bf16_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, static_cast<__nv_bfloat16*>(out), ldo, M, C / 4);
  return check_launch("cast launch");
}

extern "C" int pi3b_meanpool_tokens(void* stream, const float* x, int64_t ldx, float* out, int N, int C,
                                    int tokens_per_frame, int n_special) {
  PI3B_REQUIRE(x && out && N > 0 && C > 0 && tokens_per_frame > n_special, "pi3b_meanpool_tokens: bad args");
  dim3 grid((C + 31) / 32, N);
  meanpool_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, out, C, tokens_per_frame, n_special);
  return check_launch("meanpool launch");
}

extern "C" int pi3b_pose_from_9d(void* stream, const float* rot9, int64_t ld_r, const float* trans, int64_t ld_t,
                                 float* poses, int N) {
  PI3B_REQUIRE(rot9 && trans && poses && N > 0 && ld_r >= 9 && ld_t >= 3, "pi3b_pose_from_9d: bad args");
  pose_from_9d_kernel<<<blocks_for(N, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(rot9, ld_r, trans, ld_t, poses, N);
  return check_launch("pose_from_9d launch");
}

extern "C" int pi3b_mask_depth_edge(void* stream, const float* z, int z_stride, const float* conf, uint8_t* mask, int N,
                                    int H, int W, float conf_thr, float rtol) {
  PI3B_REQUIRE(z && conf && mask && N > 0 && H > 0 && W > 0 && z_stride > 0, "pi3b_mask_depth_edge: bad args");
  const int64_t total = int64_t(N) * H * W;
  mask_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, z_stride, conf, mask, N, H, W,
                                                                                   conf_thr, rtol);
  return check_launch("mask launch");
}
