// Post-processing kernels of the chunk path: keypoint gather (bilinear / nearest grid sampling on the
// device instead of the reference's 350-550 MB D2H + CPU grid_sample), lower-median of a masked ratio
// (MoGe metric scale), strided sample gather for the intrinsics solve, and the batched Umeyama Sim3.
#include "host_common.h"

#include <math.h>

namespace pi3b {

// ------------------------------------------------------------------------------------------------
// Keypoint gather. Reference: offline_chunk_creator.py:129-159 and keypoint_extraction.py:207-229:
//   g = (x / (W-1)) * 2 - 1  (torch elementwise ops on the keypoints' device), then, on the CPU,
//   F.grid_sample(align_corners=False, padding_mode='border'); bilinear for points, local_points and colours
//   (x255, truncated to uint8), nearest (round-half-even) for conf and mask.
// The outputs are byte / index work downstream, so the kernel restates the exact fp32 operation order of what the
// reference executes (identified empirically against torch 2.11 CPU, scripts/probe_grid_sample_order.py):
//   * grid: on a CUDA tensor ATen divides by a scalar as a multiplication by its fp32 reciprocal (grid_math = 1);
//     on a CPU tensor it is a true division (grid_math = 0);
//   * ATen's vectorised CPU grid_sample (AVX2 / AVX-512 builds, FMA-contracted): ix = fma(g + 1, W / 2, -0.5) clipped to
//     [0, W-1]; w = ix - floor(ix), e = 1 - w (likewise n, s); out = fma(se_v, n*w, fma(sw_v, n*e, fma(ne_v, s*w,
//     nw_v * (s*e)))) with out-of-range corners read as 0.
// One thread per (frame, keypoint).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float src_index(float coord, int size, int grid_math) {
  const float q = grid_math ? __fmul_rn(coord, __fdiv_rn(1.0f, float(size - 1))) : __fdiv_rn(coord, float(size - 1));
  const float g = __fsub_rn(__fmul_rn(q, 2.0f), 1.0f);
  float ix = __fmaf_rn(__fadd_rn(g, 1.0f), __fdiv_rn(float(size), 2.0f), -0.5f);
  ix = fminf(float(size - 1), fmaxf(ix, 0.0f));
  return ix;
}

__global__ void __launch_bounds__(128) keypoint_gather_kernel(
    const float* __restrict__ points, const float* __restrict__ local_points, const float* __restrict__ conf,
    const uint8_t* __restrict__ mask, const float* __restrict__ images, const float* __restrict__ kps,
    float* __restrict__ o_points, float* __restrict__ o_local, float* __restrict__ o_conf, uint8_t* __restrict__ o_mask,
    uint8_t* __restrict__ o_colors, int N, int K, int H, int W, int grid_math) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= N * K) return;
  const int n = gid / K;
  const float ix = src_index(kps[gid * 2 + 0], W, grid_math);
  const float iy = src_index(kps[gid * 2 + 1], H, grid_math);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = int(x0f), y0 = int(y0f), x1 = x0 + 1, y1 = y0 + 1;
  const float w = __fsub_rn(ix, x0f), e = __fsub_rn(1.0f, w);
  const float nn = __fsub_rn(iy, y0f), s = __fsub_rn(1.0f, nn);
  const float w_nw = __fmul_rn(s, e), w_ne = __fmul_rn(s, w), w_sw = __fmul_rn(nn, e), w_se = __fmul_rn(nn, w);
  const bool in_x0 = x0 >= 0 && x0 < W, in_x1 = x1 >= 0 && x1 < W;
  const bool in_y0 = y0 >= 0 && y0 < H, in_y1 = y1 >= 0 && y1 < H;
  const size_t fbase = size_t(n) * H * W;

  auto bil = [&](const float* src, size_t stride) {
    const float v_nw = (in_y0 && in_x0) ? src[(size_t(y0) * W + x0) * stride] : 0.f;
    const float v_ne = (in_y0 && in_x1) ? src[(size_t(y0) * W + x1) * stride] : 0.f;
    const float v_sw = (in_y1 && in_x0) ? src[(size_t(y1) * W + x0) * stride] : 0.f;
    const float v_se = (in_y1 && in_x1) ? src[(size_t(y1) * W + x1) * stride] : 0.f;
    return __fmaf_rn(v_se, w_se, __fmaf_rn(v_sw, w_sw, __fmaf_rn(v_ne, w_ne, __fmul_rn(v_nw, w_nw))));
  };
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o_points[gid * 3 + c] = bil(points + fbase * 3 + c, 3);
    o_local[gid * 3 + c] = bil(local_points + fbase * 3 + c, 3);
    if (images) {
      const float col = bil(images + (size_t(n) * 3 + c) * H * W, 1);
      o_colors[gid * 3 + c] = static_cast<uint8_t>(__fmul_rn(col, 255.0f));  // .to(torch.uint8) truncates
    }
  }
  const int xn = int(nearbyintf(ix)), yn = int(nearbyintf(iy));
  const bool in_n = xn >= 0 && xn < W && yn >= 0 && yn < H;
  o_conf[gid] = in_n ? conf[fbase + size_t(yn) * W + xn] : 0.f;
  o_mask[gid] = in_n ? (mask[fbase + size_t(yn) * W + xn] != 0) : 0;
}

// ------------------------------------------------------------------------------------------------
// Strided sample gather for the intrinsics solve (geometry_torch.py:140-142: nearest F.interpolate of
// local_points / mask to 64x64). out_pts [N, hs*ws, 3], out_valid [N, hs*ws] (sigmoid(conf) > thr).
// ------------------------------------------------------------------------------------------------
__global__ void sample_grid_kernel(const float* __restrict__ local_points, const float* __restrict__ conf,
                                   const int* __restrict__ ys, const int* __restrict__ xs, float* __restrict__ out_pts,
                                   uint8_t* __restrict__ out_valid, int N, int H, int W, int hs, int ws, float thr) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= N * hs * ws) return;
  const int n = gid / (hs * ws);
  const int r = gid % (hs * ws);
  const int y = ys[r / ws], x = xs[r % ws];
  const size_t pix = (size_t(n) * H + y) * W + x;
  out_pts[gid * 3 + 0] = local_points[pix * 3 + 0];
  out_pts[gid * 3 + 1] = local_points[pix * 3 + 1];
  out_pts[gid * 3 + 2] = local_points[pix * 3 + 2];
  out_valid[gid] = (1.0f / (1.0f + expf(-conf[pix]))) > thr ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Lower median of a[i] / b[i] over mask[i] != 0 (offline_chunk_creator.py:121-127: torch.median returns
// the lower of the two middle elements). Exact radix select on the order-preserving uint32 key of the fp32
// ratio: 4 passes of 256-bin histograms, single block (the input is one frame, ~1e5 elements).
// out[0] = median (NaN if the mask is empty), out[1] = count.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(1024) median_ratio_kernel(const float* __restrict__ a, int a_stride,
                                                            const float* __restrict__ b, int b_stride,
                                                            const uint8_t* __restrict__ mask, int n,
                                                            float* __restrict__ out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_rank, s_count;
  const int tid = threadIdx.x;
  if (tid == 0) s_count = 0;
  __syncthreads();
  unsigned int local = 0;
  for (int i = tid; i < n; i += blockDim.x) local += mask[i] != 0;
  atomicAdd(&s_count, local);
  __syncthreads();
  const unsigned int count = s_count;
  if (count == 0) {
    if (tid == 0) { out[0] = __int_as_float(0x7fc00000); out[1] = 0.f; }
    return;
  }
  // torch.median propagates NaN: if ANY masked ratio is NaN (0/0, inf/inf) the result is NaN, whatever its rank
  __shared__ unsigned int s_nan;
  if (tid == 0) { s_prefix = 0; s_rank = (count - 1) / 2; s_nan = 0; }
  __syncthreads();
  {
    unsigned int any_nan = 0;
    for (int i = tid; i < n; i += blockDim.x)
      if (mask[i]) any_nan |= isnan(a[size_t(i) * a_stride] / b[size_t(i) * b_stride]) ? 1u : 0u;
    if (any_nan) atomicOr(&s_nan, 1u);
  }
  __syncthreads();
  if (s_nan) {
    if (tid == 0) { out[0] = __int_as_float(0x7fc00000); out[1] = float(count); }
    return;
  }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int hi_mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < n; i += blockDim.x) {
      if (!mask[i]) continue;
      const float r = a[size_t(i) * a_stride] / b[size_t(i) * b_stride];
      const uint32_t key = isnan(r) ? 0xffffffffu : float_key(r);
      if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int rank = s_rank, acc = 0;
      int bin = 0;
      for (; bin < 256; ++bin) {
        if (acc + hist[bin] > rank) break;
        acc += hist[bin];
      }
      s_rank = rank - acc;
      s_prefix = prefix | (static_cast<unsigned int>(bin) << shift);
    }
    __syncthreads();
  }
  if (tid == 0) {
    out[0] = s_prefix == 0xffffffffu ? __int_as_float(0x7fc00000) : key_float(s_prefix);
    out[1] = float(count);
  }
}

// ------------------------------------------------------------------------------------------------
// Batched Umeyama similarity (Umeyama 1991): for each problem b find (s, R, t) minimising
// sum_i w_i | ref_i - (s R qry_i + t) |^2 over the valid correspondences. Stands in for pytheia's
// OptimizeAlignmentSim3(points_qry, points_ref, perform_optimization=False)
// (utils/reconstruction_alignment.py:89-105): maps QUERY-chunk points onto REFERENCE-chunk points.
// One block per problem, fp64 accumulation; R = U diag(1,1,det(UV^T)) V^T of the 3x3 covariance from a
// Jacobi eigen-solve. Output per problem: 4x4 row-major [sR | t; 0 0 0 1] + scale + #valid.
// ------------------------------------------------------------------------------------------------
__device__ void jacobi_eig3_d(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
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
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
  const int tid = threadIdx.x;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double tot = 0;
  for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
  return tot;
}

__global__ void __launch_bounds__(256) umeyama_kernel(const float* __restrict__ qry, const float* __restrict__ ref,
                                                      const uint8_t* __restrict__ valid, int n_per,
                                                      float* __restrict__ out) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  const float* q = qry + size_t(b) * n_per * 3;
  const float* r = ref + size_t(b) * n_per * 3;
  const uint8_t* v = valid ? valid + size_t(b) * n_per : nullptr;
  double cnt = 0, mq[3] = {0, 0, 0}, mr[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < n_per; i += blockDim.x) {
    if (v && !v[i]) continue;
    cnt += 1;
    for (int c = 0; c < 3; ++c) { mq[c] += q[i * 3 + c]; mr[c] += r[i * 3 + c]; }
  }
  cnt = block_sum(cnt, red);
  for (int c = 0; c < 3; ++c) { mq[c] = block_sum(mq[c], red); mr[c] = block_sum(mr[c], red); }
  float* o = out + size_t(b) * 18;
  if (cnt < 3) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? 1.f : 0.f;
      o[16] = 1.f; o[17] = float(cnt);
    }
    return;
  }
  for (int c = 0; c < 3; ++c) { mq[c] /= cnt; mr[c] /= cnt; }
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, varq = 0;  // S = sum (r - mr)(q - mq)^T
  for (int i = threadIdx.x; i < n_per; i += blockDim.x) {
    if (v && !v[i]) continue;
    double dq[3], dr[3];
    for (int c = 0; c < 3; ++c) { dq[c] = q[i * 3 + c] - mq[c]; dr[c] = r[i * 3 + c] - mr[c]; }
    varq += dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) S[a][c] += dr[a] * dq[c];
  }
  varq = block_sum(varq, red);
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) S[a][c] = block_sum(S[a][c], red);
  if (threadIdx.x != 0) return;
  // SVD of S via eigen-decomposition of S^T S = V diag(sig^2) V^T; u_i = S v_i / sig_i
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = S[0][i] * S[0][j] + S[1][i] * S[1][j] + S[2][i] * S[2][j];
  jacobi_eig3_d(A, V);
  int ord[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (A[ord[j]][ord[j]] > A[ord[i]][ord[i]]) { const int tmp = ord[i]; ord[i] = ord[j]; ord[j] = tmp; }
  double v1[3], v2[3], u1[3], u2[3];
  for (int k = 0; k < 3; ++k) { v1[k] = V[k][ord[0]]; v2[k] = V[k][ord[1]]; }
  double n1 = 0, n2 = 0;
  for (int a = 0; a < 3; ++a) {
    u1[a] = S[a][0] * v1[0] + S[a][1] * v1[1] + S[a][2] * v1[2];
    u2[a] = S[a][0] * v2[0] + S[a][1] * v2[1] + S[a][2] * v2[2];
    n1 += u1[a] * u1[a];
  }
  n1 = sqrt(n1);
  for (int a = 0; a < 3; ++a) u1[a] /= n1;
  const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  for (int a = 0; a < 3; ++a) { u2[a] -= d12 * u1[a]; n2 += u2[a] * u2[a]; }
  n2 = sqrt(n2);
  for (int a = 0; a < 3; ++a) u2[a] /= n2;
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
  // R = [u1 u2 u1xu2][v1 v2 v1xv2]^T (proper rotation; equals U diag(1,1,det) V^T).
  double R[3][3];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) R[a][c] = u1[a] * v1[c] + u2[a] * v2[c] + u3[a] * v3[c];
  // trace(D Sigma) = sig1 + sig2 + d*sig3 with d*sig3 = u3^T S v3 (signed)
  double s3 = 0;
  for (int a = 0; a < 3; ++a) s3 += u3[a] * (S[a][0] * v3[0] + S[a][1] * v3[1] + S[a][2] * v3[2]);
  const double sig1 = sqrt(fmax(A[ord[0]][ord[0]], 0.0)), sig2 = sqrt(fmax(A[ord[1]][ord[1]], 0.0));
  const double scale = varq > 0 ? (sig1 + sig2 + s3) / varq : 1.0;
  for (int a = 0; a < 3; ++a) {
    double t = mr[a];
    for (int c = 0; c < 3; ++c) {
      o[a * 4 + c] = float(scale * R[a][c]);
      t -= scale * R[a][c] * mq[c];
    }
    o[a * 4 + 3] = float(t);
  }
  o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  o[16] = float(scale);
  o[17] = float(cnt);
}


// ------------------------------------------------------------------------------------------------
// Batched focal / z-shift recovery: per frame,  min_{shift} | f(shift) * xy / (z + shift) - uv |
// with the focal eliminated in closed form — what utils/geometry_numpy.py:79-96 hands to
// scipy.optimize.least_squares(x0=0, ftol=1e-3, method='lm'). The reference's RESULT is defined by that solver's early
// stop (on degenerate geometry it returns after two evaluations), so this is a port of the algorithm scipy executes for
// one parameter, not a generic optimiser: scipy's 2-point finite-difference Jacobian (step sqrt(eps)*sign(x)*max(1,|x|))
// driving MINPACK lmder / lmpar / qrsolv with internal scaling (mode 1), factor 100, xtol = gtol = 1e-8, 100 evaluations
// at most — oracle/lm_oracle.py is the same code in numpy and is pinned against scipy itself. For n = 1 every matrix is a
// scalar (R = -|J|, Q^T f = -(J.f)/|J|); the only data-parallel work is the three reductions per function evaluation.
// One block per frame, fp64, all threads walk the (uniform) control flow. out[n] = (shift, focal, nfev, #valid).
// ------------------------------------------------------------------------------------------------
struct LmData {
  const float* pts;       // [S][3] of this frame
  const uint8_t* valid;   // [S]
  const float* uv;        // [S][2]
  int S;
};

// focal that minimises the residual at `shift`: sum(xy_proj . uv) / sum(|xy_proj|^2)
__device__ double lm_focal(const LmData& d, double shift, double* red) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < d.S; i += blockDim.x) {
    if (!d.valid[i]) continue;
    const double inv = 1.0 / (double(d.pts[i * 3 + 2]) + shift);
    const double px = double(d.pts[i * 3 + 0]) * inv, py = double(d.pts[i * 3 + 1]) * inv;
    a += px * double(d.uv[i * 2 + 0]) + py * double(d.uv[i * 2 + 1]);
    b += px * px + py * py;
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  return a / b;
}
// |fvec|^2 at `shift`
__device__ double lm_sqnorm(const LmData& d, double shift, double f, double* red) {
  double q = 0.0;
  for (int i = threadIdx.x; i < d.S; i += blockDim.x) {
    if (!d.valid[i]) continue;
    const double inv = 1.0 / (double(d.pts[i * 3 + 2]) + shift);
    const double rx = f * double(d.pts[i * 3 + 0]) * inv - double(d.uv[i * 2 + 0]);
    const double ry = f * double(d.pts[i * 3 + 1]) * inv - double(d.uv[i * 2 + 1]);
    q += rx * rx + ry * ry;
  }
  return block_sum(q, red);
}
// forward-difference Jacobian column J = (fvec(x1) - fvec(x0)) / dx: returns |J|^2 and J . fvec(x0)
__device__ void lm_jacobian(const LmData& d, double x0, double f0, double x1, double f1, double dx, double* red,
                            double& jj, double& jf) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < d.S; i += blockDim.x) {
    if (!d.valid[i]) continue;
    const double z = double(d.pts[i * 3 + 2]), X = double(d.pts[i * 3 + 0]), Y = double(d.pts[i * 3 + 1]);
    const double u = double(d.uv[i * 2 + 0]), v = double(d.uv[i * 2 + 1]);
    const double i0 = 1.0 / (z + x0), i1 = 1.0 / (z + x1);
    const double r0x = f0 * X * i0 - u, r0y = f0 * Y * i0 - v;
    const double jx = ((f1 * X * i1 - u) - r0x) / dx, jy = ((f1 * Y * i1 - v) - r0y) / dx;
    a += jx * jx + jy * jy;
    b += jx * r0x + jy * r0y;
  }
  jj = block_sum(a, red);
  jf = block_sum(b, red);
}

__device__ __forceinline__ void lm_qrsolv1(double r, double diag_in, double qtb, double& x, double& sdiag) {
  double rn = r, wa = qtb;
  if (diag_in != 0.0) {
    double sn, cs;
    if (fabs(r) < fabs(diag_in)) {
      const double cotan = r / diag_in;
      sn = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
      cs = sn * cotan;
    } else {
      const double tn = diag_in / r;
      cs = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
      sn = cs * tn;
    }
    rn = cs * r + sn * diag_in;
    wa = cs * qtb;
  }
  sdiag = rn;
  x = rn != 0.0 ? wa / rn : 0.0;
}

__device__ void lm_lmpar1(double r, double dg, double qtb, double delta, double& par, double& x) {
  constexpr double kDwarf = 2.2250738585072014e-308;
  x = r != 0.0 ? qtb / r : 0.0;
  double dxnorm = fabs(dg * x);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    par = 0.0;
    return;
  }
  double parl = 0.0;
  if (r != 0.0) {
    const double w = dg * (dg * x / dxnorm) / r;
    parl = ((fp / delta) / fabs(w)) / fabs(w);
  }
  const double gnorm = fabs(r * qtb / dg);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = kDwarf / fmin(delta, 0.1);
  par = fmin(fmax(par, parl), paru);
  if (par == 0.0) par = gnorm / dxnorm;
  for (int it = 1;; ++it) {
    if (par == 0.0) par = fmax(kDwarf, 0.001 * paru);
    double sdiag;
    lm_qrsolv1(r, sqrt(par) * dg, qtb, x, sdiag);
    dxnorm = fabs(dg * x);
    const double temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || it == 10) break;
    const double w = dg * (dg * x / dxnorm) / sdiag;
    const double parc = ((fp / delta) / fabs(w)) / fabs(w);
    if (fp > 0.0) parl = fmax(parl, par);
    if (fp < 0.0) paru = fmin(paru, par);
    par = fmax(parl, par + parc);
  }
}

__global__ void __launch_bounds__(256) focal_shift_lm_kernel(const float* __restrict__ pts, const uint8_t* __restrict__ valid,
                                                             const float* __restrict__ uv, int S, float ftol_f,
                                                             float* __restrict__ out) {
  __shared__ double red[8];
  constexpr double kEps = 2.220446049250313e-16, kFactor = 100.0, kXtol = 1e-8, kGtol = 1e-8;
  constexpr int kMaxFev = 100;
  const double ftol = double(ftol_f);
  const int n = blockIdx.x;
  LmData d{pts + size_t(n) * S * 3, valid + size_t(n) * S, uv, S};
  double cnt = 0.0;
  for (int i = threadIdx.x; i < S; i += blockDim.x) cnt += d.valid[i] ? 1.0 : 0.0;
  cnt = block_sum(cnt, red);
  if (cnt < 2.0) {   // reference: focal 1, shift 0 when fewer than two valid samples (geometry_torch.py:150-154)
    if (threadIdx.x == 0) { out[n * 4 + 0] = 0.f; out[n * 4 + 1] = 1.f; out[n * 4 + 2] = 0.f; out[n * 4 + 3] = float(cnt); }
    return;
  }
  double x = 0.0;
  double fx = lm_focal(d, x, red);
  double fnorm = sqrt(lm_sqnorm(d, x, fx, red));
  int nfev = 1, iter = 1, info = 0;
  double par = 0.0, diag = 0.0, delta = 0.0, xnorm = 0.0;
  while (info == 0) {
    const double h = sqrt(kEps) * (x >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(x));
    const double x1 = x + h;
    const double dx = x1 - x;
    const double f1 = lm_focal(d, x1, red);
    double jj, jf;
    lm_jacobian(d, x, fx, x1, f1, dx, red, jj, jf);
    const double jnorm = sqrt(jj);
    const double r = -jnorm;
    const double qtf = jnorm != 0.0 ? -jf / jnorm : 0.0;
    if (iter == 1) {
      diag = jnorm != 0.0 ? jnorm : 1.0;
      xnorm = fabs(diag * x);
      delta = kFactor * xnorm;
      if (delta == 0.0) delta = kFactor;
    }
    double gnorm = 0.0;
    if (fnorm != 0.0 && jnorm != 0.0) gnorm = fabs(r * (qtf / fnorm) / jnorm);
    if (gnorm <= kGtol) {
      info = 4;
      break;
    }
    diag = fmax(diag, jnorm);
    for (;;) {
      double p;
      lm_lmpar1(r, diag, qtf, delta, par, p);
      p = -p;
      const double x_new = x + p;
      const double pnorm = fabs(diag * p);
      if (iter == 1) delta = fmin(delta, pnorm);
      const double f_new = lm_focal(d, x_new, red);
      const double fnorm1 = sqrt(lm_sqnorm(d, x_new, f_new, red));
      ++nfev;
      double actred = -1.0;
      if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
      const double temp1 = fabs(r * p) / fnorm;
      const double temp2 = sqrt(par) * pnorm / fnorm;
      const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
      const double dirder = -(temp1 * temp1 + temp2 * temp2);
      const double ratio = prered != 0.0 ? actred / prered : 0.0;
      if (ratio <= 0.25) {
        double temp = actred >= 0.0 ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        delta = temp * fmin(delta, pnorm / 0.1);
        par /= temp;
      } else if (par == 0.0 || ratio >= 0.75) {
        delta = pnorm / 0.5;
        par *= 0.5;
      }
      if (ratio >= 1e-4) {
        x = x_new;
        fx = f_new;
        fnorm = fnorm1;
        xnorm = fabs(diag * x);
        ++iter;
      }
      const bool ftest = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
      if (ftest) info = 1;
      if (delta <= kXtol * xnorm) info = 2;
      if (ftest && info == 2) info = 3;
      if (info != 0) break;
      if (nfev >= kMaxFev) info = 5;
      if (fabs(actred) <= kEps && prered <= kEps && 0.5 * ratio <= 1.0) info = 6;
      if (delta <= kEps * xnorm) info = 7;
      if (gnorm <= kEps) info = 8;
      if (info != 0 || ratio >= 1e-4) break;
    }
  }
  // the reference re-evaluates the focal at the fp32-rounded shift (geometry_numpy.py:91-94)
  const float shift32 = float(x);
  const double focal = lm_focal(d, double(shift32), red);
  if (threadIdx.x == 0) {
    out[n * 4 + 0] = shift32;
    out[n * 4 + 1] = float(focal);
    out[n * 4 + 2] = float(nfev);
    out[n * 4 + 3] = float(cnt);
  }
}


// ------------------------------------------------------------------------------------------------
// Stage-2 front end: the reprojection of every keypoint of a chunk into its neighbouring views
// and the image-bounds test that ChunkPTRecon.create_recon_from_chunk runs as Python loops
// (utils/chunk_reconstruction.py:158-185 with _project_3d_points_to_camera :494-509): for a (source frame, target frame)
// pair, X_cam = W2T_target [X 1], p = X_cam / X_cam.z, (u, v) = (K_target p).xy (or p.xy without intrinsics), kept when
// 0 <= u < W and 0 <= v < H. fp64 like the reference's numpy (world-to-target inverses are computed by the caller exactly
// as the reference does, numpy.linalg.inv in fp32). One block per pair; the in-bounds keypoints of the pair are written
// COMPACTED AND IN KEYPOINT ORDER (ballot + prefix), so that the observation list keeps the reference's AddObservation
// order. A second launch scans the per-pair counts and scatters the rows into one flat list.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) project_observations_kernel(const float* __restrict__ points, const double* __restrict__ w2t,
                                                                   const double* __restrict__ kmat, const int* __restrict__ pairs,
                                                                   int K, double W, double H, double* __restrict__ row_uv,
                                                                   int* __restrict__ row_kp, int* __restrict__ row_count) {
  __shared__ int warp_cnt[4];
  __shared__ int base_s;
  const int r = blockIdx.x;
  const int src = pairs[2 * r], dst = pairs[2 * r + 1];
  const double* T = w2t + size_t(dst) * 16;
  const double* Km = kmat ? kmat + size_t(dst) * 9 : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += blockDim.x) {
    const int k = k0 + threadIdx.x;
    bool inb = false;
    double u = 0.0, v = 0.0;
    if (k < K) {
      const float* P = points + (size_t(src) * K + k) * 3;
      const double x = double(P[0]), y = double(P[1]), z = double(P[2]);
      const double cx = T[0] * x + T[1] * y + T[2] * z + T[3];
      const double cy = T[4] * x + T[5] * y + T[6] * z + T[7];
      const double cz = T[8] * x + T[9] * y + T[10] * z + T[11];
      const double px = cx / cz, py = cy / cz, pz = cz / cz;
      if (Km) {
        u = Km[0] * px + Km[1] * py + Km[2] * pz;
        v = Km[3] * px + Km[4] * py + Km[5] * pz;
      } else {
        u = px;
        v = py;
      }
      inb = (0.0 <= u) && (u < W) && (0.0 <= v) && (v < H);   // NaN / inf (cz == 0) compare false, as in the reference
    }
    const unsigned ball = __ballot_sync(0xffffffffu, inb);
    if (lane == 0) warp_cnt[warp] = __popc(ball);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (inb) {
      const int pos = off + __popc(ball & ((1u << lane) - 1u));
      row_uv[(size_t(r) * K + pos) * 2 + 0] = u;
      row_uv[(size_t(r) * K + pos) * 2 + 1] = v;
      row_kp[size_t(r) * K + pos] = k;
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s += warp_cnt[0] + warp_cnt[1] + warp_cnt[2] + warp_cnt[3];
    __syncthreads();
  }
  if (threadIdx.x == 0) row_count[r] = base_s;
}

// exclusive scan of the per-pair counts (single block; R is a few thousand) -> row_offset[0..R]
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int* __restrict__ cnt, int R, int* __restrict__ off) {
  __shared__ int part[1024];
  const int per = (R + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(R, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += cnt[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < blockDim.x; ++i) { const int t = part[i]; part[i] = run; run += t; }
    off[R] = run;
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { off[i] = run; run += cnt[i]; }
}

__global__ void __launch_bounds__(128) scatter_observations_kernel(const double* __restrict__ row_uv, const int* __restrict__ row_kp,
                                                                   const int* __restrict__ row_count, const int* __restrict__ row_offset,
                                                                   const int* __restrict__ pairs, int K, int capacity,
                                                                   int* __restrict__ obs_src, int* __restrict__ obs_kp,
                                                                   int* __restrict__ obs_view, double* __restrict__ obs_xy) {
  const int r = blockIdx.x;
  const int n = row_count[r], o0 = row_offset[r];
  const int src = pairs[2 * r], dst = pairs[2 * r + 1];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int o = o0 + i;
    if (o >= capacity) return;
    obs_src[o] = src;
    obs_kp[o] = row_kp[size_t(r) * K + i];
    obs_view[o] = dst;
    obs_xy[size_t(o) * 2 + 0] = row_uv[(size_t(r) * K + i) * 2 + 0];
    obs_xy[size_t(o) * 2 + 1] = row_uv[(size_t(r) * K + i) * 2 + 1];
  }
}

}  // namespace pi3b

using namespace pi3b;

extern "C" int pi3b_keypoint_gather(void* stream, const float* points, const float* local_points, const float* conf,
                                    const uint8_t* mask, const float* images, const float* keypoints, float* out_points,
                                    float* out_local, float* out_conf, uint8_t* out_mask, uint8_t* out_colors, int N,
                                    int K, int H, int W, int grid_math) {
  PI3B_REQUIRE(points && local_points && conf && mask && keypoints && out_points && out_local && out_conf && out_mask,
               "pi3b_keypoint_gather: null pointer");
  PI3B_REQUIRE(!images || out_colors, "pi3b_keypoint_gather: images given without out_colors");
  PI3B_REQUIRE(N > 0 && K > 0 && H > 1 && W > 1, "pi3b_keypoint_gather: bad shape");
  PI3B_REQUIRE(grid_math == 0 || grid_math == 1, "pi3b_keypoint_gather: grid_math must be 0 (CPU) or 1 (CUDA)");
  const int total = N * K;
  keypoint_gather_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      points, local_points, conf, mask, images, keypoints, out_points, out_local, out_conf, out_mask, out_colors, N, K,
      H, W, grid_math);
  return check_launch("keypoint_gather launch");
}

extern "C" int pi3b_sample_grid(void* stream, const float* local_points, const float* conf, const int* ys,
                                const int* xs, float* out_pts, uint8_t* out_valid, int N, int H, int W, int hs, int ws,
                                float conf_thr) {
  PI3B_REQUIRE(local_points && conf && ys && xs && out_pts && out_valid, "pi3b_sample_grid: null pointer");
  PI3B_REQUIRE(N > 0 && hs > 0 && ws > 0, "pi3b_sample_grid: bad shape");
  const int total = N * hs * ws;
  sample_grid_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      local_points, conf, ys, xs, out_pts, out_valid, N, H, W, hs, ws, conf_thr);
  return check_launch("sample_grid launch");
}

extern "C" int pi3b_median_ratio(void* stream, const float* a, int a_stride, const float* b, int b_stride,
                                 const uint8_t* mask, int n, float* out2) {
  PI3B_REQUIRE(a && b && mask && out2 && n > 0 && a_stride > 0 && b_stride > 0, "pi3b_median_ratio: bad args");
  median_ratio_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a, a_stride, b, b_stride, mask, n, out2);
  return check_launch("median_ratio launch");
}

extern "C" int pi3b_umeyama_sim3_batched(void* stream, const float* qry, const float* ref, const uint8_t* valid,
                                         int num_problems, int n_per, float* out18) {
  PI3B_REQUIRE(qry && ref && out18 && num_problems > 0 && n_per > 0, "pi3b_umeyama_sim3_batched: bad args");
  umeyama_kernel<<<num_problems, 256, 0, static_cast<cudaStream_t>(stream)>>>(qry, ref, valid, n_per, out18);
  return check_launch("umeyama launch");
}

extern "C" int pi3b_focal_shift_lm(void* stream, const float* pts, const uint8_t* valid, const float* uv, int N, int S,
                                   float ftol, float* out4) {
  PI3B_REQUIRE(pts && valid && uv && out4, "pi3b_focal_shift_lm: null pointer");
  PI3B_REQUIRE(N > 0 && S > 0 && ftol > 0.f, "pi3b_focal_shift_lm: bad arguments");
  focal_shift_lm_kernel<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(pts, valid, uv, S, ftol, out4);
  return check_launch("focal_shift_lm launch");
}

extern "C" int pi3b_project_observations(void* stream, const float* points, const double* w2t, const double* kmat,
                                         const int* pairs, int R, int K, double width, double height, double* row_uv,
                                         int* row_kp, int* row_count, int* row_offset, int* obs_src, int* obs_kp,
                                         int* obs_view, double* obs_xy, int capacity) {
  PI3B_REQUIRE(points && w2t && pairs && row_uv && row_kp && row_count && row_offset && obs_src && obs_kp && obs_view && obs_xy,
               "pi3b_project_observations: null pointer");
  PI3B_REQUIRE(R > 0 && K > 0 && capacity > 0, "pi3b_project_observations: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  project_observations_kernel<<<R, 128, 0, st>>>(points, w2t, kmat, pairs, K, width, height, row_uv, row_kp, row_count);
  if (int rc = check_launch("project_observations launch")) return rc;
  scan_counts_kernel<<<1, 1024, 0, st>>>(row_count, R, row_offset);
  if (int rc = check_launch("scan_counts launch")) return rc;
  scatter_observations_kernel<<<R, 128, 0, st>>>(row_uv, row_kp, row_count, row_offset, pairs, K, capacity, obs_src, obs_kp,
                                                 obs_view, obs_xy);
  return check_launch("scatter_observations launch");
}
