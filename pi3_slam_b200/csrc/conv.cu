// Dense-prediction kernels of the MoGe metric-depth pass (reference moge/model/v2.py:128-290,
// moge/model/modules.py:18-254). Feature maps are channels-last fp32 matrices [H*W, C]; every convolution is a
// tcgen05 GEMM (gemm.cu) over a bf16 operand produced here:
//   1x1 conv           -> GEMM on the map itself
//   3x3 replicate conv -> im2col3x3_kernel (GroupNorm + ReLU of the block fused into the gather) -> GEMM
//   ConvTranspose 2x2  -> GEMM to [H*W, 4*Cout] (column block = output sub-pixel); the 3x3 conv that follows
//                         gathers straight out of that blocked layout, so the pixel shuffle is never materialised
// plus the HBM-bound pieces around them: GroupNorm statistics, the uv (view-plane) 1x1 conv, separable
// resampling with host-built tap tables (antialiased input resize, bilinear output resize), output remap and the
// final depth assembly. None of these reshapes its work into a GEMM.
#include "host_common.h"

#include <cuda_bf16.h>
#include <math.h>

namespace pi3b {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics (modules.py:48-56: GroupNorm(1, C) "layer_norm" and GroupNorm(C/32, C)): per group
// sum and sum of squares over cpg channels x all pixels. Each thread owns one 4-channel column quad (so its
// group is fixed), accumulates <= 32 rows in fp32, then the block folds into fp64 shared bins and issues one
// fp64 atomic per group: the result is order-independent to ~1e-16.
// ------------------------------------------------------------------------------------------------
constexpr int kGnRowsPerThread = 32;

__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows,
                                                              int C, int cpg, double* __restrict__ stats) {
  __shared__ double bins[64];
  const int cq = C >> 2;
  const int rows_per_iter = 256 / cq;
  const int G = C / cpg;
  if (threadIdx.x < 2 * G) bins[threadIdx.x] = 0.0;
  __syncthreads();
  const int col = threadIdx.x % cq, sub = threadIdx.x / cq;
  if (sub < rows_per_iter) {
    const int64_t row0 = int64_t(blockIdx.x) * rows_per_iter * kGnRowsPerThread + sub;
    float s = 0.f, q = 0.f;
#pragma unroll 4
    for (int i = 0; i < kGnRowsPerThread; ++i) {
      const int64_t r = row0 + int64_t(i) * rows_per_iter;
      if (r < rows) {
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + col * 4);
        s += (v.x + v.y) + (v.z + v.w);
        q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    }
    const int g = (col * 4) / cpg;
    atomicAdd(&bins[2 * g], double(s));
    atomicAdd(&bins[2 * g + 1], double(q));
  }
  __syncthreads();
  if (threadIdx.x < 2 * G) atomicAdd(&stats[threadIdx.x], bins[threadIdx.x]);
}

// stats[g] = (sum, sumsq) -> mean_rstd[g] = (mean, 1/sqrt(biased var + eps))
__global__ void groupnorm_finalize_kernel(const double* __restrict__ stats, float2* __restrict__ mean_rstd, int G,
                                          double count, float eps) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const double mean = stats[2 * g] / count;
  double var = stats[2 * g + 1] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[g] = make_float2(float(mean), float(1.0 / sqrt(var + double(eps))));
}

// ------------------------------------------------------------------------------------------------
// im2col for a 3x3 / stride 1 / replicate-padded convolution on a channels-last fp32 map, with the block's
// normalisation + ReLU applied on the fly (a pointwise op commutes with replicate padding):
//   out[y*W + x][t*C + c] = act(norm(src(clamp(y + t/3 - 1), clamp(x + t%3 - 1), c)))          t = 0..8
// `blocked` != 0: the source is the [H/2 * W/2, 4*C] output of a 2x2/stride-2 transposed-conv GEMM whose column
// block (dy*2 + dx) holds output sub-pixel (2Y+dy, 2X+dx). One thread = one (pixel, tap, 8 channels): two
// 16-byte loads, one 16-byte store; the 9-fold re-read of the source is served by L2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float* __restrict__ x, int64_t ldx, int H, int W, int C,
                                                        int blocked, const float2* __restrict__ mean_rstd, int cpg,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int relu, __nv_bfloat16* __restrict__ out, int64_t ldo,
                                                        int64_t total) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c8n = C >> 3;
  const int c8 = int(gid % c8n);
  const int64_t pt = gid / c8n;
  const int t = int(pt % 9);
  const int64_t p = pt / 9;
  const int y = int(p / W), xx = int(p % W);
  int sy = y + t / 3 - 1, sx = xx + t % 3 - 1;
  sy = sy < 0 ? 0 : (sy >= H ? H - 1 : sy);
  sx = sx < 0 ? 0 : (sx >= W ? W - 1 : sx);
  const int c0 = c8 * 8;
  const float* src;
  if (blocked) {
    src = x + (int64_t(sy >> 1) * (W >> 1) + (sx >> 1)) * ldx + ((sy & 1) * 2 + (sx & 1)) * C + c0;
  } else {
    src = x + (int64_t(sy) * W + sx) * ldx + c0;
  }
  float4 a = *reinterpret_cast<const float4*>(src);
  float4 b = *reinterpret_cast<const float4*>(src + 4);
  if (mean_rstd) {
    const float2 mr = mean_rstd[c0 / cpg];
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    a.x = (a.x - mr.x) * mr.y * g0.x + b0.x; a.y = (a.y - mr.x) * mr.y * g0.y + b0.y;
    a.z = (a.z - mr.x) * mr.y * g0.z + b0.z; a.w = (a.w - mr.x) * mr.y * g0.w + b0.w;
    b.x = (b.x - mr.x) * mr.y * g1.x + b1.x; b.y = (b.y - mr.x) * mr.y * g1.y + b1.y;
    b.z = (b.z - mr.x) * mr.y * g1.z + b1.z; b.w = (b.w - mr.x) * mr.y * g1.w + b1.w;
  }
  if (relu) {
    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
  }
  uint4 o = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
  *reinterpret_cast<uint4*>(out + p * ldo + int64_t(t) * C + c0) = o;
}

// ------------------------------------------------------------------------------------------------
// 1x1 conv on the 2-channel view-plane uv map (v2.py:141-147 concatenates uv to every pyramid level; the
// ConvStack input block is a 1x1 conv, modules.py:208-211): x[p, c] (+)= w[c][0]*u[p%W] + w[c][1]*v[p/W] + b[c].
// u_tab / v_tab are the reference's linspace values (geometry_torch.py:40-52), per-shape constants.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) uv_affine_kernel(float* __restrict__ x, int64_t ldx, int H, int W, int C,
                                                        const float* __restrict__ w, int ldw,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ u_tab,
                                                        const float* __restrict__ v_tab, int accumulate,
                                                        int64_t total) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cq = C >> 2;
  const int c0 = int(gid % cq) * 4;
  const int64_t p = gid / cq;
  const float u = u_tab[p % W], v = v_tab[p / W];
  float4* dst = reinterpret_cast<float4*>(x + p * ldx + c0);
  float4 o = accumulate ? *dst : make_float4(0.f, 0.f, 0.f, 0.f);
  float r[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float wu = __ldg(w + int64_t(c0 + i) * ldw), wv = __ldg(w + int64_t(c0 + i) * ldw + 1);
    r[i] += wu * u + wv * v + (bias ? __ldg(bias + c0 + i) : 0.f);
  }
  *dst = make_float4(r[0], r[1], r[2], r[3]);
}

// ------------------------------------------------------------------------------------------------
// Separable resampling with host-built tap tables (ATen's upsample_bilinear2d and _upsample_bilinear2d_aa
// weights): out(c, y, x) = sum_j wy[y][j] * sum_i wx[x][i] * in(c, y0[y] + j, x0[x] + i), horizontal sum first
// like ATen. Element (c, y, x) of in / out sits at c*sc + y*sy + x*sx (planar images and channels-last maps).
// Used for encoder input F.interpolate(antialias=True) (modules.py:121) and the heads' bilinear resize (v2.py:157).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resample2d_kernel(const float* __restrict__ in, int64_t isc, int64_t isy,
                                                         int64_t isx, int Hi, int Wi, float* __restrict__ out,
                                                         int64_t osc, int64_t osy, int64_t osx, int C, int Ho, int Wo,
                                                         const int* __restrict__ y0, const float* __restrict__ wy,
                                                         int ty, const int* __restrict__ x0,
                                                         const float* __restrict__ wx, int tx, int64_t total) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  // x fastest for planar output, channel fastest for channels-last output
  int c, y, x;
  if (osx == 1) {
    x = int(gid % Wo); y = int((gid / Wo) % Ho); c = int(gid / (int64_t(Wo) * Ho));
  } else {
    c = int(gid % C); x = int((gid / C) % Wo); y = int(gid / (int64_t(C) * Wo));
  }
  const int ys = y0[y], xs = x0[x];
  float acc = 0.f;
  for (int j = 0; j < ty; ++j) {
    const float wj = wy[y * ty + j];
    if (wj == 0.f) continue;
    const int yy = min(ys + j, Hi - 1);
    const float* row = in + c * isc + yy * isy;
    float h = 0.f;
    for (int i = 0; i < tx; ++i) {
      const float wi = wx[x * tx + i];
      if (wi != 0.f) h += wi * row[min(xs + i, Wi - 1) * isx];
    }
    acc += wj * h;
  }
  out[c * osc + y * osy + x * osx] = acc;
}

// ------------------------------------------------------------------------------------------------
// Output remap (v2.py:112-126, 160-169): points (H*W, 3) in place — 0 linear, 1 sinh, 2 exp (z = e^z, xy *= z),
// 3 sinh_exp; mask_prob = sigmoid(mask_logit).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) moge_remap_kernel(float* __restrict__ pts, const float* __restrict__ mask_logit,
                                                         float* __restrict__ mask_prob, int mode, int64_t n) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float x = pts[p * 3], y = pts[p * 3 + 1], z = pts[p * 3 + 2];
  if (mode == 1) { x = sinhf(x); y = sinhf(y); z = sinhf(z); }
  else if (mode == 2) { z = expf(z); x *= z; y *= z; }
  else if (mode == 3) { x = sinhf(x); y = sinhf(y); z = expf(z); }
  pts[p * 3] = x; pts[p * 3 + 1] = y; pts[p * 3 + 2] = z;
  if (mask_logit) mask_prob[p] = 1.0f / (1.0f + expf(-mask_logit[p]));
}

// ------------------------------------------------------------------------------------------------
// Depth assembly of MoGeModel.infer (v2.py:236-275): z' = z + shift; mask = prob > 0.5 && z' > 0;
// depth = mask ? z' * e^{log_scale} : +inf. log_scale: device pointer to the scale head's raw output (or NULL).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) moge_depth_kernel(const float* __restrict__ pts, const float* __restrict__ mask_prob,
                                                         float shift, const float* __restrict__ log_scale,
                                                         int apply_mask, float* __restrict__ depth,
                                                         uint8_t* __restrict__ mask_out, int64_t n) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float z = pts[p * 3 + 2] + shift;
  bool m = true;
  if (mask_prob) m = mask_prob[p] > 0.5f && z > 0.f;
  float d = z;
  if (log_scale) d *= expf(log_scale[0]);
  if (apply_mask && mask_prob && !m) d = INFINITY;
  depth[p] = d;
  if (mask_out) mask_out[p] = m ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Pillow-exact 8-bit BILINEAR resize + ToTensor (reference loader: transforms.Resize(target)(PIL image) -> ToTensor,
// datasets/image_datasets.py:193-199): in uint8 [N, Hi, Wi, 3] -> out fp32 [N, 3, Ho, Wo] in [0, 1].
// Integer arithmetic of libImaging/Resample.c: horizontal pass to uint8 (clip8((2^21 + sum px * k) >> 22)), then the
// vertical pass on those uint8 values, then / 255.0f (IEEE division, as torch's). The horizontal intermediate is
// recomputed per vertical tap instead of materialised (<= 5 x 5 taps for the path's ~1.5x down-scale).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clip8(int v) {
  v >>= 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) resize_u8_pil_kernel(const uint8_t* __restrict__ in, int Hi, int Wi,
                                                            float* __restrict__ out, int Ho, int Wo,
                                                            const int* __restrict__ ymin, const int* __restrict__ ycnt,
                                                            const int* __restrict__ ky, int ksy,
                                                            const int* __restrict__ xmin, const int* __restrict__ xcnt,
                                                            const int* __restrict__ kx, int ksx, int64_t total) {
  const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int x = int(gid % Wo), y = int((gid / Wo) % Ho);
  const int64_t n = gid / (int64_t(Wo) * Ho);
  const int x0 = xmin[x], nx = xcnt[x], y0 = ymin[y], ny = ycnt[y];
  const uint8_t* img = in + n * int64_t(Hi) * Wi * 3;
  int v0 = 1 << 21, v1 = 1 << 21, v2 = 1 << 21;
  for (int j = 0; j < ny; ++j) {
    const uint8_t* row = img + (int64_t(y0 + j) * Wi + x0) * 3;
    int h0 = 1 << 21, h1 = 1 << 21, h2 = 1 << 21;
    for (int i = 0; i < nx; ++i) {
      const int k = __ldg(kx + x * ksx + i);
      h0 += int(row[i * 3 + 0]) * k;
      h1 += int(row[i * 3 + 1]) * k;
      h2 += int(row[i * 3 + 2]) * k;
    }
    const int k = __ldg(ky + y * ksy + j);
    v0 += clip8(h0) * k;
    v1 += clip8(h1) * k;
    v2 += clip8(h2) * k;
  }
  const int64_t plane = int64_t(Ho) * Wo;
  float* o = out + n * 3 * plane + int64_t(y) * Wo + x;
  o[0] = float(clip8(v0)) / 255.0f;
  o[plane] = float(clip8(v1)) / 255.0f;
  o[2 * plane] = float(clip8(v2)) / 255.0f;
}

}  // namespace pi3b

using namespace pi3b;

static inline unsigned blocks64(int64_t total, int per) { return unsigned((total + per - 1) / per); }

extern "C" int pi3b_groupnorm_stats(void* stream, const float* x, int64_t ldx, int64_t rows, int C, int groups,
                                    float eps, double* workspace, float* mean_rstd) {
  PI3B_REQUIRE(x && workspace && mean_rstd, "pi3b_groupnorm_stats: null pointer");
  PI3B_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= 1024 && ldx % 4 == 0 && ldx >= C, "pi3b_groupnorm_stats: bad C/ld");
  PI3B_REQUIRE(groups > 0 && groups <= 32 && C % groups == 0 && (C / groups) % 4 == 0, "pi3b_groupnorm_stats: bad groups");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * groups, st) != cudaSuccess) return check_launch("groupnorm memset");
  const int rows_per_block = (256 / (C / 4)) * kGnRowsPerThread;
  groupnorm_stats_kernel<<<blocks64(rows, rows_per_block), 256, 0, st>>>(x, ldx, rows, C, C / groups, workspace);
  int rc = check_launch("groupnorm_stats launch");
  if (rc) return rc;
  groupnorm_finalize_kernel<<<1, 32, 0, st>>>(workspace, reinterpret_cast<float2*>(mean_rstd), groups,
                                              double(rows) * double(C / groups), eps);
  return check_launch("groupnorm_finalize launch");
}

extern "C" int pi3b_im2col3x3(void* stream, const float* x, int64_t ldx, int H, int W, int C, int blocked,
                              const float* mean_rstd, int groups, const float* gamma, const float* beta, int relu,
                              void* out, int64_t ldo) {
  PI3B_REQUIRE(x && out && H > 0 && W > 0 && C > 0 && C % 8 == 0, "pi3b_im2col3x3: bad shape (C must be a multiple of 8)");
  PI3B_REQUIRE(ldo >= 9 * int64_t(C) && ldo % 8 == 0 && ldx % 4 == 0, "pi3b_im2col3x3: bad ld");
  PI3B_REQUIRE(!blocked || (H % 2 == 0 && W % 2 == 0 && ldx >= 4 * int64_t(C)), "pi3b_im2col3x3: blocked source needs even H, W and ldx >= 4C");
  PI3B_REQUIRE(blocked || ldx >= C, "pi3b_im2col3x3: ldx < C");
  PI3B_REQUIRE(!mean_rstd || (gamma && beta && groups > 0 && C % groups == 0 && (C / groups) % 8 == 0),
               "pi3b_im2col3x3: norm needs gamma/beta and 8 | C/groups");
  const int64_t total = int64_t(H) * W * 9 * (C / 8);
  im2col3x3_kernel<<<blocks64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, H, W, C, blocked, reinterpret_cast<const float2*>(mean_rstd), mean_rstd ? C / groups : C, gamma, beta, relu,
      static_cast<__nv_bfloat16*>(out), ldo, total);
  return check_launch("im2col3x3 launch");
}

extern "C" int pi3b_uv_affine(void* stream, float* x, int64_t ldx, int H, int W, int C, const float* w, int ldw,
                              const float* bias, const float* u_tab, const float* v_tab, int accumulate) {
  PI3B_REQUIRE(x && w && u_tab && v_tab && H > 0 && W > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldx >= C && ldw >= 2,
               "pi3b_uv_affine: bad args");
  const int64_t total = int64_t(H) * W * (C / 4);
  uv_affine_kernel<<<blocks64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, H, W, C, w, ldw, bias,
                                                                                        u_tab, v_tab, accumulate, total);
  return check_launch("uv_affine launch");
}

extern "C" int pi3b_resample2d(void* stream, const float* in, int64_t in_sc, int64_t in_sy, int64_t in_sx, int Hi, int Wi,
                               float* out, int64_t out_sc, int64_t out_sy, int64_t out_sx, int C, int Ho, int Wo,
                               const int* y0, const float* wy, int taps_y, const int* x0, const float* wx, int taps_x) {
  PI3B_REQUIRE(in && out && y0 && wy && x0 && wx, "pi3b_resample2d: null pointer");
  PI3B_REQUIRE(C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && taps_y > 0 && taps_x > 0, "pi3b_resample2d: bad shape");
  const int64_t total = int64_t(C) * Ho * Wo;
  resample2d_kernel<<<blocks64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, in_sc, in_sy, in_sx, Hi, Wi, out, out_sc, out_sy, out_sx, C, Ho, Wo, y0, wy, taps_y, x0, wx, taps_x, total);
  return check_launch("resample2d launch");
}

extern "C" int pi3b_moge_remap(void* stream, float* points, const float* mask_logit, float* mask_prob, int mode,
                               int64_t n) {
  PI3B_REQUIRE(points && n > 0 && mode >= 0 && mode <= 3 && (!mask_logit || mask_prob), "pi3b_moge_remap: bad args");
  moge_remap_kernel<<<blocks64(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(points, mask_logit, mask_prob, mode, n);
  return check_launch("moge_remap launch");
}

extern "C" int pi3b_moge_depth(void* stream, const float* points, const float* mask_prob, float shift,
                               const float* log_scale, int apply_mask, float* depth, uint8_t* mask_out, int64_t n) {
  PI3B_REQUIRE(points && depth && n > 0, "pi3b_moge_depth: bad args");
  moge_depth_kernel<<<blocks64(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(points, mask_prob, shift, log_scale,
                                                                                   apply_mask, depth, mask_out, n);
  return check_launch("moge_depth launch");
}

extern "C" int pi3b_resize_u8_pil(void* stream, const uint8_t* in, int N, int Hi, int Wi, float* out, int Ho, int Wo,
                                  const int* ymin, const int* ycnt, const int* ky, int ksize_y, const int* xmin,
                                  const int* xcnt, const int* kx, int ksize_x) {
  PI3B_REQUIRE(in && out && ymin && ycnt && ky && xmin && xcnt && kx, "pi3b_resize_u8_pil: null pointer");
  PI3B_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && ksize_y > 0 && ksize_x > 0, "pi3b_resize_u8_pil: bad shape");
  const int64_t total = int64_t(N) * Ho * Wo;
  resize_u8_pil_kernel<<<blocks64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, Hi, Wi, out, Ho, Wo, ymin, ycnt, ky, ksize_y, xmin, xcnt, kx, ksize_x, total);
  return check_launch("resize_u8_pil launch");
}
