/*
 * pi3b200.h — C ABI of libpi3b200.so: the B200 (sm_100a) kernels behind Pi3_SLAM's offline chunk-creation
 * hot path. This is the drop-in boundary: plain pointers and sizes, no torch types.
 *
 * Conventions (all entry points):
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and the call returns
 *     immediately. No entry point synchronises, allocates device memory or throws.
 *   - every pointer is a DEVICE pointer unless marked "host".
 *   - matrices are row-major with unit inner stride; `ld*` are row pitches in ELEMENTS.
 *   - return value: 0 = ok; -1 bad argument; -2 CUDA launch/runtime error; -3 driver (TMA descriptor) error.
 *     pi3b_last_error() returns a thread-local message for the last failing call on this thread.
 *   - thread-safe for concurrent calls on distinct streams and on distinct devices (per-device function attributes and
 *     SM counts; the caller makes the tensors' device current before the call, as the CUDA runtime requires).
 *
 * "replaces:" cites the reference interface (file:line in urbste/Pi3_SLAM) each entry point stands in for.
 */
#ifndef PI3B200_H_
#define PI3B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int pi3b_version(void);
const char* pi3b_last_error(void);

/* out[M,N] = [resid +] gamma * act(A[M,K] * W[N,K]^T + bias)      (tcgen05 tensor cores, fp32 accumulate)
 * in_dtype 0: A, W bf16 (kind::f16); 1: A, W fp32 consumed as TF32 (kind::tf32).
 * bias, gamma: fp32 [N] or NULL. resid: fp32 [.., ld_resid] or NULL, may alias out. out: fp32 (out_f32=1) or bf16.
 * act: 0 none, 1 GELU(erf), 2 ReLU. Row remap (grp_in > 0): GEMM row m is written to / reads resid from row
 * (m / grp_in) * grp_out + grp_off + m % grp_in. lda/ldw must give 16-byte row pitches.
 * replaces: every nn.Linear on the path — pi3/models/layers/attention.py:256-258 (qkv, proj),
 *   pi3/models/dinov2/layers/mlp.py:29-31 (fc1+GELU, fc2), pi3/models/dinov2/layers/layer_scale.py:26-27 and
 *   the residual adds of pi3/models/layers/block.py:333-334 (fused), pi3/models/layers/transformer_head.py:24,46,68,
 *   pi3/models/layers/camera_head.py:21-23,40-46, and the patch-embed conv pi3/models/dinov2/layers/patch_embed.py:65,75. */
int pi3b_gemm(void* stream, int in_dtype, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
              const float* bias, const float* gamma, const float* resid, int64_t ld_resid, void* out, int64_t ldo,
              int out_f32, int act, int grp_in, int grp_out, int grp_off);

/* The QKV projection with the q/k post-processing fused into the epilogue: out[M, 3*64*num_heads] bf16 =
 * (A[M,K] * W[3*64*num_heads, K]^T + bias) rounded to bf16, then on every 64-channel head of the q and k thirds
 * LayerNorm(64) (qk_norm = fp32 [4][64]: q weight, q bias, k weight, k bias; NULL to skip) and 2-D RoPE (cs_table fp32
 * [table_rows][16][2] = (cos, sin) of pos * base^(-j/16); NULL to skip; token geometry as pi3b_qknorm_rope); the v third
 * is stored as is. q_scale (1 = none): the query heads are multiplied by it before their final rounding to bf16 — pass
 * softmax_scale * log2(e) and call the attention with scale = ln 2: the long-sequence kernel then feeds its accumulator
 * straight to exp2 ("zero-reference" softmax, no scale-and-subtract per score). Bit-identical to pi3b_gemm followed by pi3b_qknorm_rope, without that pass's read-modify-write of the
 * q and k thirds (822 MB per layer at 100 400 tokens). A, W bf16.
 * replaces: qkv Linear + q_norm / k_norm + rope of pi3/models/layers/attention.py:326-334 in one launch; supersedes the
 *   reference's native rope_2d_cuda_kernel (pi3/models/curope/kernels.cu:17-108), which is a separate in-place pass. */
int pi3b_gemm_qkv(void* stream, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int num_heads, int K,
                  const float* bias, void* out, int64_t ldo, const float* qk_norm, float eps, const float* cs_table,
                  int table_rows, int tokens_per_frame, int n_special, int patch_w, float q_scale);

/* out = softmax(q k^T * scale) v, head_dim 64, bf16, per (sequence, head). Sequence s owns query / output rows
 * [s*q_len, (s+1)*q_len) and key / value rows [s*kv_len, (s+1)*kv_len) (q_len != kv_len: the sequence-parallel case,
 * local queries against all-gathered K/V); head h lives at columns {q,k,v}_col0 + 64*h of its operand; out at 64*h.
 * replaces: F.scaled_dot_product_attention at pi3/models/layers/attention.py:102-104 and :336-338
 *   (frame-local and chunk-global attention of pi3/models/pi3.py:156-166). */
int pi3b_attn_fwd_hd64(void* stream, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                       int64_t ldv, int q_col0, int k_col0, int v_col0, void* out, int64_t ldo, int num_seqs,
                       int q_len, int kv_len, int num_heads, float scale);

/* Same, and additionally lse[h * num_seqs*q_len + row] = max_j(x_j) + log2(sum_j 2^(x_j - max)) with
 * x_j = (q . k_j) * scale * log2(e): the log2-domain log-sum-exp of the row, fp32 [num_heads][num_seqs*q_len].
 * Lets attention over disjoint key/value sets be computed separately and merged (pi3b_attn_merge).
 * replaces: nothing in the reference (its SDPA call is monolithic); exists for the sequence-parallel chunk-global layer
 *   of BASELINE configs[4], where the local K/V block is attended to while the remote blocks cross NVLink. */
int pi3b_attn_fwd_hd64_lse(void* stream, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                           int64_t ldv, int q_col0, int k_col0, int v_col0, void* out, int64_t ldo, float* lse,
                           int num_seqs, int q_len, int kv_len, int num_heads, float scale);

/* pi3b_attn_fwd_hd64 / _lse (lse may be NULL) with a caller-provided scratch buffer: workspace (device, 4-byte aligned,
 * >= 4 * num_seqs * num_heads bytes; NULL = none). With it, sequences longer than 3000 keys (the three-tile kernel) run the BOUNDED-REFERENCE
 * softmax: one streaming pass over K finds max |k|^2 per (sequence, head), |q| max|k| * scale bounds every score of a
 * row, and the softmax is evaluated against that fixed reference instead of tracking the running maximum (identical
 * mathematics, within rounding; rows whose bound would be looser than 2^100 stay on the online path). The workspace
 * must not be shared by calls running concurrently on different streams.
 * replaces: the same call sites as pi3b_attn_fwd_hd64. */
int pi3b_attn_fwd_hd64_ws(void* stream, const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                          int64_t ldv, int q_col0, int k_col0, int v_col0, void* out, int64_t ldo, float* lse,
                          int num_seqs, int q_len, int kv_len, int num_heads, float scale, void* workspace,
                          int64_t workspace_bytes);

/* out[row, 64h + c] = sum_i w_i(row, h) * parts[i][row, 64h + c], w_i = 2^(lses[i] - max_i) / sum_i 2^(lses[i] - max_i):
 * exact combination of n_parts (1..4) attention results computed over disjoint key/value sets. parts: HOST array of
 * n_parts device pointers to bf16 [rows, 64*num_heads] (row pitch ld_part); lses: HOST array of device pointers to fp32
 * [num_heads][rows]; lse_out (optional): merged log-sum-exp. out may alias parts[0].
 * replaces: nothing in the reference; the softmax of pi3/models/layers/attention.py:336-338 split over key blocks. */
int pi3b_attn_merge(void* stream, int n_parts, const void* const* parts, const float* const* lses, int64_t ld_part,
                    void* out, int64_t ldo, float* lse_out, int rows, int num_heads);

/* out[M,D] = LayerNorm(x[M,D]) * gamma + beta (fp32 in; bf16 or fp32 out; in place allowed).
 * If override_rows != NULL, rows r with r % period < n_override are set to override_rows[r % period] instead.
 * D = 384, 768 or 1024.
 * replaces: nn.LayerNorm(eps=1e-6) in pi3/models/dinov2/layers/block.py:88-113, pi3/models/layers/block.py:310-335,
 *   encoder norm + token slicing pi3/models/dinov2/models/vision_transformer.py:271-275 and the register-token
 *   concat of pi3/models/pi3.py:140-144. */
int pi3b_layernorm(void* stream, const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                   void* out, int64_t ldo, int out_f32, int M, int D, const float* override_rows, int period,
                   int n_override);

/* In place on the q and k thirds of packed qkv [M, 3*64*heads] bf16: optional LayerNorm(64) (q/k-norm params
 * fp32 [64], all four NULL to skip) then 2-D RoPE (do_rope) with cs_table fp32 [table_rows][16][2] = (cos, sin) of
 * pos * base^(-j/16). Token t = row % tokens_per_frame; t < n_special => position (0,0) (identity);
 * otherwise (y, x) = ((t-n_special) / patch_w + 1, (t-n_special) % patch_w + 1). q_scale as in pi3b_gemm_qkv.
 * replaces: q_norm/k_norm + rope of pi3/models/layers/attention.py:330-334, RoPE2D pi3/models/layers/pos_embed.py:112-159,
 *   PositionGetter :162-174, and the reference's only native kernel rope_2d_cuda_kernel pi3/models/curope/kernels.cu:17-108
 *   (pybind rope_2d, pi3/models/curope/curope.cpp:49-69). */
int pi3b_qknorm_rope(void* stream, void* qkv, int64_t ld, int M, int num_heads, int tokens_per_frame, int n_special,
                     int patch_w, const float* qn_w, const float* qn_b, const float* kn_w, const float* kn_b, float eps,
                     const float* cs_table, int table_rows, int do_rope, float q_scale);

/* imgs fp32 [N,3,H,W] in [0,1] -> out bf16 [N*(H/14)*(W/14), ldk]: normalised ((x-mean)/std) im2col of the 14x14/s14
 * patch conv, column c*196 + 14*i + j, columns [588, ldk) zero. mean3/std3: HOST pointers to 3 floats.
 * replaces: pi3/models/pi3.py:174 + the input side of pi3/models/dinov2/layers/patch_embed.py:68-81. */
int pi3b_patch_pack(void* stream, const float* imgs, void* out, int N, int H, int W, int ldk, const float* mean3,
                    const float* std3);

/* out[g, r, :] = table[r, :] for g < num_groups (fp32, D % 4 == 0).
 * replaces: cls / register / pos-embed token assembly pi3/models/dinov2/models/vision_transformer.py:215-234. */
int pi3b_broadcast_rows(void* stream, const float* table, float* out, int num_groups, int rows_per_group, int D);

/* fp32 [M,C] -> bf16 [M,C] with independent row pitches (C % 4 == 0).
 * replaces: the autocast input casts of nn.Linear and the torch.cat of pi3/models/pi3.py:168-171. */
int pi3b_cast_f32_bf16(void* stream, const float* x, int64_t ldx, void* out, int64_t ldo, int M, int C);

/* Head unpatchify: local_points[n,y,x,:] = (fx*e^fz, fy*e^fz, e^fz), points = R local + t, conf = raw logit, where
 * f*[..] = feat[n*tokens_per_frame + n_special + (y/14)*(W/14) + x/14][c*196 + 14*(y%14) + x%14]. poses: fp32 [N,4,4].
 * replaces: LinearPts3d pixel_shuffle pi3/models/layers/transformer_head.py:76-81 and pi3/models/pi3.py:194-209. */
int pi3b_unpatchify_points_conf(void* stream, const float* feat_pts, int64_t ld_pts, const float* feat_conf,
                                int64_t ld_conf, const float* poses, float* local_points, float* points, float* conf,
                                int N, int H, int W, int tokens_per_frame, int n_special);

/* out[n, c] = mean over patch tokens t >= n_special of x[n*tokens_per_frame + t, c].
 * replaces: AdaptiveAvgPool2d in pi3/models/layers/camera_head.py:55. */
int pi3b_meanpool_tokens(void* stream, const float* x, int64_t ldx, float* out, int N, int C, int tokens_per_frame,
                         int n_special);

/* poses[n] = [[R, t], [0,0,0,1]] with R = the rotation nearest to the row-normalised 3x3 rot9[n] (fp64 solve).
 * replaces: svd_orthogonalize / convert_pose_to_4x4 pi3/models/layers/camera_head.py:66-93 (torch.svd + torch.det). */
int pi3b_pose_from_9d(void* stream, const float* rot9, int64_t ld_r, const float* trans, int64_t ld_t, float* poses,
                      int N);

/* mask[n,y,x] = sigmoid(conf) > conf_thr && !((max3x3 z - min3x3 z) / z > rtol); z read with element stride z_stride.
 * replaces: OfflineChunkCreator._compute_masks slam/offline_chunk_creator.py:114-119 + depth_edge pi3/utils/geometry.py:347-375. */
int pi3b_mask_depth_edge(void* stream, const float* z, int z_stride, const float* conf, uint8_t* mask, int N, int H,
                         int W, float conf_thr, float rtol);

/* Sample dense maps at keypoints [N,K,2] (x,y): grid_sample(align_corners=False, border) semantics with
 * g = 2*x/(W-1) - 1; bilinear for points / local_points [N,H,W,3] and images [N,3,H,W] (x255 -> uint8, may be NULL),
 * nearest for conf [N,H,W] and mask [N,H,W] (uint8). Bit-exact with what the reference executes (fp32 operation order
 * of ATen's vectorised CPU grid_sample); grid_math selects how the reference computed g: 1 = keypoints lived on a CUDA
 * device (ATen multiplies by the fp32 reciprocal of W-1), 0 = on the CPU (true division).
 * replaces: _interpolate_world_points_for_keypoints slam/offline_chunk_creator.py:129-159 and
 *   GridKeypointExtractor._interpolate_colors utils/keypoint_extraction.py:207-229 (CPU grid_sample after a dense D2H). */
int pi3b_keypoint_gather(void* stream, const float* points, const float* local_points, const float* conf,
                         const uint8_t* mask, const float* images, const float* keypoints, float* out_points,
                         float* out_local, float* out_conf, uint8_t* out_mask, uint8_t* out_colors, int N, int K, int H,
                         int W, int grid_math);

/* out_pts[n, i*ws + j, :] = local_points[n, ys[i], xs[j], :]; out_valid = sigmoid(conf) > conf_thr.
 * replaces: the nearest F.interpolate to 64x64 in recover_focal_shift utils/geometry_torch.py:140-142. */
int pi3b_sample_grid(void* stream, const float* local_points, const float* conf, const int* ys, const int* xs,
                     float* out_pts, uint8_t* out_valid, int N, int H, int W, int hs, int ws, float conf_thr);

/* Batched focal / z-shift recovery: per frame n, the shift that minimises |f * xy / (z + shift) - uv| over the valid
 * samples (f eliminated in closed form), by the algorithm scipy.optimize.least_squares(x0=0, ftol, method='lm') executes
 * for one parameter (2-point finite differences + MINPACK lmder, mode 1, factor 100, <= 100 evaluations): the reference's
 * result is defined by that solver's early stop. pts fp32 [N,S,3], valid uint8 [N,S], uv fp32 [S,2] (pi3b_sample_grid
 * layout). out4[n] = (shift as fp32, focal at that shift, #function evaluations, #valid); fewer than 2 valid samples:
 * (0, 1, 0, count).
 * replaces: solve_optimal_focal_shift utils/geometry_numpy.py:79-96 looped over frames by recover_focal_shift
 *   utils/geometry_torch.py:114-169 (host scipy, 67 ms per 100-frame chunk). */
int pi3b_focal_shift_lm(void* stream, const float* pts, const uint8_t* valid, const float* uv, int N, int S, float ftol,
                        float* out4);

/* out2[0] = lower median of a[i*a_stride] / b[i*b_stride] over mask[i] != 0 (NaN if the mask is empty or any masked
 * ratio is NaN, as torch.median), out2[1] = count.
 * replaces: _get_scale_factor_for_pi3 slam/offline_chunk_creator.py:121-127 (boolean indexing + torch.median). */
int pi3b_median_ratio(void* stream, const float* a, int a_stride, const float* b, int b_stride, const uint8_t* mask,
                      int n, float* out2);

/* Batched closed-form similarity (Umeyama): per problem p, (s,R,t) minimising sum |ref - (s R qry + t)|^2 over
 * valid[p,i] != 0 (valid may be NULL). qry, ref: fp32 [P, n_per, 3]. out18[p] = 4x4 row-major [sR|t], scale, count.
 * replaces: pt.sfm.OptimizeAlignmentSim3(points_qry, points_ref, perform_optimization=False) +
 *   pt.math.Sim3d.exp(...).matrix() at utils/reconstruction_alignment.py:89-105 (un-vendored pytheia==0.2.9). */
int pi3b_umeyama_sim3_batched(void* stream, const float* qry, const float* ref, const uint8_t* valid,
                              int num_problems, int n_per, float* out18);

/* Stage-2 front end: reprojection of a chunk's keypoints into neighbouring views + image-bounds test, as one flat,
 * reference-ordered observation list. points fp32 [N,K,3] (world); w2t fp64 [N,16] = inverse camera poses (row-major,
 * computed by the caller as the reference does); kmat fp64 [N,9] intrinsics or NULL (normalised coordinates);
 * pairs int32 [R,2] = (source frame, target frame) in the reference's loop order. For pair r the keypoints k whose
 * projection (u, v) satisfies 0 <= u < width, 0 <= v < height are emitted in keypoint order. Scratch: row_uv fp64 [R,K,2],
 * row_kp int32 [R,K], row_count int32 [R]. Outputs: row_offset int32 [R+1] (row_offset[R] = number of observations) and,
 * for o < min(total, capacity): obs_src, obs_kp, obs_view (int32) and obs_xy (fp64 [.,2]).
 * replaces: the per-keypoint Python loops of ChunkPTRecon.create_recon_from_chunk utils/chunk_reconstruction.py:158-185
 *   (_project_points_to_other_cams :445-479, _project_3d_points_to_camera :481-509); pytheia only receives the list. */
int pi3b_project_observations(void* stream, const float* points, const double* w2t, const double* kmat, const int* pairs,
                              int R, int K, double width, double height, double* row_uv, int* row_kp, int* row_count,
                              int* row_offset, int* obs_src, int* obs_kp, int* obs_view, double* obs_xy, int capacity);

/* ---- MoGe metric-depth pass (row a15): dense-prediction pieces around pi3b_gemm. Feature maps are channels-last
 * fp32 matrices [H*W, C] (row pitch ld, in elements). ---------------------------------------------------------- */

/* GroupNorm statistics over one sample: mean_rstd[g] = (mean, 1/sqrt(var_biased + eps)) over C/groups channels x rows
 * pixels. workspace: 2*groups doubles (zeroed by the call). groups <= 32, C <= 1024.
 * replaces: nn.GroupNorm(1, C) / nn.GroupNorm(C/32, C) statistics in moge/model/modules.py:48-56. */
int pi3b_groupnorm_stats(void* stream, const float* x, int64_t ldx, int64_t rows, int C, int groups, float eps,
                         double* workspace, float* mean_rstd);

/* out bf16 [H*W, ldo]: column t*C + c of row (y, x) = act(norm(x[clamp(y + t/3 - 1), clamp(x + t%3 - 1)][c])), t = 0..8:
 * the A operand of a 3x3 / stride 1 / replicate-padded conv as a GEMM with W[Cout][t*C + c]. mean_rstd != NULL applies
 * (v - mean_g) * rstd_g * gamma[c] + beta[c] first; relu != 0 applies max(., 0). blocked != 0: the source is the
 * [H/2*W/2, 4*C] output of a 2x2 / stride-2 ConvTranspose GEMM (column block dy*2+dx = output pixel (2Y+dy, 2X+dx)).
 * replaces: GroupNorm -> ReLU -> Conv2d(3x3, padding_mode='replicate') of ResidualConvBlock moge/model/modules.py:47-60
 *   and the ConvTranspose2d -> Conv2d pair of Resampler('conv_transpose') moge/model/modules.py:160-165. */
int pi3b_im2col3x3(void* stream, const float* x, int64_t ldx, int H, int W, int C, int blocked, const float* mean_rstd,
                   int groups, const float* gamma, const float* beta, int relu, void* out, int64_t ldo);

/* x[y*W + x', c] (+)= w[c*ldw]*u_tab[x'] + w[c*ldw + 1]*v_tab[y] + bias[c]   (accumulate != 0: +=, else =).
 * replaces: the uv concat of moge/model/v2.py:141-147 followed by the 1x1 input conv moge/model/modules.py:208-211,245
 *   (normalized_view_plane_uv moge/utils/geometry_torch.py:40-52 supplies u_tab / v_tab). */
int pi3b_uv_affine(void* stream, float* x, int64_t ldx, int H, int W, int C, const float* w, int ldw, const float* bias,
                   const float* u_tab, const float* v_tab, int accumulate);

/* out(c, y, x) = sum_j wy[y*taps_y + j] * sum_i wx[x*taps_x + i] * in(c, min(y0[y] + j, Hi-1), min(x0[x] + i, Wi-1));
 * element (c, y, x) sits at c*sc + y*sy + x*sx (elements). Tap tables are the caller's (ATen bilinear / bilinear-aa).
 * replaces: F.interpolate(bilinear, antialias=True) moge/model/modules.py:121 and F.interpolate(bilinear)
 *   moge/model/v2.py:157. */
int pi3b_resample2d(void* stream, const float* in, int64_t in_sc, int64_t in_sy, int64_t in_sx, int Hi, int Wi, float* out,
                    int64_t out_sc, int64_t out_sy, int64_t out_sx, int C, int Ho, int Wo, const int* y0, const float* wy,
                    int taps_y, const int* x0, const float* wx, int taps_x);

/* points [n,3] in place: mode 0 linear, 1 sinh, 2 exp (z = e^z; xy *= z), 3 sinh_exp; mask_prob = sigmoid(mask_logit)
 * (both may be NULL). replaces: MoGeModel._remap_points + mask sigmoid moge/model/v2.py:112-126, 160-167. */
int pi3b_moge_remap(void* stream, float* points, const float* mask_logit, float* mask_prob, int mode, int64_t n);

/* z' = points[:,2] + shift; m = mask_prob > 0.5 && z' > 0; depth = (m || !apply_mask) ? z' * exp(*log_scale) : +inf;
 * mask_out = m. mask_prob, log_scale (device scalar), mask_out may be NULL.
 * replaces: the depth / mask / metric-scale tail of MoGeModel.infer moge/model/v2.py:236-275. */
int pi3b_moge_depth(void* stream, const float* points, const float* mask_prob, float shift, const float* log_scale,
                    int apply_mask, float* depth, uint8_t* mask_out, int64_t n);

/* Input pipeline: in uint8 [N, Hi, Wi, 3] (decoded RGB frames) -> out fp32 [N, 3, Ho, Wo] in [0,1], bit-identical to
 * PIL.Image.resize((Wo, Ho), BILINEAR) followed by torchvision ToTensor: Pillow's fixed-point coefficients (22 bits,
 * caller-built: min index, tap count, int coefficient rows of length ksize per output row / column), horizontal pass
 * rounded to uint8, then vertical pass, then / 255.
 * replaces: transforms.Resize(target_size) + ToTensor on the PIL image in the DataLoader workers,
 *   datasets/image_datasets.py:193-199 and utils/image_utils.py:98-103. */
int pi3b_resize_u8_pil(void* stream, const uint8_t* in, int N, int Hi, int Wi, float* out, int Ho, int Wo, const int* ymin,
                       const int* ycnt, const int* ky, int ksize_y, const int* xmin, const int* xcnt, const int* kx,
                       int ksize_x);

#ifdef __cplusplus
}
#endif
#endif /* PI3B200_H_ */
