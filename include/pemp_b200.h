/* pemp_b200.h - C ABI of libpemp_b200.so: the B200 (sm_100a) prototype-matching head of Jarvis73/PEMP.
 *
 * Every entry point replaces a piece of the reference's PyTorch hot path; the reference location is
 * cited as `file:line` relative to the reference repository root.
 *
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device; tensors are dense, row-major, in the
 *     layouts written next to each argument; float = IEEE binary32;
 *   - the caller owns every buffer (inputs, outputs, workspace).  The library never allocates, frees or
 *     keeps a pointer after the call returns;
 *   - calls only enqueue work on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     and never synchronise; they are re-entrant and may be issued from several host threads;
 *   - return value: 0 = PEMP_OK, negative = PEMP_E_* argument error (nothing was launched), positive =
 *     the cudaError_t reported by the launch.  No exceptions, no abort, no output on stdout/stderr;
 *   - feature tensors take an *episode stride*: image i of a [B*S, c, hw] operand lives at
 *       base + (i / S) * episode_stride + (i % S) * c * hw        (floats),
 *     so the support (or query) maps can be read in place from the encoder output
 *     `features [B, S+Q, c, h, w]` (episode_stride = (S+Q)*c*hw, networks/pemp_stage1.py:141-144) without the
 *     copy `features[:, :S].reshape(...)` makes for B > 1.  0 means dense (episode_stride = S*c*hw);
 *   - workspace: `pemp_<op>_workspace_bytes(...)` gives the scratch size for the same dimensions; pass a
 *     256-byte aligned device buffer of at least that size.
 */
#ifndef PEMP_B200_H_
#define PEMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEMP_ABI_VERSION 1

enum {
  PEMP_OK = 0,
  PEMP_E_SHAPE = -1,     /* a dimension is <= 0 or outside the supported range             */
  PEMP_E_ALIGN = -2,     /* a pointer does not meet the documented alignment                */
  PEMP_E_WORKSPACE = -3, /* workspace missing or smaller than pemp_<op>_workspace_bytes()   */
  PEMP_E_ARCH = -4,      /* the current device is not sm_100 (B200)                         */
  PEMP_E_NULL = -5       /* a required pointer is NULL                                      */
};

typedef void* pemp_stream_t; /* cudaStream_t */

int pemp_abi_version(void);
const char* pemp_strerror(int code);
/* 0 if the current device can run this library (compute capability 10.x), else PEMP_E_ARCH / cudaError_t. */
int pemp_check_device(void);

/* ---- K0  nearest-neighbour mask down-sampling ------------------------------------------------------
 * replaces  F.interpolate(sup_mask, (h, w), mode="nearest")      networks/pemp_stage1.py:146-148,
 *                                                                networks/pemp_stage2.py:145-147
 * in  [planes, H, W]  ->  out [planes, h, w];  src index = min(floor(dst * (float)in/out), in-1).     */
int pemp_mask_nearest(const float* in, int planes, int H, int W, int h, int w, float* out, pemp_stream_t stream);
/* Same down-sampling fed by the label map the data set stores (uint8: 1 = object, 0 = background, 255 = boundary) instead of
 * the two float planes the loader expands it to:  sup_mask = stack((label == 1), (label == 0))   data_kits/pascal_voc.py:209-210
 * labels [planes, H, W] uint8  ->  out [planes, 2, h, w] float (fg, bg); bit-identical to pemp_mask_nearest on the expansion. */
int pemp_mask_nearest_labels(const uint8_t* labels, int planes, int H, int W, int h, int w, float* out, pemp_stream_t stream);

/* ---- K1  low-resolution masked average pooling (+ K8 Weighted_GAP) ----------------------------------
 * replaces  sum(f*m,-1)/(m.sum(-1)+1e-5); view(B,S,c).mean(1)    networks/pemp_stage1.py:223-227,
 *           pemp_stage2.py:196-200, canet.py:176-178, panet.py:181-186 (query side of alignLoss)
 * fts [B*S, c, hw];  fg / bg: one weight per pixel, image i at  fg + i*mask_stride  (floats); bg may be
 * NULL (then bg_proto is ignored).  Outputs fg_proto / bg_proto [B, c] = mean over the S shots of
 * sum_x f*m / (sum_x m + eps).                                                                         */
size_t pemp_map_pool_workspace_bytes(int B, int S, int c, int hw);
int pemp_map_pool_lowres(const float* fts, long long fts_episode_stride, const float* fg, const float* bg, long long mask_stride,
                         int B, int S, int c, int hw, float eps,
                         float* fg_proto, float* bg_proto,
                         void* workspace, size_t workspace_bytes, pemp_stream_t stream);
/* replaces  Weighted_GAP(supp_feat, mask)                        networks/pfenet.py:15-20  (eps = 5e-4)
 * supp_feat [B, c, hw], mask [B, hw] -> out [B, c].                                                    */
int pemp_weighted_gap(const float* supp_feat, const float* mask, int B, int c, int hw, float* out,
                      void* workspace, size_t workspace_bytes, pemp_stream_t stream);

/* ---- K2  meta-prototype attention -------------------------------------------------------------------
 * replaces  the `self.ctr is not None` branch of mpm()           networks/pemp_stage1.py:202-213,
 *                                                                networks/pemp_stage2.py:174-186
 * fts [B*S, c, hw]; ctr [c, 2p] (columns 0..p-1 = foreground group); fg/bg masks as in K1 (both required).
 * D[k,x] = -sum_c (f[c,x]-ctr[c,k])^2; softmax over the p members of each group; * group mask;
 * centre[c,k] = sum_x f*D / (sum_x D + eps); mean over shots.
 * Outputs fg_proto, bg_proto [B, c, p] and (nullable) adaptive_p [B, c, 2p] (fg columns first,
 * pemp_stage2.py:185).  1 <= p <= 4.                                                                    */
size_t pemp_meta_proto_attn_workspace_bytes(int B, int S, int c, int hw, int p);
int pemp_meta_proto_attn(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg, const float* bg, long long mask_stride,
                         int B, int S, int c, int hw, int p, float eps,
                         float* fg_proto, float* bg_proto, float* adaptive_p,
                         void* workspace, size_t workspace_bytes, pemp_stream_t stream);
/* Diagnostic (tests only; process-wide, not thread-safe): K2 has a TMA-fed persistent kernel for c = 512, p = 3,
 * hw >= 32 and a generic kernel for every other shape.  mode 1 forces the generic kernel so the two can be
 * compared on the same input; mode 0 restores the automatic choice.  Returns the previous mode.         */
int pemp_debug_mpa_path(int mode);

/* ---- K3  cosine matching ----------------------------------------------------------------------------
 * replaces  compute_similarity() [+ .max(dim=2), response map]   networks/pemp_stage1.py:214-222,233-261,
 *           pemp_stage2.py:187-194,205-233, baseline.py:121-149, panet.py:122-156
 * qry [N, c, hw]; fg_proto / bg_proto [Bp, c, P] (P = 1 for [Bp, c]); query n is matched against
 * prototype set n / (N / Bp) (b-major expansion, panet.py:145-149).  cos = sum_c (q/max(|q|,1e-8)) *
 * (p/max(|p|,1e-8)) (F.cosine_similarity of torch >= 2), times `scalar` (dist_scalar = 20).
 * Outputs (each nullable, at least one required):
 *   sim      [N, 2, P, hw]  all similarities, channel 0 = background, 1 = foreground
 *   pred     [N, 2, hw]     max over the P prototypes of each class
 *   response [N, hw] int64  bg-argmax where bg wins, fg-argmax + 3 where fg wins (pemp_stage1.py:217-222)
 * 1 <= P <= 4.                                                                                          */
int pemp_cosine_match(const float* qry, long long qry_episode_stride, const float* fg_proto, const float* bg_proto,
                      int N, int Bp, int c, int hw, int P, float scalar,
                      float* sim, float* pred, int64_t* response, pemp_stream_t stream);
/* Diagnostic (tests only), as pemp_debug_mpa_path: K3 has a TMA-fed persistent kernel for c = 512, P in {1, 3},
 * hw >= 32; mode 1 forces the generic kernel, mode 0 restores the automatic choice.  Returns the previous mode. */
int pemp_debug_cosine_path(int mode);
/* The same switch for K1 / K6 / K7 / K8 (masked average pooling): TMA-fed kernel for c in {256, 512}, hw >= 32. */
int pemp_debug_pool_path(int mode);

/* ---- K4  bilinear up-sampling (align_corners=True) + 2-way argmax -----------------------------------
 * replaces  F.interpolate(pred, out_shape, 'bilinear', align_corners=True) and logits.argmax(1)
 *           networks/pemp_stage1.py:157-162, baseline.py:117, panet.py:116; entry/pemp_stage1.py:52,
 *           entry/pemp_stage2.py:60,64
 * pred [N, 2, h, w] -> (each nullable, at least one required) logits [N, 2, H, W] float,
 * mask8 [N, H, W] uint8, mask64 [N, H, W] int64 (1 where logits[:,1] > logits[:,0]; ties -> 0).
 * Arithmetic order is ATen's: fma(l0, a, l1*b) per axis, scale = (in-1)/(out-1) in float.              */
int pemp_upsample_argmax(const float* pred, int N, int h, int w, int H, int W,
                         float* logits, uint8_t* mask8, int64_t* mask64, pemp_stream_t stream);
/* generic single-plane version of the same resampler: in [planes, h, w] -> out [planes, H, W]
 * (PFENet's mask resize, networks/pfenet.py:191,205)                                                    */
int pemp_bilinear_resize(const float* in, int planes, int h, int w, int H, int W, float* out, pemp_stream_t stream);
/* K5: response.float() nearest-upsampled and cast back (pemp_stage1.py:158-159): in [planes,h,w] int64 */
int pemp_nearest_resize_i64(const int64_t* in, int planes, int h, int w, int H, int W, int64_t* out, pemp_stream_t stream);

/* ---- K6  full-resolution masked average pooling (Baseline / PANet) ----------------------------------
 * replaces  F.interpolate(sup_fts,(H,W),'bilinear',True); sum(up*m,(2,3))/(m.sum((2,3))+1e-5); mean(1)
 *           networks/baseline.py:100-110, panet.py:99-109
 * Uses sum_YX m*(U f) = sum_yx f*(U^T m): the mask is pushed through the transposed bilinear operator
 * to a [B*S, 2, h, w] weight map, the features are read once at low resolution.
 * fts [B*S, c, h*w]; sup_mask [B*S, 2, H, W] (fg, bg) -> fg_proto, bg_proto [B, c].                     */
size_t pemp_map_pool_fullres_workspace_bytes(int B, int S, int c, int h, int w);
int pemp_map_pool_fullres(const float* fts, long long fts_episode_stride, const float* sup_mask, int B, int S, int c, int h, int w, int H, int W,
                          float eps, float* fg_proto, float* bg_proto,
                          void* workspace, size_t workspace_bytes, pemp_stream_t stream);
/* The same pooling fed by the label map the data set stores (uint8 [B*S, H, W]: 1 object, 0 background, 255 boundary) instead of
 * its float expansion  sup_mask = stack((label == 1), (label == 0))   data_kits/pascal_voc.py:209-210, 226-231.
 * Identical prototypes; an eighth of the mask bytes.                                                               */
size_t pemp_map_pool_fullres_labels_workspace_bytes(int B, int S, int c, int h, int w, int H, int W);
int pemp_map_pool_fullres_labels(const float* fts, long long fts_episode_stride, const uint8_t* labels, int B, int S, int c,
                                 int h, int w, int H, int W, float eps, float* fg_proto, float* bg_proto, void* workspace,
                                 size_t workspace_bytes, pemp_stream_t stream);
/* the adjoint resampler on its own: mask [planes, H, W] -> wt [planes, h, w], msum [planes] (nullable) =
 * plain sum of the mask plane.                                                                          */
int pemp_bilinear_adjoint(const float* mask, int planes, int H, int W, int h, int w, float* wt, float* msum,
                          pemp_stream_t stream);

/* ---- K7  PANet prototype-alignment reverse pass ------------------------------------------------------
 * replaces  PANet.alignLoss()                                    networks/panet.py:158-194
 * qry_fts [B*Q, c, h*w]; pred [B*Q, 2, h*w] low-res logits; sup_fts [B*S, c, h*w];
 * sup_mask_fg: plane i at sup_mask_fg + i*mask_stride, [H, W] floats used as class labels (long cast).
 * loss[0] = mean over B*S*H*W of -log_softmax(upsampled reverse logits)[label].                         */
size_t pemp_panet_align_workspace_bytes(int B, int S, int Q, int c, int h, int w, int H, int W);
int pemp_panet_align(const float* qry_fts, long long qry_episode_stride, const float* pred,
                     const float* sup_fts, long long sup_episode_stride,
                     const float* sup_mask_fg, long long mask_stride,
                     int B, int S, int Q, int c, int h, int w, int H, int W, float scalar,
                     float* loss, void* workspace, size_t workspace_bytes, pemp_stream_t stream);
/* The same loss against the uint8 label map (plane i at labels + i*label_stride, [H, W]; 1 = object, everything else is not
 * foreground) that `sup_mask_fg = (label == 1)` was expanded from  (data_kits/pascal_voc.py:209-210).                 */
int pemp_panet_align_labels(const float* qry_fts, long long qry_episode_stride, const float* pred,
                            const float* sup_fts, long long sup_episode_stride,
                            const uint8_t* labels, long long label_stride,
                            int B, int S, int Q, int c, int h, int w, int H, int W, float scalar,
                            float* loss, void* workspace, size_t workspace_bytes, pemp_stream_t stream);

/* ---- K9  PFENet prior mask ----------------------------------------------------------------------------
 * replaces  the prior block of PFENet.forward()                  networks/pfenet.py:201-231
 * q4 [B, C, hw_q]; s4 [S, B, C, hw_s] (layer-4 support features per shot); smask [S, B, hw_s] the support
 * mask already resized to the feature size (pemp_bilinear_resize).  Per shot: sim = (s*m)^T q /
 * (|s*m| |q|^T + 1e-7); max over support pixels; min-max normalise over query pixels (+1e-7); mean over
 * shots -> prior [B, hw_q].
 * precision: 0 = bf16 tcgen05 GEMM (fast; ~6e-4 norm-wise on the pre-normalisation cosine),
 *            1 = fp32 CUDA-core GEMM (reference-grade), 2 = 3-term bf16 split on tcgen05 (fp32-grade).
 * rowmax (nullable) [S, B, hw_q] receives the pre-normalisation maxima for tolerance statements.        */
size_t pemp_prior_mask_workspace_bytes(int B, int S, int C, int hw_s, int hw_q, int precision);
int pemp_prior_mask(const float* q4, const float* s4, const float* smask,
                    int B, int S, int C, int hw_s, int hw_q, int precision,
                    float* prior, float* rowmax, void* workspace, size_t workspace_bytes, pemp_stream_t stream);

/* ---- K10  FewShotMetric confusion counts --------------------------------------------------------------
 * replaces  FewShotMetric.update()                               core/metrics.py:9-23
 * pred, ref [N, npix] uint8; cls [N] int64 in 1..num_classes; label 255 in ref is ignored.
 * stat [(num_classes+1), 3] int64 (tp, fp, fn) is ACCUMULATED: row 0 += background counts of every
 * episode, row cls[i] += foreground counts of episode i.  Integer atomics: order-independent, exact.     */
int pemp_iou_hist(const uint8_t* pred, const uint8_t* ref, const int64_t* cls, int N, long long npix,
                  int num_classes, int64_t* stat, pemp_stream_t stream);
/* K4 + K10 in one launch (what Evaluator.test_step + FewShotMetric.update do per batch, entry/pemp_stage2.py:63-65,
 * core/metrics.py:9-23): up-sample pred [N,2,h,w] to (H,W), write the argmax mask8 [N,H,W] and accumulate the counts of
 * (mask8, ref [N,H,W] uint8, cls [N]) into stat [(num_classes+1), 3].  mask8 and ref must be 4-byte aligned.          */
int pemp_upsample_argmax_hist(const float* pred, int N, int h, int w, int H, int W, uint8_t* mask8,
                              const uint8_t* ref, const int64_t* cls, int num_classes, int64_t* stat,
                              pemp_stream_t stream);

/* ---- K11 communication module of the Stage-2 backbones ("next" row) ------------------------------------
 * replaces  ResNetCM.comm / VGG16CM.comm                         networks/backbones.py:208-222, 469-479
 *   mask_out = max_pool2d(mask_in, 3, stride, 1);  p = x * mask_out;
 *   feat = linear(cat(mean_hw(p).view(B, spq, c).mean(1), max_hw(p).view(B, spq, c).mean(1)));  out = feat broadcast
 * x [N, c, h, w] with N = B*spq; mask_in [N, 1, Hm, Wm]; weight [n_out, 2c] and bias [n_out] (nullable) as in
 * nn.Linear; outputs mask_out [N, 1, h, w] and out [N, n_out, h, w].  h, w must equal the pooled size
 * floor((Hm + 2 - 3) / stride) + 1.                                                                      */
size_t pemp_comm_workspace_bytes(int N, int c, int spq, int n_out);
int pemp_comm_module(const float* x, const float* mask_in, int N, int c, int h, int w, int Hm, int Wm, int stride,
                     int spq, const float* weight, const float* bias, int n_out, float* mask_out, float* out,
                     void* workspace, size_t workspace_bytes, pemp_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K12  Training path: what autograd records for the head in entry/pemp_stage1.py:57-65.
 * pemp_meta_proto_attn_train = pemp_meta_proto_attn (same kernels, same fg_proto / bg_proto) that also stores the
 *   per-shot centres shot_centre [B*S, c, 2p] (foreground columns first) and denominators shot_den [B*S, 2p]
 *   (sum of the attention + eps) for the backward.  Workspace: pemp_meta_proto_attn_workspace_bytes.
 * pemp_meta_proto_attn_bwd: g_fg / g_bg [B, c, p] = gradient of fg_proto / bg_proto  ->  d_fts (image (b, s) at
 *   d_fts + b * d_fts_episode_stride + s * c * hw; stride 0 = dense [B*S, c, hw] - a non-zero stride writes straight into
 *   the support half of the gradient of the encoder output [B, S+Q, c, hw]), d_ctr [c, 2p].  Masks get no gradient.
 * pemp_cosine_match_bwd: g_pred [N, 2, hw] = gradient of pred (after the max over prototypes; the arg-max is recomputed,
 *   first maximum wins)  ->  d_qry (same addressing with d_qry_episode_stride), d_fg / d_bg [Bp, c, P].  c <= 1024.
 * pemp_cosine_sim_bwd: the same for the per-prototype maps of `compute_similarity` used on their own under autograd
 *   (networks/pemp_stage1.py:233-261, baseline.py:121-149, panet.py:122-156): g_sim [N, 2, P, hw] = gradient of
 *   `sim` of pemp_cosine_match (channel 0 background), no arg-max.  Workspace: pemp_cosine_match_bwd_workspace_bytes. */
int pemp_meta_proto_attn_train(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg,
                               const float* bg, long long mask_stride, int B, int S, int c, int hw, int p, float eps,
                               float* fg_proto, float* bg_proto, float* shot_centre, float* shot_den, void* workspace,
                               size_t workspace_bytes, pemp_stream_t stream);
size_t pemp_meta_proto_attn_bwd_workspace_bytes(int B, int S, int c, int hw, int p);
int pemp_meta_proto_attn_bwd(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg,
                             const float* bg, long long mask_stride, const float* shot_centre, const float* shot_den,
                             const float* g_fg, const float* g_bg, int B, int S, int c, int hw, int p, float* d_fts,
                             long long d_fts_episode_stride, float* d_ctr, void* workspace, size_t workspace_bytes,
                             pemp_stream_t stream);
/* Diagnostic (tests only), as pemp_debug_mpa_path: the backward of K2 (p = 3, c in {128, 256, 512, 1024}) and of K3 (P = 3,
 * c in {256, 512}) have kernels that run the three per-tile products on the warp-level tensor path (3 x TF32, fp32-grade) and
 * CUDA-core kernels for every other shape; mode 1 forces the latter, mode 0 restores the automatic choice.  Returns the
 * previous mode.                                                                                                        */
int pemp_debug_bwd_path(int mode);
size_t pemp_cosine_match_bwd_workspace_bytes(int N, int Bp, int c, int hw, int P);
int pemp_cosine_match_bwd(const float* qry, long long qry_episode_stride, const float* fg_proto, const float* bg_proto,
                          const float* g_pred, int N, int Bp, int c, int hw, int P, float scalar, float* d_qry,
                          long long d_qry_episode_stride, float* d_fg, float* d_bg, void* workspace, size_t workspace_bytes,
                          pemp_stream_t stream);
int pemp_cosine_sim_bwd(const float* qry, long long qry_episode_stride, const float* fg_proto, const float* bg_proto,
                        const float* g_sim, int N, int Bp, int c, int hw, int P, float scalar, float* d_qry,
                        long long d_qry_episode_stride, float* d_fg, float* d_bg, void* workspace, size_t workspace_bytes,
                        pemp_stream_t stream);

/* backward of pemp_map_pool_lowres (K1; training path of the baseline and PANet heads, entry/panet.py:108-115):
 * g_fg / g_bg [B, c] -> d_fts, addressed like pemp_meta_proto_attn_bwd's.  bg and g_bg may both be NULL.            */
int pemp_map_pool_lowres_bwd(const float* fg, const float* bg, long long mask_stride, const float* g_fg, const float* g_bg,
                             int B, int S, int c, int hw, float eps, float* d_fts, long long d_fts_episode_stride,
                             pemp_stream_t stream);
/* K13  loss of the training step and its gradient (entry/pemp_stage1.py:51,57-60): mean cross entropy, 255 ignored, of the
 * bilinear (align_corners) up-sampling of pred [N, 2, h, w] to the target [N, H, W] (int64, or uint8 when target_is_u8).
 * loss [1]; d_pred [N, 2, h, w] = d loss / d pred (nullable: loss only); weight [N, H, W] nullable (see K14).          */
size_t pemp_upsample_ce_workspace_bytes(int N, int h, int w, int H, int W);
int pemp_upsample_ce(const float* pred, const void* target, int target_is_u8, const float* weight, int N, int h, int w,
                     int H, int W, float* loss, float* d_pred, void* workspace, size_t workspace_bytes,
                     pemp_stream_t stream);
/* K14  CELossDT (core/losses.py:17-43): weight [N, H, W] = exp(-d / sigma^2) + 1, d = exact Euclidean
 * distance to the nearest boundary pixel of (target == 1) - the reference computes it with scipy on the host every step.
 * Passing the result as `weight` to pemp_upsample_ce gives CELossDT's loss: sum(w * ce) / sum(w) (weight == NULL: plain
 * mean over the valid pixels).                                                                                          */
size_t pemp_boundary_weight_workspace_bytes(int N, int H, int W);
int pemp_boundary_weight(const void* target, int target_is_u8, int N, int H, int W, float sigma, float* weight,
                         void* workspace, size_t workspace_bytes, pemp_stream_t stream);

/* K15  CaNet dense-comparison input (networks/canet.py:172-180): out [N, 2c, hw] = the query maps
 * (read in place through qry_episode_stride, as in pemp_cosine_match) concatenated with z [Bp, c] tiled over hw.  z is
 * pemp_map_pool_lowres with the nearest-resized foreground mask (bg = NULL), eps 1e-5.                               */
int pemp_canet_concat(const float* qry, long long qry_episode_stride, const float* z, int N, int Bp, int c, int hw,
                      float* out, pemp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PEMP_B200_H_ */
