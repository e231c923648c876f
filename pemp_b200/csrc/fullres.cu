// K6  full-resolution masked average pooling of Baseline / PANet without the up-sampled feature copy.
//
// replaces  sup_fts = F.interpolate(sup_fts, (H, W), 'bilinear', align_corners=True);
//           sum(sup_fts * mask, (2,3)) / (mask.sum((2,3)) + 1e-5); view(B,S,-1).mean(1)
//   networks/baseline.py:100-110, networks/panet.py:99-109       (329 MB / shot of temporaries at 401x401)
//
// Identity:  sum_{YX} m[Y,X] (U f)[Y,X] = sum_{yx} f[y,x] (U^T m)[y,x]  with U the bilinear operator.  The
// mask (2*H*W floats per shot) is pushed through U^T once (resample.cu: adjoint_rows_kernel), then the
// features are pooled at low resolution by the K1 kernel with the exact mask sums as denominators.
// Algorithmic bytes per shot: c*h*w*4 + 2*H*W*4.
#include "common.cuh"

int pemp_pool_launch(const float* fts, long long ep_stride, const float* fg, const float* bg, long long mask_stride, int B, int S, int c,
                     int hw, float eps, const float* den_override, float* fg_proto, float* bg_proto, void* workspace,
                     size_t workspace_bytes, cudaStream_t st);

size_t pemp_adjoint_scratch_bytes(int planes, int h, int w);
int pemp_adjoint_launch(const float* mask, int planes, int H, int W, int h, int w, float* wt, float* msum, char* scratch,
                        cudaStream_t st);

size_t pemp_adjoint_labels_extra_bytes(int images, int H, int W, int h, int w);
int pemp_adjoint_launch_labels(const uint8_t* labels, int images, int H, int W, int h, int w, float* wt, float* msum, char* scratch,
                               float* expanded, cudaStream_t st);

namespace {
struct Plan {
  size_t off_wt, off_sum, off_adj, off_pool, total;
};
Plan make_plan(int B, int S, int c, int h, int w) {
  Plan p;
  size_t planes = static_cast<size_t>(B) * S * 2;
  p.off_wt = 0;
  p.off_sum = align_up(planes * h * w * sizeof(float), 256);
  p.off_adj = p.off_sum + align_up(planes * sizeof(float), 256);
  p.off_pool = p.off_adj + pemp_adjoint_scratch_bytes(static_cast<int>(planes), h, w);
  p.total = p.off_pool + pemp_map_pool_workspace_bytes(B, S, c, h * w);
  return p;
}
}  // namespace

extern "C" size_t pemp_map_pool_fullres_workspace_bytes(int B, int S, int c, int h, int w) {
  if (B <= 0 || S <= 0 || c <= 0 || h <= 0 || w <= 0) return 0;
  return make_plan(B, S, c, h, w).total;
}

extern "C" int pemp_map_pool_fullres(const float* fts, long long fts_episode_stride, const float* sup_mask, int B, int S, int c, int h, int w, int H,
                                     int W, float eps, float* fg_proto, float* bg_proto, void* workspace,
                                     size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(fts && sup_mask && fg_proto && bg_proto, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && c > 0 && h > 0 && w > 0 && H > 0 && W > 0, PEMP_E_SHAPE);
  Plan pl = make_plan(B, S, c, h, w);
  PEMP_REQUIRE(workspace && workspace_bytes >= pl.total, PEMP_E_WORKSPACE);
  char* ws = static_cast<char*>(workspace);
  float* wt = reinterpret_cast<float*>(ws + pl.off_wt);
  float* msum = reinterpret_cast<float*>(ws + pl.off_sum);
  int planes = B * S * 2;
  int rc = pemp_adjoint_launch(sup_mask, planes, H, W, h, w, wt, msum, ws + pl.off_adj, as_stream(stream));
  if (rc != PEMP_OK) return rc;
  const int hw = h * w;
  return pemp_pool_launch(fts, fts_episode_stride, wt, wt + hw, 2LL * hw, B, S, c, hw, eps, msum, fg_proto, bg_proto, ws + pl.off_pool,
                          workspace_bytes - pl.off_pool, as_stream(stream));
}

// ---- the same pooling fed by the label map the data set stores (1 object / 0 background / 255 boundary) ---------------------
// instead of the two float planes the loader expands it to (data_kits/pascal_voc.py:209-210): identical prototypes (the planes
// are formed on the fly, fg = (label == 1), bg = (label == 0)), an eighth of the mask bytes from HBM and from the host.
extern "C" size_t pemp_map_pool_fullres_labels_workspace_bytes(int B, int S, int c, int h, int w, int H, int W) {
  if (B <= 0 || S <= 0 || c <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return 0;
  return make_plan(B, S, c, h, w).total + pemp_adjoint_labels_extra_bytes(B * S, H, W, h, w);
}

extern "C" int pemp_map_pool_fullres_labels(const float* fts, long long fts_episode_stride, const uint8_t* labels, int B, int S, int c,
                                            int h, int w, int H, int W, float eps, float* fg_proto, float* bg_proto, void* workspace,
                                            size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(fts && labels && fg_proto && bg_proto, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && c > 0 && h > 0 && w > 0 && H > 0 && W > 0, PEMP_E_SHAPE);
  Plan pl = make_plan(B, S, c, h, w);
  const size_t extra = pemp_adjoint_labels_extra_bytes(B * S, H, W, h, w);
  PEMP_REQUIRE(workspace && workspace_bytes >= pl.total + extra, PEMP_E_WORKSPACE);
  char* ws = static_cast<char*>(workspace);
  float* wt = reinterpret_cast<float*>(ws + pl.off_wt);
  float* msum = reinterpret_cast<float*>(ws + pl.off_sum);
  float* expanded = extra ? reinterpret_cast<float*>(ws + pl.total) : nullptr;
  int rc = pemp_adjoint_launch_labels(labels, B * S, H, W, h, w, wt, msum, ws + pl.off_adj, expanded, as_stream(stream));
  if (rc != PEMP_OK) return rc;
  const int hw = h * w;
  return pemp_pool_launch(fts, fts_episode_stride, wt, wt + hw, 2LL * hw, B, S, c, hw, eps, msum, fg_proto, bg_proto, ws + pl.off_pool,
                          pl.total - pl.off_pool, as_stream(stream));
}
