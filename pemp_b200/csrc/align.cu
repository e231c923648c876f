// K7  PANet prototype-alignment reverse pass (alignLoss).
//
// replaces networks/panet.py:158-194:
//   pred.argmax(1) -> query fg/bg masks -> masked average pooling of the QUERY features (eps 1e-5), mean over Q
//   -> cosine of every SUPPORT feature map against those prototypes (x dist_scalar) -> bilinear up-sampling to
//   (H, W) -> F.cross_entropy against sup_mask_fg.long() (mean over B*S*H*W, no ignore label).
//
// Built from the K1 pooling and K3 matching kernels plus two small kernels here: the argmax mask and a fused
// "up-sample + 2-class log-softmax + NLL + block sum" that never writes the [BS, 2, H, W] logits.
// Algorithmic bytes: (Q+S)*c*hw*4 + 2*hw*4 + S*H*W*4 per episode.
#include "common.cuh"

int pemp_pool_launch(const float* fts, long long ep_stride, const float* fg, const float* bg, long long mask_stride, int B, int S, int c,
                     int hw, float eps, const float* den_override, float* fg_proto, float* bg_proto, void* workspace,
                     size_t workspace_bytes, cudaStream_t st);

namespace {

constexpr int kCeThreads = 256;

// class of a pixel of the support mask: float plane `sup_mask_fg` (`.long()` truncation, panet.py:190) or the uint8 label map it
// was expanded from (1 object; 0 background and 255 boundary are not foreground)
template <typename LT>
__device__ __forceinline__ int ce_label(const LT* __restrict__ label, long long idx) {
  if constexpr (sizeof(LT) == 1) {
    return __ldg(label + idx) == 1 ? 1 : 0;
  } else {
    return static_cast<int>(__ldg(label + idx));
  }
}

// pred [N, 2, hw] -> mask [N, 2, hw]: plane 0 = (argmax == 1), plane 1 = (argmax == 0)
__global__ void argmax_masks_kernel(const float* __restrict__ pred, float* __restrict__ mask, long long total, int hw) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long n = i / hw;
    int x = static_cast<int>(i - n * hw);
    float bgv = __ldg(pred + (n * 2 + 0) * hw + x), fgv = __ldg(pred + (n * 2 + 1) * hw + x);
    float isfg = fgv > bgv ? 1.f : 0.f;
    mask[(n * 2 + 0) * hw + x] = isfg;
    mask[(n * 2 + 1) * hw + x] = 1.f - isfg;
  }
}

// rev [N, 2, h, w] low-res reverse logits, label planes [N][H*W] floats -> partial[blockIdx] = sum of NLL
template <typename LT>
__global__ void __launch_bounds__(kCeThreads)
upsample_ce_kernel(const float* __restrict__ rev, const LT* __restrict__ label, long long label_stride, long long total,
                   int h, int w, int H, int W, float sy, float sx, float* __restrict__ partial) {
  const long long HW = static_cast<long long>(H) * W;
  const int hw = h * w;
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long n = i / HW;
    long long r = i - n * HW;
    int Y = static_cast<int>(r / W), X = static_cast<int>(r - static_cast<long long>(Y) * W);
    Lerp ly = lerp_coeff(Y, sy, h), lx = lerp_coeff(X, sx, w);
    float v[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* p = rev + (n * 2 + ch) * hw;
      float a = __ldg(p + ly.i0 * w + lx.i0), b = __ldg(p + ly.i0 * w + lx.i1);
      float c = __ldg(p + ly.i1 * w + lx.i0), d = __ldg(p + ly.i1 * w + lx.i1);
      v[ch] = lerp2(ly.l0, lerp2(lx.l0, a, lx.l1, b), ly.l1, lerp2(lx.l0, c, lx.l1, d));
    }
    const int lab = ce_label<LT>(label, n * label_stride + r);
    const float m = fmaxf(v[0], v[1]);
    const float lse = m + logf(expf(v[0] - m) + expf(v[1] - m));
    acc += lse - (lab == 1 ? v[1] : v[0]);
  }
  __shared__ float part[kCeThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < kCeThreads / 32 ? part[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

// Banded version (same idea as upsample_argmax_band_kernel in resample.cu): the bilinear sample is
// fma(ly.l0, Hrow(i0, X), ly.l1 * Hrow(i1, X)) with Hrow(y, X) = fma(lx.l0, p[y][x0], lx.l1 * p[y][x1]); a CTA owns a band
// of kCeBand output rows of one image, computes Hrow once for the few source rows the band touches (both channels) and
// then needs 4 shared loads, 2 lerps, one exp and one log per pixel:  lse - v_label = max + log(1 + exp(-|v0 - v1|)) -
// v_label.  The label rows of a band are one contiguous, coalesced range.  partial[image * bands + band].
constexpr int kCeBand = 16, kCeMaxSrc = 8;
template <typename LT>
__global__ void __launch_bounds__(kCeThreads)
upsample_ce_band_kernel(const float* __restrict__ rev, const LT* __restrict__ label, long long label_stride, int h, int w,
                        int H, int W, float sy, float sx, int bands, float* __restrict__ partial) {
  extern __shared__ float hrow[];                    // [nsrc][2][W]
  const int n = blockIdx.x / bands, band = blockIdx.x - n * bands;
  const int Y0 = band * kCeBand, Y1 = min(H, Y0 + kCeBand);
  const int src0 = lerp_coeff(Y0, sy, h).i0, nsrc = lerp_coeff(Y1 - 1, sy, h).i1 - src0 + 1;
  const int hw = h * w;
  const float* p0 = rev + static_cast<long long>(n) * 2 * hw + src0 * w;
  for (int X = threadIdx.x; X < W; X += blockDim.x) {
    const Lerp lx = lerp_coeff(X, sx, w);
    for (int r = 0; r < nsrc; ++r)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float* row = p0 + ch * hw + r * w;
        hrow[(r * 2 + ch) * W + X] = lerp2(lx.l0, __ldg(row + lx.i0), lx.l1, __ldg(row + lx.i1));
      }
  }
  __syncthreads();
  const long long lab0 = n * label_stride;
  float acc = 0.f;
  for (int Y = Y0; Y < Y1; ++Y) {
    const Lerp ly = lerp_coeff(Y, sy, h);
    const float* rt = hrow + (ly.i0 - src0) * 2 * W;
    const float* rb = hrow + (ly.i1 - src0) * 2 * W;
    const long long lrow = lab0 + static_cast<long long>(Y) * W;
    for (int X = threadIdx.x; X < W; X += blockDim.x) {
      const float v0 = lerp2(ly.l0, rt[X], ly.l1, rb[X]);
      const float v1 = lerp2(ly.l0, rt[W + X], ly.l1, rb[W + X]);
      const int lb = ce_label<LT>(label, lrow + X);
      const float m = fmaxf(v0, v1);
      acc += (m + logf(1.0f + expf(-fabsf(v0 - v1)))) - (lb == 1 ? v1 : v0);
    }
  }
  __shared__ float part[kCeThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s2 = threadIdx.x < kCeThreads / 32 ? part[threadIdx.x] : 0.f;
    s2 = warp_sum(s2);
    if (threadIdx.x == 0) partial[blockIdx.x] = s2;
  }
}

// Round-2 version of the banded kernel.  With two classes the loss of a pixel depends on the logits only through their
// difference d = v1 - v0:  lse(v0, v1) - v_label = softplus(-d) for label 1 and softplus(d) for label 0, and bilinear
// up-sampling is linear, so ONE plane (rev1 - rev0, formed while the horizontal pass reads the low-res rows) is up-sampled
// instead of two: half the shared-memory rows, loads and lerps per pixel.  softplus(t) = max(t, 0) + log(1 + exp(-|t|)) with
// ex2 / lg2 approximations (|error| < 3e-7 per pixel, the loss is a mean over S*H*W pixels compared at 1e-5).  A thread keeps
// its columns and walks the rows of the band with the label loads of four rows in flight.
#ifndef PEMP_CE_BAND
#define PEMP_CE_BAND 32
#endif
constexpr int kCeBand2 = PEMP_CE_BAND;
template <typename LT>
__global__ void __launch_bounds__(kCeThreads)
upsample_ce_band2_kernel(const float* __restrict__ rev, const LT* __restrict__ label, long long label_stride, int h, int w,
                         int H, int W, float sy, float sx, int bands, float* __restrict__ partial) {
  extern __shared__ float hrow[];                    // [nsrc][W]  horizontal pass of (rev1 - rev0)
  const int n = blockIdx.x / bands, band = blockIdx.x - n * bands;
  const int Y0 = band * kCeBand2, Y1 = min(H, Y0 + kCeBand2);
  const int src0 = lerp_coeff(Y0, sy, h).i0, nsrc = lerp_coeff(Y1 - 1, sy, h).i1 - src0 + 1;
  const int hw = h * w;
  const float* p0 = rev + static_cast<long long>(n) * 2 * hw + src0 * w;
  for (int X = threadIdx.x; X < W; X += kCeThreads) {
    const Lerp lx = lerp_coeff(X, sx, w);
    for (int r = 0; r < nsrc; ++r) {
      const float* row = p0 + r * w;
      const float a = __ldg(row + hw + lx.i0) - __ldg(row + lx.i0), b = __ldg(row + hw + lx.i1) - __ldg(row + lx.i1);
      hrow[r * W + X] = lerp2(lx.l0, a, lx.l1, b);
    }
  }
  __syncthreads();
  float acc = 0.f;
  for (int X = threadIdx.x; X < W; X += kCeThreads) {
    const long long lcol = n * label_stride + static_cast<long long>(Y0) * W + X;
#pragma unroll 4
    for (int Y = Y0; Y < Y1; ++Y) {
      const Lerp ly = lerp_coeff(Y, sy, h);
      const float d = lerp2(ly.l0, hrow[(ly.i0 - src0) * W + X], ly.l1, hrow[(ly.i1 - src0) * W + X]);
      const int lb = ce_label<LT>(label, lcol + static_cast<long long>(Y - Y0) * W);
      const float t = lb == 1 ? -d : d;
      acc += fmaxf(t, 0.f) + __logf(1.0f + __expf(-fabsf(t)));
    }
  }
  __shared__ float part[kCeThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s2 = threadIdx.x < kCeThreads / 32 ? part[threadIdx.x] : 0.f;
    s2 = warp_sum(s2);
    if (threadIdx.x == 0) partial[blockIdx.x] = s2;
  }
}

// single CTA: add the block partials in index order in double, divide by the element count
__global__ void ce_finalize_kernel(const float* __restrict__ partial, int n, double count, float* __restrict__ loss) {
  __shared__ double part[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += static_cast<double>(partial[i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += part[i];
    loss[0] = static_cast<float>(t / count);
  }
}

struct Plan {
  int ce_blocks;
  size_t off_qmask, off_proto, off_rev, off_partial, off_pool, total;
};
Plan make_plan(int B, int S, int Q, int c, int h, int w, int H, int W) {
  Plan p;
  const size_t hw = static_cast<size_t>(h) * w;
  long long total = static_cast<long long>(B) * S * H * W;
  p.ce_blocks = static_cast<int>(llmin((total + kCeThreads - 1) / kCeThreads, 148LL * 16));
  const long long band_blocks = static_cast<long long>(B) * S * ((H + (kCeBand < kCeBand2 ? kCeBand : kCeBand2) - 1) / (kCeBand < kCeBand2 ? kCeBand : kCeBand2));
  if (band_blocks > p.ce_blocks) p.ce_blocks = static_cast<int>(band_blocks);   // the banded kernel writes one partial per band
  p.off_qmask = 0;
  p.off_proto = align_up(static_cast<size_t>(B) * Q * 2 * hw * sizeof(float), 256);
  p.off_rev = p.off_proto + align_up(static_cast<size_t>(B) * c * 2 * sizeof(float), 256);
  p.off_partial = p.off_rev + align_up(static_cast<size_t>(B) * S * 2 * hw * sizeof(float), 256);
  p.off_pool = p.off_partial + align_up(static_cast<size_t>(p.ce_blocks) * sizeof(float), 256);
  p.total = p.off_pool + pemp_map_pool_workspace_bytes(B, Q, c, static_cast<int>(hw));
  return p;
}

}  // namespace

extern "C" size_t pemp_panet_align_workspace_bytes(int B, int S, int Q, int c, int h, int w, int H, int W) {
  if (B <= 0 || S <= 0 || Q <= 0 || c <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return 0;
  return make_plan(B, S, Q, c, h, w, H, W).total;
}

namespace {
template <typename LT>
int panet_align_impl(const float* qry_fts, long long qry_episode_stride, const float* pred, const float* sup_fts,
                     long long sup_episode_stride, const LT* sup_mask_fg, long long mask_stride, int B, int S, int Q,
                     int c, int h, int w, int H, int W, float scalar, float* loss, void* workspace, size_t workspace_bytes,
                     pemp_stream_t stream) {
  PEMP_REQUIRE(qry_fts && pred && sup_fts && sup_mask_fg && loss, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && Q > 0 && c > 0 && h > 0 && w > 0 && H > 0 && W > 0, PEMP_E_SHAPE);
  Plan pl = make_plan(B, S, Q, c, h, w, H, W);
  PEMP_REQUIRE(workspace && workspace_bytes >= pl.total, PEMP_E_WORKSPACE);
  char* ws = static_cast<char*>(workspace);
  float* qmask = reinterpret_cast<float*>(ws + pl.off_qmask);
  float* fgp = reinterpret_cast<float*>(ws + pl.off_proto);
  float* bgp = fgp + static_cast<size_t>(B) * c;
  float* rev = reinterpret_cast<float*>(ws + pl.off_rev);
  float* partial = reinterpret_cast<float*>(ws + pl.off_partial);
  cudaStream_t st = as_stream(stream);
  const int hw = h * w;

  long long npx = static_cast<long long>(B) * Q * hw;
  argmax_masks_kernel<<<static_cast<unsigned>(llmin((npx + 255) / 256, 148LL * 8)), 256, 0, st>>>(pred, qmask, npx, hw);
  // query prototypes: "shots" of the pooling kernel are the Q queries of an episode (panet.py:181-186)
  int rc = pemp_pool_launch(qry_fts, qry_episode_stride, qmask, qmask + hw, 2LL * hw, B, Q, c, hw, 1e-5f, nullptr, fgp, bgp, ws + pl.off_pool,
                            workspace_bytes - pl.off_pool, st);
  if (rc != PEMP_OK) return rc;
  // reverse matching: every support map against its episode's query prototypes (panet.py:189, b-major expansion)
  rc = pemp_cosine_match(sup_fts, sup_episode_stride, fgp, bgp, B * S, B, c, hw, 1, scalar, nullptr, rev, nullptr, stream);
  if (rc != PEMP_OK) return rc;
  long long total = static_cast<long long>(B) * S * H * W;
  const float sy = lerp_scale(h, H), sx = lerp_scale(w, W);
  int nparts;
#ifndef PEMP_CE_V1
  const int nsrc2 = static_cast<int>((kCeBand2 - 1) * sy) + 3;
  const size_t smem2 = static_cast<size_t>(nsrc2 < h ? nsrc2 : h) * W * sizeof(float);
  if (smem2 <= 48 * 1024) {
    const int bands = (H + kCeBand2 - 1) / kCeBand2;
    nparts = B * S * bands;
    upsample_ce_band2_kernel<LT><<<nparts, kCeThreads, smem2, st>>>(rev, sup_mask_fg, mask_stride, h, w, H, W, sy, sx, bands, partial);
    ce_finalize_kernel<<<1, 256, 0, st>>>(partial, nparts, static_cast<double>(total), loss);
    return launch_status();
  }
#endif
  const int nsrc_max = static_cast<int>((kCeBand - 1) * sy) + 3;
  const size_t smem = static_cast<size_t>(nsrc_max < h ? nsrc_max : h) * 2 * W * sizeof(float);
  if (nsrc_max <= kCeMaxSrc && smem <= 48 * 1024) {
    const int bands = (H + kCeBand - 1) / kCeBand;
    nparts = B * S * bands;
    upsample_ce_band_kernel<LT><<<nparts, kCeThreads, smem, st>>>(rev, sup_mask_fg, mask_stride, h, w, H, W, sy, sx, bands, partial);
  } else {
    nparts = static_cast<int>(llmin((total + kCeThreads - 1) / kCeThreads, 148LL * 16));
    upsample_ce_kernel<LT><<<nparts, kCeThreads, 0, st>>>(rev, sup_mask_fg, mask_stride, total, h, w, H, W, sy, sx, partial);
  }
  ce_finalize_kernel<<<1, 256, 0, st>>>(partial, nparts, static_cast<double>(total), loss);
  return launch_status();
}
}  // namespace

extern "C" int pemp_panet_align(const float* qry_fts, long long qry_episode_stride, const float* pred,
                                const float* sup_fts, long long sup_episode_stride, const float* sup_mask_fg,
                                long long mask_stride, int B, int S, int Q, int c, int h, int w, int H, int W,
                                float scalar, float* loss, void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  return panet_align_impl<float>(qry_fts, qry_episode_stride, pred, sup_fts, sup_episode_stride, sup_mask_fg, mask_stride, B, S, Q, c, h, w,
                          H, W, scalar, loss, workspace, workspace_bytes, stream);
}

// the same loss against the uint8 label map [B*S, H, W] (1 = object) the float `sup_mask_fg` was expanded from
extern "C" int pemp_panet_align_labels(const float* qry_fts, long long qry_episode_stride, const float* pred,
                                       const float* sup_fts, long long sup_episode_stride, const uint8_t* labels,
                                       long long label_stride, int B, int S, int Q, int c, int h, int w, int H, int W,
                                       float scalar, float* loss, void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  return panet_align_impl<uint8_t>(qry_fts, qry_episode_stride, pred, sup_fts, sup_episode_stride, labels, label_stride, B, S, Q, c, h, w,
                          H, W, scalar, loss, workspace, workspace_bytes, stream);
}
