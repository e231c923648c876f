// K13  Training loss of the PEMP head and its gradient in one pass.
//
// replaces  F.interpolate(pred, size, mode="bilinear", align_corners=True) + CrossEntropyLoss(ignore_index=255)
//           and their autograd                                      entry/pemp_stage1.py:51,57-60, networks/pemp_stage1.py:157-162
// The reference materialises the [N, 2, H, W] logits (and, backward, their gradient and the soft-max) - about 20 bytes per
// output pixel each way.  With two classes the gradient of the loss with respect to the logits is one plane,
//   e(Y, X) = valid * (sigmoid(v1 - v0) - [label == 1]),   d v1 = e / n_valid,  d v0 = -e / n_valid,
// so the forward writes e (4 bytes per pixel) while it sums the loss, the adjoint resampler of K6 folds e back to the
// feature resolution, and a finalize kernel scales by 1 / n_valid.  Same banded structure as upsample_ce_band_kernel.
#include "common.cuh"

size_t pemp_adjoint_scratch_bytes(int planes, int h, int w);
int pemp_adjoint_launch(const float* mask, int planes, int H, int W, int h, int w, float* wt, float* msum, char* scratch,
                        cudaStream_t st);

namespace {

constexpr int kLossThreads = 256, kLossBand = 16;

template <typename LabelT>
__global__ void __launch_bounds__(kLossThreads)
upsample_ce_grad_kernel(const float* __restrict__ pred, const LabelT* __restrict__ target, int h, int w, int H, int W, float sy,
                        float sx, int bands, float* __restrict__ e, float* __restrict__ part_loss, float* __restrict__ part_cnt) {
  extern __shared__ float hrow[];                    // [nsrc][2][W]
  const int n = blockIdx.x / bands, band = blockIdx.x - n * bands;
  const int Y0 = band * kLossBand, Y1 = min(H, Y0 + kLossBand);
  const int src0 = lerp_coeff(Y0, sy, h).i0, nsrc = lerp_coeff(Y1 - 1, sy, h).i1 - src0 + 1;
  const int hw = h * w;
  const float* p0 = pred + static_cast<long long>(n) * 2 * hw + src0 * w;
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
  float acc = 0.f, cnt = 0.f;
  for (int Y = Y0; Y < Y1; ++Y) {
    const Lerp ly = lerp_coeff(Y, sy, h);
    const float* rt = hrow + (ly.i0 - src0) * 2 * W;
    const float* rb = hrow + (ly.i1 - src0) * 2 * W;
    const long long base = (static_cast<long long>(n) * H + Y) * W;
    for (int X = threadIdx.x; X < W; X += blockDim.x) {
      const float v0 = lerp2(ly.l0, rt[X], ly.l1, rb[X]);
      const float v1 = lerp2(ly.l0, rt[W + X], ly.l1, rb[W + X]);
      const long long lb = static_cast<long long>(target[base + X]);
      const bool valid = lb != 255;
      const float d = v1 - v0, t = expf(-fabsf(d));
      const float lse = fmaxf(v0, v1) + log1pf(t);
      const float p1 = d >= 0.f ? 1.0f / (1.0f + t) : t / (1.0f + t);          // sigmoid(v1 - v0)
      if (valid) {
        acc += lse - (lb == 1 ? v1 : v0);
        cnt += 1.0f;
      }
      e[base + X] = valid ? p1 - (lb == 1 ? 1.0f : 0.0f) : 0.0f;
    }
  }
  __shared__ float pa[kLossThreads / 32], pc[kLossThreads / 32];
  acc = warp_sum(acc);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) {
    pa[threadIdx.x >> 5] = acc;
    pc[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c2 = 0.f;
    for (int i = 0; i < kLossThreads / 32; ++i) {
      a += pa[i];
      c2 += pc[i];
    }
    part_loss[blockIdx.x] = a;
    part_cnt[blockIdx.x] = c2;       // <= 16 * W: exact in fp32
  }
}

// block 0 of the first launch phase: totals in double, index order; every block then scales its share of d_pred
__global__ void __launch_bounds__(256)
ce_grad_finalize_kernel(const float* __restrict__ part_loss, const float* __restrict__ part_cnt, int nparts,
                        const float* __restrict__ wt, int N, int hw, float* __restrict__ loss, float* __restrict__ d_pred) {
  __shared__ double sa[256], sc[256];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) {      // every block recomputes the (small) totals: no grid sync needed
    a += static_cast<double>(part_loss[i]);
    c += static_cast<double>(part_cnt[i]);
  }
  sa[threadIdx.x] = a;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sc[threadIdx.x] += sc[threadIdx.x + o];
    }
    __syncthreads();
  }
  const double total = sa[0], count = sc[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) loss[0] = static_cast<float>(total / count);      // NaN when nothing is valid, as torch
  if (!d_pred) return;
  const float inv = static_cast<float>(1.0 / count);
  const long long tot = static_cast<long long>(N) * hw;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < tot; i += gridDim.x * 256LL) {
    const long long n = i / hw, r = i - n * hw;
    const float g = wt[i] * inv;
    d_pred[(n * 2 + 0) * hw + r] = -g;
    d_pred[(n * 2 + 1) * hw + r] = g;
  }
}

struct LossPlan {
  int bands, nparts;
  size_t smem, off_e, off_wt, off_pl, off_pc, off_adj, total;
};
LossPlan loss_plan(int N, int h, int w, int H, int W) {
  LossPlan p;
  p.bands = (H + kLossBand - 1) / kLossBand;
  p.nparts = N * p.bands;
  const float sy = lerp_scale(h, H);
  int nsrc = static_cast<int>((kLossBand - 1) * sy) + 3;
  if (nsrc > h) nsrc = h;
  p.smem = static_cast<size_t>(nsrc) * 2 * W * sizeof(float);
  p.off_e = 0;
  p.off_wt = align_up(static_cast<size_t>(N) * H * W * sizeof(float), 256);
  p.off_pl = p.off_wt + align_up(static_cast<size_t>(N) * h * w * sizeof(float), 256);
  p.off_pc = p.off_pl + align_up(static_cast<size_t>(p.nparts) * sizeof(float), 256);
  p.off_adj = p.off_pc + align_up(static_cast<size_t>(p.nparts) * sizeof(float), 256);
  p.total = p.off_adj + pemp_adjoint_scratch_bytes(N, h, w);
  return p;
}

}  // namespace

extern "C" size_t pemp_upsample_ce_workspace_bytes(int N, int h, int w, int H, int W) {
  if (N <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return 0;
  return loss_plan(N, h, w, H, W).total;
}

extern "C" int pemp_upsample_ce(const float* pred, const void* target, int target_is_u8, int N, int h, int w, int H, int W,
                                float* loss, float* d_pred, void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(pred && target && loss, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0, PEMP_E_SHAPE);
  const LossPlan pl = loss_plan(N, h, w, H, W);
  PEMP_REQUIRE(pl.smem <= 200 * 1024 && static_cast<long long>(N) * pl.bands < (1LL << 31), PEMP_E_SHAPE);
  PEMP_REQUIRE(workspace && workspace_bytes >= pl.total, PEMP_E_WORKSPACE);
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  float* e = reinterpret_cast<float*>(ws + pl.off_e);
  float* wt = reinterpret_cast<float*>(ws + pl.off_wt);
  float* part_loss = reinterpret_cast<float*>(ws + pl.off_pl);
  float* part_cnt = reinterpret_cast<float*>(ws + pl.off_pc);
  const float sy = lerp_scale(h, H), sx = lerp_scale(w, W);
  cudaError_t err;
  if (target_is_u8) {
    err = cudaFuncSetAttribute(upsample_ce_grad_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
    if (err != cudaSuccess) return static_cast<int>(err);
    upsample_ce_grad_kernel<uint8_t><<<pl.nparts, kLossThreads, pl.smem, st>>>(pred, static_cast<const uint8_t*>(target), h, w, H, W, sy, sx,
                                                                                pl.bands, e, part_loss, part_cnt);
  } else {
    err = cudaFuncSetAttribute(upsample_ce_grad_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
    if (err != cudaSuccess) return static_cast<int>(err);
    upsample_ce_grad_kernel<int64_t><<<pl.nparts, kLossThreads, pl.smem, st>>>(pred, static_cast<const int64_t*>(target), h, w, H, W, sy, sx,
                                                                                pl.bands, e, part_loss, part_cnt);
  }
  if (d_pred) {
    const int rc = pemp_adjoint_launch(e, N, H, W, h, w, wt, nullptr, ws + pl.off_adj, st);
    if (rc != PEMP_OK) return rc;
  }
  const long long tot = static_cast<long long>(N) * h * w;
  const int blocks = d_pred ? static_cast<int>(llmin((tot + 255) / 256, 148LL * 4)) : 1;
  ce_grad_finalize_kernel<<<blocks, 256, 0, st>>>(part_loss, part_cnt, pl.nparts, wt, N, h * w, loss, d_pred);
  return launch_status();
}
