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
upsample_ce_grad_kernel(const float* __restrict__ pred, const LabelT* __restrict__ target, const float* __restrict__ weight, int h,
                        int w, int H, int W, float sy, float sx, int bands, float* __restrict__ e, float* __restrict__ part_loss,
                        float* __restrict__ part_cnt) {
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
      if (weight) {          // CELossDT: weighted sum over the sum of ALL weights, ignored pixels included (losses.py:42-43)
        const float wv = __ldg(weight + base + X);
        if (valid) acc = fmaf(wv, lse - (lb == 1 ? v1 : v0), acc);
        cnt += wv;
        e[base + X] = valid ? wv * (p1 - (lb == 1 ? 1.0f : 0.0f)) : 0.0f;
      } else {
        if (valid) {
          acc += lse - (lb == 1 ? v1 : v0);
          cnt += 1.0f;
        }
        e[base + X] = valid ? p1 - (lb == 1 ? 1.0f : 0.0f) : 0.0f;
      }
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
    part_cnt[blockIdx.x] = c2;       // unweighted: <= 16 * W, exact in fp32
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

extern "C" int pemp_upsample_ce(const float* pred, const void* target, int target_is_u8, const float* weight, int N, int h, int w,
                                int H, int W, float* loss, float* d_pred, void* workspace, size_t workspace_bytes,
                                pemp_stream_t stream) {
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
    upsample_ce_grad_kernel<uint8_t><<<pl.nparts, kLossThreads, pl.smem, st>>>(pred, static_cast<const uint8_t*>(target), weight, h, w, H, W, sy, sx,
                                                                                pl.bands, e, part_loss, part_cnt);
  } else {
    err = cudaFuncSetAttribute(upsample_ce_grad_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pl.smem));
    if (err != cudaSuccess) return static_cast<int>(err);
    upsample_ce_grad_kernel<int64_t><<<pl.nparts, kLossThreads, pl.smem, st>>>(pred, static_cast<const int64_t*>(target), weight, h, w, H, W, sy, sx,
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

// ------------------------------------------------------------------------------------------------ K14
// CELossDT.boundary2weight on the device (core/losses.py:23-40).  The reference copies the boundary map to
// the host, runs scipy's exact Euclidean distance transform per image and copies the weights back - a CPU round trip in
// every training step.  Here: (1) 3x3 boundary map; (2) per column, the vertical distance to the nearest boundary pixel
// (two scans); (3) per row, D^2(x) = min over x' of (x - x')^2 + g(x')^2 in integers - exact, so sqrt in double equals
// scipy's value - and weight = exp(-D / sigma^2) + 1 in double, stored as float.  A plane without boundary pixels gets
// scipy's distances to its virtual zero at (row -1, column 0).
namespace {

constexpr int kInfDist = 1 << 20;

template <typename LabelT>
__global__ void boundary_kernel(const LabelT* __restrict__ target, int N, int H, int W, uint8_t* __restrict__ bnd, int* __restrict__ any) {
  const long long total = static_cast<long long>(N) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const long long t = i / W;
    const int y = static_cast<int>(t % H);
    const long long n = t / H;
    const LabelT* p = target + n * H * W;
    int s = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += static_cast<long long>(p[static_cast<long long>(yy) * W + xx]) == 1;
      }
    const bool m = static_cast<long long>(p[static_cast<long long>(y) * W + x]) == 1;
    const bool b = m ? s < 9 : s > 0;
    bnd[i] = b;
    if (b) any[n] = 1;           // benign race: every writer stores 1
  }
}

// g[y][x] = distance to the nearest boundary pixel of column x (kInfDist if none).  A CTA stages 32 columns of the plane in
// shared memory with coalesced loads, 32 threads scan their column down and up there (a scan through global memory pays the
// load latency 2 H times in a row), and all threads write the result back coalesced.
constexpr int kColTile = 32;
__global__ void __launch_bounds__(256)
edt_columns_kernel(const uint8_t* __restrict__ bnd, int H, int W, int* __restrict__ g) {
  extern __shared__ int col[];             // [H][kColTile + 1]
  const int x0 = blockIdx.x * kColTile, n = blockIdx.y;
  const uint8_t* b = bnd + static_cast<long long>(n) * H * W;
  int* out = g + static_cast<long long>(n) * H * W;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int y = ty; y < H; y += 8)
    col[y * (kColTile + 1) + tx] = (x0 + tx < W) ? b[static_cast<long long>(y) * W + x0 + tx] : 0;
  __syncthreads();
  if (threadIdx.x < kColTile) {
    int* cp = col + threadIdx.x;
    int d = kInfDist;
    for (int y = 0; y < H; ++y) {
      d = cp[y * (kColTile + 1)] ? 0 : (d < kInfDist ? d + 1 : kInfDist);
      cp[y * (kColTile + 1)] = d;
    }
    d = kInfDist;
    for (int y = H - 1; y >= 0; --y) {
      const int cur = cp[y * (kColTile + 1)];
      d = cur == 0 ? 0 : (d < kInfDist ? d + 1 : kInfDist);
      if (d < cur) cp[y * (kColTile + 1)] = d;
    }
  }
  __syncthreads();
  for (int y = ty; y < H; y += 8)
    if (x0 + tx < W) out[static_cast<long long>(y) * W + x0 + tx] = col[y * (kColTile + 1) + tx];
}

// one CTA per (row, image): exact squared distance by the lower envelope search over the row, then the weight
__global__ void __launch_bounds__(256)
edt_rows_weight_kernel(const int* __restrict__ g, const int* __restrict__ any, int H, int W, double inv_sigma2,
                       float* __restrict__ weight) {
  extern __shared__ long long g2[];      // [W] squared column distances
  const int y = blockIdx.x, n = blockIdx.y;
  const int* row = g + (static_cast<long long>(n) * H + y) * W;
  const bool has = any[n] != 0;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const long long v = row[x];
    g2[x] = v >= kInfDist ? (1LL << 60) : v * v;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    long long best;
    if (!has) {
      best = static_cast<long long>(y + 1) * (y + 1) + static_cast<long long>(x) * x;
    } else {
      best = g2[x];
      // walk outwards; a candidate at horizontal distance d cannot win once d^2 >= best
      for (int d = 1; d < W; ++d) {
        const long long dd = static_cast<long long>(d) * d;
        if (dd >= best) break;
        if (x - d >= 0) best = min(best, dd + g2[x - d]);
        if (x + d < W) best = min(best, dd + g2[x + d]);
      }
    }
    const double dist = sqrt(static_cast<double>(best));
    weight[(static_cast<long long>(n) * H + y) * W + x] = static_cast<float>(exp(-dist * inv_sigma2) + 1.0);
  }
}

}  // namespace

extern "C" size_t pemp_boundary_weight_workspace_bytes(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  const size_t px = static_cast<size_t>(N) * H * W;
  return align_up(px, 256) + align_up(px * sizeof(int), 256) + align_up(static_cast<size_t>(N) * sizeof(int), 256);
}

extern "C" int pemp_boundary_weight(const void* target, int target_is_u8, int N, int H, int W, float sigma, float* weight,
                                    void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(target && weight, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && N <= 65535 && H > 0 && H <= 65535 && W > 0 && W <= 16384 && sigma > 0.f, PEMP_E_SHAPE);
  PEMP_REQUIRE(workspace && workspace_bytes >= pemp_boundary_weight_workspace_bytes(N, H, W), PEMP_E_WORKSPACE);
  cudaStream_t st = as_stream(stream);
  const size_t px = static_cast<size_t>(N) * H * W;
  char* ws = static_cast<char*>(workspace);
  uint8_t* bnd = reinterpret_cast<uint8_t*>(ws);
  int* g = reinterpret_cast<int*>(ws + align_up(px, 256));
  int* any = reinterpret_cast<int*>(ws + align_up(px, 256) + align_up(px * sizeof(int), 256));
  cudaError_t e = cudaMemsetAsync(any, 0, static_cast<size_t>(N) * sizeof(int), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  const unsigned blocks = static_cast<unsigned>(llmin((static_cast<long long>(px) + 255) / 256, 148LL * 16));
  if (target_is_u8)
    boundary_kernel<uint8_t><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(target), N, H, W, bnd, any);
  else
    boundary_kernel<int64_t><<<blocks, 256, 0, st>>>(static_cast<const int64_t*>(target), N, H, W, bnd, any);
  const size_t col_smem = static_cast<size_t>(H) * (kColTile + 1) * sizeof(int);
  PEMP_REQUIRE(col_smem <= 200 * 1024, PEMP_E_SHAPE);
  cudaError_t ec = cudaFuncSetAttribute(edt_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(col_smem));
  if (ec != cudaSuccess) return static_cast<int>(ec);
  edt_columns_kernel<<<dim3((W + kColTile - 1) / kColTile, N), 256, col_smem, st>>>(bnd, H, W, g);
  const double s2 = static_cast<double>(sigma) * static_cast<double>(sigma);
  edt_rows_weight_kernel<<<dim3(H, N), 256, static_cast<size_t>(W) * sizeof(long long), st>>>(g, any, H, W, 1.0 / s2, weight);
  return launch_status();
}
