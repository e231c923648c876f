// K11  communication module of the Stage-2 backbones ("next" row 2).
//
// replaces ResNetCM.comm / VGG16CM.comm                       networks/backbones.py:208-222, 469-479
//   mask'  = max_pool2d(mask, 3, stride, 1)                                        [N, 1, h, w]
//   mean   = (x * mask').mean(hw);  max = (x * mask').max(hw)                      [N, c] each
//   feat   = linear(cat(mean.view(B, spq, c).mean(1), max.view(B, spq, c).mean(1)))   [B, n]
//   out    = feat broadcast to [N, n, h, w]  (every image of episode b gets feat[b])
// The reference makes three full-size temporaries of x (x*mask', and one copy per reduction); here x is read
// once: algorithmic bytes = N*c*hw*4 + N*(Hm*Wm + hw)*4 + N*n*hw*4.
//
// Kernels: (a) 3x3 max-pool of the mask (-inf padding, as ATen); (b) one CTA per (image, channel chunk) stages
// the pooled mask of its image in shared memory and its warps walk channel rows with coalesced loads, 16 independent
// accumulators per lane for the sum and the max of x*mask'; (c) one CTA per episode: means over the spq images,
// the n x 2c linear layer (fixed summation order) and (d) the broadcast store.
#include <math_constants.h>

#include "common.cuh"

namespace {

__global__ void comm_maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int Hm, int Wm, int h, int w,
                                    int stride) {
  const long long total = static_cast<long long>(N) * h * w;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    const long long t = i / w;
    const int y = static_cast<int>(t % h);
    const float* p = in + (t / h) * Hm * Wm;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y * stride - 1 + ky;
      if (iy < 0 || iy >= Hm) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x * stride - 1 + kx;
        if (ix < 0 || ix >= Wm) continue;
        const float v = __ldg(p + static_cast<long long>(iy) * Wm + ix);
        m = (v > m || v != v) ? v : m;            // NaN propagates, as in ATen's max_pool2d
      }
    }
    out[i] = m;
  }
}

constexpr int kCommThreads = 256;
#ifndef PEMP_COMM_U
#define PEMP_COMM_U 32
#endif
#ifndef PEMP_COMM_CTAS_PER_SM
#define PEMP_COMM_CTAS_PER_SM 8
#endif
constexpr int kCU = PEMP_COMM_U;   // independent loads per lane (ncu: the kernel waits for loads, long_scoreboard 24 warps per issue)
constexpr int kCA = 4;             // accumulator chains per lane (the batch's values fold into them round robin)

// stats[n][0][ch] = sum_x x*m / hw,  stats[n][1][ch] = max_x x*m.   kSmemMask: the image's mask is staged in smem.
template <bool kSmemMask>
__global__ void __launch_bounds__(kCommThreads)
comm_pool_kernel(const float* __restrict__ x, const float* __restrict__ mask, int c, int hw, int rows_per_cta,
                 float* __restrict__ stats) {
  extern __shared__ float msk[];
  const int n = blockIdx.y, c0 = blockIdx.x * rows_per_cta, c1 = min(c, c0 + rows_per_cta);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* mg = mask + static_cast<long long>(n) * hw;
  if (kSmemMask) {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) msk[i] = __ldg(mg + i);
    __syncthreads();
  }
  const float* mp = kSmemMask ? msk : mg;
  const float inv_hw = 1.0f / static_cast<float>(hw);
  for (int ch = c0 + warp; ch < c1; ch += kCommThreads / 32) {
    const float* row = x + (static_cast<long long>(n) * c + ch) * hw;
    float s[kCA], mx[kCA];
#pragma unroll
    for (int u = 0; u < kCA; ++u) {
      s[u] = 0.f;
      mx[u] = -CUDART_INF_F;
    }
    int i = lane;
    for (; i + (kCU - 1) * 32 < hw; i += kCU * 32) {      // full batches: kCU unpredicated loads per lane in flight
      float v[kCU];
#pragma unroll
      for (int u = 0; u < kCU; ++u) v[u] = __ldg(row + i + 32 * u);
#pragma unroll
      for (int u = 0; u < kCU; ++u) {
        const float p = v[u] * (kSmemMask ? mp[i + 32 * u] : __ldg(mp + i + 32 * u));
        s[u % kCA] += p;
        mx[u % kCA] = fmaxf(mx[u % kCA], p);
      }
    }
    if (i < hw) {     // the last, partial batch of the row, issued the same way (hw = 10201: 31 of 32 loads)
      float v[kCU];
#pragma unroll
      for (int u = 0; u < kCU; ++u) v[u] = (i + 32 * u < hw) ? __ldg(row + i + 32 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < kCU; ++u) {
        if (i + 32 * u < hw) {
          const float p = v[u] * (kSmemMask ? mp[i + 32 * u] : __ldg(mp + i + 32 * u));
          s[u % kCA] += p;
          mx[u % kCA] = fmaxf(mx[u % kCA], p);
        }
      }
    }
    float st = 0.f, mt = -CUDART_INF_F;
#pragma unroll
    for (int u = 0; u < kCA; ++u) {      // fixed order
      st += s[u];
      mt = fmaxf(mt, mx[u]);
    }
    st = warp_sum(st);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(kFull, mt, o));
    if (lane == 0) {
      stats[(static_cast<long long>(n) * 2 + 0) * c + ch] = st * inv_hw;
      stats[(static_cast<long long>(n) * 2 + 1) * c + ch] = mt;
    }
  }
}

// one CTA per episode: v = [mean over spq of means | mean over spq of maxima]; feat = W v + b
__global__ void comm_linear_kernel(const float* __restrict__ stats, const float* __restrict__ weight, const float* __restrict__ bias,
                                   int spq, int c, int n_out, float* __restrict__ feat) {
  extern __shared__ float v[];     // [2c]
  const int b = blockIdx.x;
  const float inv = 1.0f / static_cast<float>(spq);
  for (int k = threadIdx.x; k < 2 * c; k += blockDim.x) {
    const int which = k / c, ch = k - which * c;
    float a = 0.f;
    for (int s = 0; s < spq; ++s) a += stats[((static_cast<long long>(b) * spq + s) * 2 + which) * c + ch];
    v[k] = a * inv;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < n_out; j += blockDim.x / 32) {
    float a = 0.f;
    for (int k = lane; k < 2 * c; k += 32) a = fmaf(__ldg(weight + static_cast<long long>(j) * 2 * c + k), v[k], a);
    a = warp_sum(a);
    if (lane == 0) feat[b * n_out + j] = a + (bias ? __ldg(bias + j) : 0.f);
  }
}

__global__ void comm_expand_kernel(const float* __restrict__ feat, int spq, int n_out, int hw, long long total, float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long plane = i / hw;                 // n * n_out + j
    const int j = static_cast<int>(plane % n_out);
    const long long n = plane / n_out;
    out[i] = __ldg(feat + (n / spq) * n_out + j);
  }
}

}  // namespace

extern "C" size_t pemp_comm_workspace_bytes(int N, int c, int spq, int n_out) {
  if (N <= 0 || c <= 0 || spq <= 0 || n_out <= 0 || N % spq) return 0;
  return align_up(static_cast<size_t>(N) * 2 * c * sizeof(float), 256) + align_up(static_cast<size_t>(N / spq) * n_out * sizeof(float), 256);
}

extern "C" int pemp_comm_module(const float* x, const float* mask_in, int N, int c, int h, int w, int Hm, int Wm, int stride,
                                int spq, const float* weight, const float* bias, int n_out, float* mask_out, float* out,
                                void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(x && mask_in && weight && mask_out && out, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && c > 0 && h > 0 && w > 0 && Hm > 0 && Wm > 0 && spq > 0 && n_out > 0 && N % spq == 0, PEMP_E_SHAPE);
  PEMP_REQUIRE(stride >= 1 && (Hm + 2 - 3) / stride + 1 == h && (Wm + 2 - 3) / stride + 1 == w, PEMP_E_SHAPE);
  PEMP_REQUIRE(2 * c * sizeof(float) <= 48 * 1024, PEMP_E_SHAPE);
  PEMP_REQUIRE(workspace && workspace_bytes >= pemp_comm_workspace_bytes(N, c, spq, n_out), PEMP_E_WORKSPACE);
  cudaStream_t st = as_stream(stream);
  float* stats = static_cast<float*>(workspace);
  float* feat = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up(static_cast<size_t>(N) * 2 * c * sizeof(float), 256));
  const int hw = h * w;
  const long long npix = static_cast<long long>(N) * hw;
  comm_maxpool_kernel<<<static_cast<unsigned>(llmin((npix + 255) / 256, 148LL * 16)), 256, 0, st>>>(mask_in, mask_out, N, Hm, Wm, h, w, stride);
  // channel chunks: enough CTAs for ~4 per SM, at least 8 rows (one per warp) each
  int chunks = (PEMP_COMM_CTAS_PER_SM * 148 + N - 1) / N;
  if (chunks > c / 8) chunks = c / 8;
  if (chunks < 1) chunks = 1;
  int rows = (c + chunks - 1) / chunks;
  rows = (rows + 7) / 8 * 8;             // a whole number of rows per warp: 20 rows on 8 warps leaves half of them idle a third of the time
  chunks = (c + rows - 1) / rows;
  const size_t smem = static_cast<size_t>(hw) * sizeof(float);
  if (smem <= 96 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(comm_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    comm_pool_kernel<true><<<dim3(chunks, N), kCommThreads, smem, st>>>(x, mask_out, c, hw, rows, stats);
  } else {
    comm_pool_kernel<false><<<dim3(chunks, N), kCommThreads, 0, st>>>(x, mask_out, c, hw, rows, stats);
  }
  comm_linear_kernel<<<N / spq, 256, 2 * c * sizeof(float), st>>>(stats, weight, bias, spq, c, n_out, feat);
  const long long total = npix * n_out;
  comm_expand_kernel<<<static_cast<unsigned>(llmin((total + 255) / 256, 148LL * 16)), 256, 0, st>>>(feat, spq, n_out, hw, total, out);
  return launch_status();
}
