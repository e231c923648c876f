// K3  cosine matching of query features against K = 2P prototypes (P per class).
//
// replaces compute_similarity() (+ max over prototypes, response map):
//   networks/pemp_stage1.py:214-222,233-261, pemp_stage2.py:187-194,205-233, baseline.py:121-149, panet.py:122-156
//
// Roofline: HBM.  Per query image the kernel reads c*hw floats once (5.33 MB at c=512, hw=2601) and writes
// 2*hw floats; arithmetic is (1 + K) FMA per loaded float.  Layout: qry [N, c, hw], pixel index fastest.
//
// Mapping: one CTA owns a tile of kTile pixels of one image.  Its warps split the channel range; lane l
// of every warp owns pixels {l, l+32, l+64, l+96} of the tile, so every load instruction of a warp reads
// 128 contiguous bytes of one channel row.  Normalised prototypes sit in shared memory as [c][8] so one
// channel's K values are two broadcast 128-bit loads.  Per-pixel partial sums (|q|^2 and K dot products) are
// combined across warps through shared memory in a fixed order (bit-reproducible run to run).
#include "common.cuh"

// TMA-fed persistent fast path for c = 512, P in {1, 3} (cosine_tma.cu); PEMP_E_ALIGN = not covered, nothing launched
int pemp_cosine_tma_launch(const float* qry, long long ep_stride, const float* fg, const float* bg, int N, int Bp, int hw,
                           int P, float scalar, float* sim, float* pred, int64_t* response, cudaStream_t st);
#ifndef PEMP_COS_TMA
#define PEMP_COS_TMA 1
#endif
static int g_cos_path = 0;   // diagnostic switch, see pemp_debug_cosine_path

namespace {

constexpr int kTile = 128;       // pixels per CTA
#ifndef PEMP_COS_WARPS
#define PEMP_COS_WARPS 8
#endif
constexpr int kWarps = PEMP_COS_WARPS;        // channel-splitting warps per CTA
constexpr int kPix = kTile / 32; // pixels per lane
constexpr int kMaxK = 8;         // prototype vectors per image (2P), padded row of the smem table
static_assert(kWarps >= kMaxK, "one warp per prototype column computes inv_norm: a build with fewer warps leaves it uninitialised");
constexpr float kCosEps = 1e-8f; // F.cosine_similarity eps

template <int K>
__global__ void __launch_bounds__(kWarps * 32)
cosine_match_kernel(const float* __restrict__ qry, long long ep_stride, const float* __restrict__ fg_proto, const float* __restrict__ bg_proto,
                    int Qper, int c, int hw, float scalar, float* __restrict__ sim, float* __restrict__ pred,
                    int64_t* __restrict__ response) {
  constexpr int P = K / 2;
  extern __shared__ __align__(16) float smem[];
  float* table = smem;                                  // [c][kMaxK] normalised prototypes (bg 0..P-1, fg P..2P-1)
  float* red = smem + static_cast<size_t>(c) * kMaxK;   // [kWarps][1 + K][kTile]
  __shared__ float inv_norm[kMaxK];

  const int tiles = (hw + kTile - 1) / kTile;
  const int n = blockIdx.x / tiles;
  const int x0 = (blockIdx.x - n * tiles) * kTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = n / Qper;

  // ---- prototypes -> shared, normalised by max(|p|, eps) ------------------------------------------
  for (int i = tid; i < c * P; i += blockDim.x) {
    int ch = i / P, j = i - ch * P;
    table[ch * kMaxK + j] = __ldg(bg_proto + (static_cast<long long>(b) * c + ch) * P + j);
    table[ch * kMaxK + P + j] = __ldg(fg_proto + (static_cast<long long>(b) * c + ch) * P + j);
  }
  __syncthreads();
  if (warp < K) {   // one warp per prototype vector: sum of squares over channels
    float s = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
      float v = table[ch * kMaxK + warp];
      s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) inv_norm[warp] = 1.0f / fmaxf(sqrtf(s), kCosEps);
  }
  __syncthreads();
  for (int i = tid; i < c * K; i += blockDim.x) {
    int ch = i / K, k = i - ch * K;
    table[ch * kMaxK + k] *= inv_norm[k];
  }
  __syncthreads();

  // ---- stream the channel rows ---------------------------------------------------------------------
  float nrm[kPix], dot[kPix][K];
#pragma unroll
  for (int i = 0; i < kPix; ++i) {
    nrm[i] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) dot[i][k] = 0.f;
  }
  const float* base = qry + b * ep_stride + static_cast<long long>(n - b * Qper) * c * hw + x0;
  bool ok[kPix];
#pragma unroll
  for (int i = 0; i < kPix; ++i) ok[i] = x0 + lane + 32 * i < hw;

  const int per = (c + kWarps - 1) / kWarps;
  const int c_begin = warp * per, c_end = min(c, c_begin + per);
#ifndef PEMP_COS_U
#define PEMP_COS_U 4
#endif
  constexpr int U = PEMP_COS_U;   // channel rows in flight per warp
  for (int ch = c_begin; ch < c_end; ch += U) {
    float v[U][kPix];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* row = base + static_cast<long long>(ch + u) * hw + lane;
      bool live = ch + u < c_end;
#pragma unroll
      for (int i = 0; i < kPix; ++i) v[u][i] = (live && ok[i]) ? __ldg(row + 32 * i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ch + u < c_end) {
        const float4* t4 = reinterpret_cast<const float4*>(table + (ch + u) * kMaxK);
        float4 ta = t4[0];
        float4 tb = K > 4 ? t4[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        float t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
        for (int i = 0; i < kPix; ++i) {
          nrm[i] = fmaf(v[u][i], v[u][i], nrm[i]);
#pragma unroll
          for (int k = 0; k < K; ++k) dot[i][k] = fmaf(v[u][i], t[k], dot[i][k]);
        }
      }
    }
  }

  // ---- combine warps (fixed order), finish ----------------------------------------------------------
#pragma unroll
  for (int i = 0; i < kPix; ++i) {
    red[(warp * (1 + K) + 0) * kTile + lane + 32 * i] = nrm[i];
#pragma unroll
    for (int k = 0; k < K; ++k) red[(warp * (1 + K) + 1 + k) * kTile + lane + 32 * i] = dot[i][k];
  }
  __syncthreads();
  if (tid < kTile && x0 + tid < hw) {
    float acc[1 + K];
#pragma unroll
    for (int k = 0; k <= K; ++k) acc[k] = 0.f;
    for (int wv = 0; wv < kWarps; ++wv)
#pragma unroll
      for (int k = 0; k <= K; ++k) acc[k] += red[(wv * (1 + K) + k) * kTile + tid];
    const float qinv = 1.0f / fmaxf(sqrtf(acc[0]), kCosEps);
    const int x = x0 + tid;
    float best[2];
    int arg[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      best[g] = -INFINITY;
      arg[g] = 0;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        float s = acc[1 + g * P + j] * qinv * scalar;
        if (sim) sim[((static_cast<long long>(n) * 2 + g) * P + j) * hw + x] = s;
        if (s > best[g]) { best[g] = s; arg[g] = j; }   // first maximum wins, as torch.max
      }
    }
    if (pred) {
      pred[(static_cast<long long>(n) * 2 + 0) * hw + x] = best[0];
      pred[(static_cast<long long>(n) * 2 + 1) * hw + x] = best[1];
    }
    if (response) response[static_cast<long long>(n) * hw + x] = best[1] > best[0] ? arg[1] + 3 : arg[0];
  }
}

template <int K>
int launch(const float* qry, long long ep_stride, const float* fg, const float* bg, int N, int Bp, int c, int hw, float scalar, float* sim,
           float* pred, int64_t* response, cudaStream_t st) {
  size_t smem = (static_cast<size_t>(c) * kMaxK + static_cast<size_t>(kWarps) * (1 + K) * kTile) * sizeof(float);
  if (smem > 200 * 1024) return PEMP_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(cosine_match_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid(static_cast<unsigned>((hw + kTile - 1) / kTile) * N);
  cosine_match_kernel<K><<<grid, kWarps * 32, smem, st>>>(qry, ep_stride ? ep_stride : static_cast<long long>(N / Bp) * c * hw, fg, bg, N / Bp, c, hw, scalar,
                                                          sim, pred, response);
  return launch_status();
}

}  // namespace

extern "C" int pemp_debug_cosine_path(int mode) {
  const int old = g_cos_path;
  if (mode == 0 || mode == 1) g_cos_path = mode;
  return old;
}

extern "C" int pemp_cosine_match(const float* qry, long long qry_episode_stride, const float* fg_proto, const float* bg_proto, int N, int Bp, int c,
                                 int hw, int P, float scalar, float* sim, float* pred, int64_t* response,
                                 pemp_stream_t stream) {
  PEMP_REQUIRE(qry && fg_proto && bg_proto, PEMP_E_NULL);
  PEMP_REQUIRE(sim || pred || response, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && Bp > 0 && c > 0 && hw > 0 && N % Bp == 0, PEMP_E_SHAPE);
  PEMP_REQUIRE(P >= 1 && P <= 4, PEMP_E_SHAPE);
  cudaStream_t st = as_stream(stream);
  if (PEMP_COS_TMA && c == 512 && g_cos_path != 1) {
    const int rc = pemp_cosine_tma_launch(qry, qry_episode_stride, fg_proto, bg_proto, N, Bp, hw, P, scalar, sim, pred, response, st);
    if (rc != PEMP_E_ALIGN) return rc;
  }
  switch (P) {
    case 1: return launch<2>(qry, qry_episode_stride, fg_proto, bg_proto, N, Bp, c, hw, scalar, sim, pred, response, st);
    case 2: return launch<4>(qry, qry_episode_stride, fg_proto, bg_proto, N, Bp, c, hw, scalar, sim, pred, response, st);
    case 3: return launch<6>(qry, qry_episode_stride, fg_proto, bg_proto, N, Bp, c, hw, scalar, sim, pred, response, st);
    default: return launch<8>(qry, qry_episode_stride, fg_proto, bg_proto, N, Bp, c, hw, scalar, sim, pred, response, st);
  }
}
