// K12  backward of the PEMP head for the training path.
//
// replaces what autograd records for                                   entry/pemp_stage1.py:57-65 (loss.backward())
//   K3  compute_similarity + max over prototypes                      pemp_stage1.py:214-215, 233-261
//   K2  meta-prototype attention                                      pemp_stage1.py:202-213
// The reference's backward re-reads the [BS, c, 2P, hw] temporaries of the forward (4 x 32 MB per shot at the PEMP size).
// Here each backward reads the feature map once and writes its gradient once: 2 * c * hw * 4 bytes per image.
//
// Both kernels share one skeleton.  A CTA owns a run of 32-pixel tiles of one image:
//   load + phase A   lane = pixel, warp = channel subset: the tile goes to shared memory while the per-pixel dot products
//                    with the [c, K] tables (prototypes / centres / gradient coefficients) accumulate in registers;
//   pixel step       one warp: cross-warp sums, arg-max / softmax and their derivatives -> per-pixel coefficients;
//   phase B1         lane = pixel: the gradient tile is a rank-2K combination of table rows (plus, for K3, the tile
//                    itself) -> coalesced 128-byte stores;
//   phase B2         thread = channel: the per-channel sums over pixels (gradients of prototypes / centres) accumulate in
//                    registers across the CTA's tiles (tile rows are padded to 33 floats: conflict free both ways).
// Per-CTA partial sums go to the workspace and a finalize kernel adds them in a fixed order (no float atomics).
//
// Math.  K3: sim_gk = scalar * (q . pn_gk) / max(|q|, eps), pred_g = max_k sim_gk (first maximum wins)
//   dq  = sum_g gp_g * scalar * (pn_gk* / |q| - (q . pn_gk*) q / |q|^3)          (second term 0 when |q| < eps)
//   dpn_gk = sum_x [k = k*] gp_g * scalar * q / |q|;   dproto = (dpn - pn (pn . dpn)) / |proto|   (dpn / eps if clamped)
// K2: l_k = 2 f . ctr_k - |ctr_k|^2 (the |f|^2 term cancels in the softmax; evaluated as differences to the first prototype of the
//   group, l_k - l_g0 = 2 f . (ctr_k - ctr_g0) - (|ctr_k|^2 - |ctr_g0|^2), which the soft-max cannot tell apart), sigma = softmax over the group,
//   a_k = m_g sigma_k, centre_k = sum_x f a_k / den_k, den_k = sum_x a_k + eps, with A_k = g_centre_k / den_k:
//   da_k = f . A_k - A_k . centre_k;  dl_k = sigma_k (m da_k - sum_j sigma_j m da_j)
//   df = sum_k (a_k A_k + 2 dl_k ctr_k);   dctr_k = sum_x 2 dl_k f - ctr_k sum_x 2 dl_k
#include <math_constants.h>

#include "common.cuh"

// train_mma.cu: K2 backward with the per-tile products on the warp-level tensor path (P = 3, c in {128, 256, 512, 1024})
size_t pemp_mpa_bwd_mma_smem(int c);
bool pemp_mpa_bwd_mma_shape(int c, int p, int hw);
int pemp_mpa_bwd_mma_tiles(int hw);
size_t pemp_mpa_bwd_mma_table_bytes(int N, int c);
int pemp_mpa_bwd_mma_table_ld();
int pemp_mpa_bwd_mma_rows_per_warp(int c);
int pemp_mpa_bwd_mma_launch(const float* fts, long long ep, int B, int S, const float* ctr, const float* coef, const float* beta,
                            const float* fg, const float* bg, long long mask_stride, int c, int hw, int chunks, float* tabg,
                            float* dfts, long long d_ep, float* part, float* img_part, cudaStream_t st);
// train_mma_cos.cu: K3 backward on the same structure (P = 3, c in {256, 512})
bool pemp_cos_bwd_mma_shape(int c, int P, int hw);
size_t pemp_cos_bwd_mma_smem(int c);
size_t pemp_cos_bwd_mma_table_bytes(int Bp, int c);
int pemp_cos_bwd_mma_tiles(int hw);
int pemp_cos_bwd_mma_launch(bool dense, const float* qry, long long ep, int Bp, int Q, const float* pn, const float* g, int c, int hw,
                            int chunks, float scalar, float* tabg, float* dq, long long d_ep, float* part, cudaStream_t st);
static int g_bwd_path = 0;   // diagnostic switch, see pemp_debug_bwd_path

namespace {

#ifndef PEMP_BWD_THREADS
#define PEMP_BWD_THREADS 256
#endif
#ifndef PEMP_BWD_UNROLL
#define PEMP_BWD_UNROLL 8
#endif
constexpr int kBT = PEMP_BWD_THREADS;   // threads per CTA
constexpr int kBW = kBT / 32;           // warps
constexpr int kLd = 33;                 // padded tile row
constexpr int kMaxCPT = (1024 + kBT - 1) / kBT;   // channels per thread in phase B2  =>  c <= 1024
constexpr int kUn = PEMP_BWD_UNROLL;    // feature loads in flight per lane in phase A
constexpr float kCosEps = 1e-8f;   // F.cosine_similarity's eps

__device__ __forceinline__ float block_sum(float v, float* scratch) {   // scratch: kBW floats; every thread gets the sum
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < kBW; ++w) s += scratch[w];
  return s;
}

// ------------------------------------------------------------------------------------------------ K3 backward
// pn [Bp][c][K] normalised prototypes (k < P: background, k >= P: foreground = the channel order of pred), nrm [Bp][K]
__global__ void __launch_bounds__(kBT)
proto_norm_kernel(const float* __restrict__ fg_proto, const float* __restrict__ bg_proto, int c, int P, float* __restrict__ pn,
                  float* __restrict__ nrm) {
  __shared__ float scratch[kBW];
  const int b = blockIdx.x, K = 2 * P;
  for (int k = 0; k < K; ++k) {
    const float* src = (k < P ? bg_proto : fg_proto) + static_cast<long long>(b) * c * P + (k < P ? k : k - P);
    float s = 0.f;
    for (int ch = threadIdx.x; ch < c; ch += kBT) {
      const float v = __ldg(src + static_cast<long long>(ch) * P);
      s = fmaf(v, v, s);
    }
    const float n = sqrtf(block_sum(s, scratch));
    const float inv = 1.0f / fmaxf(n, kCosEps);
    for (int ch = threadIdx.x; ch < c; ch += kBT)
      pn[(static_cast<long long>(b) * c + ch) * K + k] = __ldg(src + static_cast<long long>(ch) * P) * inv;
    if (threadIdx.x == 0) nrm[b * K + k] = n;
  }
}

// Table rows are K floats (K even => 8-byte aligned): a row is K/2 LDS.64 whose column pairs feed the packed FFMA2 (two
// fp32 FMAs per issue slot) - the phases below are issue bound, not HBM bound, at K = 6.  (Rows padded to 8 floats would
// be two LDS.128, but push the CTA past half of the SM's shared memory at c = 512.)
template <int K>
struct Pad {
  static constexpr int KP = K;
  static constexpr int H = KP / 2;
};
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// The tile buffer is single (two CTAs per SM overlap instead), so the next tile's rows are pulled into L2 while this tile
// is being worked on: phase A then waits for L2, not for HBM (long-scoreboard stalls were the top stall reason).
#ifndef PEMP_BWD_PREFETCH
#define PEMP_BWD_PREFETCH 1
#endif
__device__ __forceinline__ void prefetch_tile(const float* src, int c, int hw, int x0) {
  if (!PEMP_BWD_PREFETCH || x0 >= hw) return;
  const int x1 = min(x0 + 31, hw - 1);
  for (int ch = threadIdx.x; ch < c; ch += kBT) {
    const float* row = src + static_cast<long long>(ch) * hw;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(row + x0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(row + x1));
  }
}
template <int KP>
__device__ __forceinline__ void load_row(const float* row, float2 (&r)[KP / 2]) {
#pragma unroll
  for (int j = 0; j < KP / 2; ++j) r[j] = reinterpret_cast<const float2*>(row)[j];
}

// Phase A / B1 of cosine_bwd_kernel; FAST as in mpa_bwd_phase_a below.
template <int K, bool FAST>
__device__ __forceinline__ void cos_bwd_phase_a(const float* __restrict__ src, int c, int hw, int x, bool inb, int warp, int lane,
                                                float* tile, const float* tab, float2 (&acc)[K / 2], float& nacc) {
  constexpr int H = K / 2;
  const float* gp = src + static_cast<long long>(warp) * hw + x;
  const long long gstep = static_cast<long long>(kBW) * hw;
  float* tp = tile + warp * kLd + lane;
  const float* rp = tab + warp * K;
  for (int ch0 = warp; ch0 < c; ch0 += kBW * kUn) {
    float vv[kUn];
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      if (FAST) {
        vv[u] = __ldg(gp + u * gstep);
      } else {
        vv[u] = (inb && ch0 + u * kBW < c) ? __ldg(gp + u * gstep) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      if (FAST || ch0 + u * kBW < c) {
        const float v = vv[u];
        tp[u * kBW * kLd] = v;
        nacc = fmaf(v, v, nacc);
        float2 r[H];
        load_row<K>(rp + u * kBW * K, r);
        const float2 v2 = make_float2(v, v);
#pragma unroll
        for (int j = 0; j < H; ++j) acc[j] = ffma2(v2, r[j], acc[j]);
      }
    }
    gp += kUn * gstep;
    tp += kUn * kBW * kLd;
    rp += kUn * kBW * K;
  }
}

template <int K, bool FAST>
__device__ __forceinline__ void cos_bwd_phase_b1(float* __restrict__ dst, int c, int hw, bool inb, int warp, int lane,
                                                 const float* tile, const float* tab, float c0, float c1, float tv, int s0, int s1) {
  const long long gstep = static_cast<long long>(kBW) * hw;
  dst += static_cast<long long>(warp) * hw;
  const float* tp = tile + warp * kLd + lane;
  const float* rp = tab + warp * K;
#pragma unroll 4
  for (int ch = warp; ch < c; ch += kBW) {
    const float val = fmaf(c0, rp[s0], fmaf(c1, rp[s1], tv * *tp));
    if (FAST || inb) *dst = val;
    dst += gstep;
    tp += kBW * kLd;
    rp += kBW * K;
  }
}

// DENSE = false: g_pred [N][2][hw], gradient of pred = max over the prototypes of a class (first maximum wins, recomputed here).
// DENSE = true : g_pred is g_sim [N][K][hw], the gradient of every per-prototype map of `compute_similarity`
//                (pemp_stage1.py:233-261; k < P background, k >= P foreground = its channel order): no arg-max, the gradient
//                tile is a rank-K combination of the table rows.
template <int K, bool DENSE>
__global__ void __launch_bounds__(kBT, 2)
cosine_bwd_kernel(const float* __restrict__ qry, long long ep_stride, int Q, const float* __restrict__ pn,
                  const float* __restrict__ g_pred, int c, int hw, int ntiles, float scalar, float* __restrict__ dq,
                  long long d_ep_stride, float* __restrict__ part) {
  constexpr int P = K / 2, KP = Pad<K>::KP, H = Pad<K>::H, NA = KP + 1;
  extern __shared__ __align__(16) float sm[];
  float* tab = sm;                           // [c][KP]
  float* wts = tab + c * KP;                 // [32][KP]  phase-B2 weights
  float* tile = wts + 32 * KP;               // [c][33]
  float* red = tile + c * kLd;               // [kBW][NA][32]
  float* sum = red + kBW * NA * 32;          // [NA][32]
  float* cg = sum + NA * 32;                 // [2][32]   coefficient of the selected prototype
  float* tq = cg + 64;                       // [32]      coefficient of q
  int* sel = reinterpret_cast<int*>(tq + 32);   // [2][32] selected table column
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.y, b = n / Q, qi = n - b * Q;
  const float* src = qry + static_cast<long long>(b) * ep_stride + static_cast<long long>(qi) * c * hw;
  for (int i = tid; i < c * KP; i += kBT) {
    const int ch = i / KP, k = i - ch * KP;
    tab[i] = k < K ? __ldg(pn + (static_cast<long long>(b) * c + ch) * K + k) : 0.f;
  }
  for (int i = tid; i < 32 * KP; i += kBT) wts[i] = 0.f;
  float2 accB[kMaxCPT][H];
#pragma unroll
  for (int i = 0; i < kMaxCPT; ++i)
#pragma unroll
    for (int k = 0; k < H; ++k) accB[i][k] = make_float2(0.f, 0.f);
  const int t0 = static_cast<int>(static_cast<long long>(ntiles) * blockIdx.x / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(ntiles) * (blockIdx.x + 1) / gridDim.x);
  __syncthreads();
  for (int t = t0; t < t1; ++t) {
    const int x = t * 32 + lane;
    const bool inb = x < hw;
    float2 acc[H];
    float nacc = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = make_float2(0.f, 0.f);
    const bool fast = (t * 32 + 32 <= hw) && (c % (kBW * kUn) == 0);
    if (fast)
      cos_bwd_phase_a<K, true>(src, c, hw, x, inb, warp, lane, tile, tab, acc, nacc);
    else
      cos_bwd_phase_a<K, false>(src, c, hw, x, inb, warp, lane, tile, tab, acc, nacc);
#pragma unroll
    for (int j = 0; j < H; ++j) {
      red[(warp * NA + 2 * j) * 32 + lane] = acc[j].x;
      red[(warp * NA + 2 * j + 1) * 32 + lane] = acc[j].y;
    }
    red[(warp * NA + KP) * 32 + lane] = nacc;
    if (t + 1 < t1) prefetch_tile(src, c, hw, (t + 1) * 32);
    __syncthreads();
    for (int i = tid; i < NA * 32; i += kBT) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kBW; ++w) s += red[w * NA * 32 + i];
      sum[i] = s;
    }
    __syncthreads();
    if (tid < 32) {
      const float nq = sqrtf(sum[KP * 32 + lane]);
      const float invq = 1.0f / fmaxf(nq, kCosEps);
      float tsum = 0.f;
      if (DENSE) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float gk = inb ? __ldg(g_pred + (static_cast<long long>(n) * K + k) * hw + x) * scalar : 0.f;
          wts[lane * KP + k] = gk * invq;
          tsum = fmaf(gk, sum[k * 32 + lane], tsum);
        }
      }
#pragma unroll
      for (int g = 0; g < 2 && !DENSE; ++g) {
        int best = 0;
        float bv = sum[(g * P) * 32 + lane];
#pragma unroll
        for (int k = 1; k < P; ++k) {
          const float v = sum[(g * P + k) * 32 + lane];
          if (v > bv) {
            bv = v;
            best = k;
          }
        }
        const float gp = inb ? __ldg(g_pred + (static_cast<long long>(n) * 2 + g) * hw + x) * scalar : 0.f;
        const float co = gp * invq;
        cg[g * 32 + lane] = co;
        sel[g * 32 + lane] = g * P + best;
        tsum = fmaf(gp, bv, tsum);
#pragma unroll
        for (int k = 0; k < P; ++k) wts[lane * KP + g * P + k] = (k == best) ? co : 0.f;
      }
      tq[lane] = (nq < kCosEps) ? 0.f : tsum * invq * invq * invq;
    }
    __syncthreads();
    if (DENSE) {   // B1: dq tile = sum_k a_k pn_k - t q
      float a[K];
#pragma unroll
      for (int k = 0; k < K; ++k) a[k] = wts[lane * KP + k];
      const float tv = -tq[lane];
      float* dst = dq + static_cast<long long>(b) * d_ep_stride + static_cast<long long>(qi) * c * hw + x +
                   static_cast<long long>(warp) * hw;
      const float* tp = tile + warp * kLd + lane;
      const float* rp = tab + warp * K;
      const long long gstep = static_cast<long long>(kBW) * hw;
      for (int ch = warp; ch < c; ch += kBW) {
        float val = tv * *tp;
#pragma unroll
        for (int k = 0; k < K; ++k) val = fmaf(a[k], rp[k], val);
        if (inb) *dst = val;
        dst += gstep;
        tp += kBW * kLd;
        rp += kBW * K;
      }
    } else {   // B1: dq tile
      const float c0 = cg[lane], c1 = cg[32 + lane], tv = -tq[lane];
      const int s0 = sel[lane], s1 = sel[32 + lane];
      float* dst = dq + static_cast<long long>(b) * d_ep_stride + static_cast<long long>(qi) * c * hw + x;
      if (fast)
        cos_bwd_phase_b1<K, true>(dst, c, hw, inb, warp, lane, tile, tab, c0, c1, tv, s0, s1);
      else
        cos_bwd_phase_b1<K, false>(dst, c, hw, inb, warp, lane, tile, tab, c0, c1, tv, s0, s1);
    }
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i) {   // B2: per-channel sums for the prototype gradients
      const int ch = tid + i * kBT;
      if (ch < c) {
#pragma unroll 4
        for (int xx = 0; xx < 32; ++xx) {
          const float v = tile[ch * kLd + xx];
          float2 r[H];
          load_row<KP>(wts + xx * KP, r);
          const float2 v2 = make_float2(v, v);
#pragma unroll
          for (int k = 0; k < H; ++k) accB[i][k] = ffma2(r[k], v2, accB[i][k]);
        }
      }
    }
    __syncthreads();
  }
  float* dstp = part + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * c * K;
#pragma unroll
  for (int i = 0; i < kMaxCPT; ++i) {
    const int ch = tid + i * kBT;
    if (ch < c) {
#pragma unroll
      for (int k = 0; k < K; ++k) dstp[ch * K + k] = (k & 1) ? accB[i][k / 2].y : accB[i][k / 2].x;
    }
  }
}

// one CTA per prototype set: add the partials of its Q query maps, then the backward of the normalisation.  A thread reads the
// whole K-float row of its channel from every partial in ONE pass with several loads in flight (the first version made a pass
// per column with a single 4-byte load in flight: 20 us for 16 query maps, a fifth of the whole backward).
__global__ void __launch_bounds__(kBT)
cosine_bwd_finalize_kernel(const float* __restrict__ part, const float* __restrict__ pn, const float* __restrict__ nrm, int Q,
                           int chunks, int c, int P, float* __restrict__ d_fg, float* __restrict__ d_bg) {
  __shared__ float scratch[kBW];
  const int b = blockIdx.x, K = 2 * P;
  float d[kMaxCPT][8];
#pragma unroll
  for (int i = 0; i < kMaxCPT; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) d[i][k] = 0.f;
    const int ch = threadIdx.x + i * kBT;
    if (ch < c) {
      const float* row = part + (static_cast<long long>(b) * Q * chunks * c + ch) * K;
#pragma unroll 4
      for (int j = 0; j < Q * chunks; ++j) {       // in index order: the same sums as before, bit for bit
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < K) d[i][k] += __ldg(row + k);
        row += static_cast<long long>(c) * K;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k >= K) break;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i) {
      const int ch = threadIdx.x + i * kBT;
      if (ch < c) dot = fmaf(pn[(static_cast<long long>(b) * c + ch) * K + k], d[i][k], dot);
    }
    dot = block_sum(dot, scratch);
    const float nv = nrm[b * K + k];
    float* out = (k < P ? d_bg : d_fg) + static_cast<long long>(b) * c * P + (k < P ? k : k - P);
#pragma unroll
    for (int i = 0; i < kMaxCPT; ++i) {
      const int ch = threadIdx.x + i * kBT;
      if (ch < c) {
        const float pv = pn[(static_cast<long long>(b) * c + ch) * K + k];
        out[static_cast<long long>(ch) * P] = (nv < kCosEps) ? d[i][k] / kCosEps : (d[i][k] - pv * dot) / nv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ K2 backward
// coef [BS][c][K] = g_centre / (S * den); beta [BS][2K] = { -sum_c coef * centre, |ctr_k|^2 - |ctr_g0|^2 }.   Columns k < P:
// foreground group.  The squared-norm differences are a bias common to every pixel's logit, so they are summed in double.
__global__ void __launch_bounds__(kBT)
mpa_bwd_prepare_kernel(const float* __restrict__ g_fg, const float* __restrict__ g_bg, const float* __restrict__ shot_centre,
                       const float* __restrict__ shot_den, const float* __restrict__ ctr, int N, int S, int c, int P,
                       float* __restrict__ coef, float* __restrict__ beta, float* __restrict__ tabg, int tab_ld, int CW) {
  // One pass over the channels with all 2P columns at once and ONE block reduction (the first version ran 2P block sums in a
  // row and had 2P threads of EVERY CTA add the c squared-norm terms serially in double: 49 us of a 430-us backward at 80
  // images).  Block N computes the squared-norm differences, which do not depend on the image, once for the launch.
  __shared__ double dscr[kBW][8];
  const int n = blockIdx.x, K = 2 * P, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.0;
  if (n == N) {
    // |ctr_k|^2 - |ctr_g0|^2 with g0 the first prototype of k's group: the soft-max of a group only sees differences of logits,
    // and the backward kernels work with ctr_k - ctr_g0 throughout, as the forward does
    for (int ch = threadIdx.x; ch < c; ch += kBT) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < K) {
          const double v = static_cast<double>(__ldg(ctr + ch * K + k)), v0 = static_cast<double>(__ldg(ctr + ch * K + (k / P) * P));
          s[k] += (v - v0) * (v + v0);
        }
      }
    }
  } else {
    const int b = n / S;
    for (int ch = threadIdx.x; ch < c; ch += kBT) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < K) {
          const float* g = (k < P ? g_fg : g_bg) + static_cast<long long>(b) * c * P + (k < P ? k : k - P);
          const float inv = 1.0f / (static_cast<float>(S) * __ldg(shot_den + n * K + k));
          const float a = __ldg(g + static_cast<long long>(ch) * P) * inv;
          coef[(static_cast<long long>(n) * c + ch) * K + k] = a;
          s[k] += static_cast<double>(a * __ldg(shot_centre + (static_cast<long long>(n) * c + ch) * K + k));
        }
      }
      if (tabg) {
        // the image's table for the tensor-path kernel (train_mma.cu), in its row order: row R = w CW + r <-> channel
        // 4 (CW q + r) + e with w = e + 4 q, CW = rows per product warp:  { coef[ch][0..K) | ctr_k - ctr_g0 of the non-first
        // prototypes (exact in double, one rounding) | 0 .. }
        const int g4 = ch >> 2, R = ((ch & 3) + 4 * (g4 / CW)) * CW + g4 % CW;
        float* row = tabg + (static_cast<long long>(n) * c + R) * tab_ld;
        int col = 0;
        for (int k = 0; k < K; ++k) row[col++] = coef[(static_cast<long long>(n) * c + ch) * K + k];
        for (int k = 0; k < K; ++k) {
          if (k % P == 0) continue;
          row[col++] = static_cast<float>(static_cast<double>(__ldg(ctr + ch * K + k)) - static_cast<double>(__ldg(ctr + ch * K + (k / P) * P)));
        }
        for (; col < tab_ld; ++col) row[col] = 0.f;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(kFull, s[k], o);
    if (lane == 0) dscr[warp][k] = s[k];
  }
  __syncthreads();
  if (n == N) {
    for (int i = threadIdx.x; i < N * K; i += kBT) {
      const int k = i % K;
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kBW; ++w) t += dscr[w][k];
      beta[(i / K) * 2 * K + K + k] = static_cast<float>(t);
    }
  } else if (threadIdx.x < K) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kBW; ++w) t += dscr[w][threadIdx.x];
    beta[n * 2 * K + threadIdx.x] = static_cast<float>(-t);
  }
}

// Phase A / B1 of mpa_bwd_kernel.  FAST: the tile lies inside the row (no per-element bounds predicate) and c is a multiple of
// the per-iteration channel count (no channel predicate) - the common case runs without the branch / predicate scaffolding,
// which was ~40 % of the instructions of the first version.  Pointers advance by constants instead of being recomputed.
// One table row per channel: { ctr[ch][0..K), coef[ch][0..K) } = 2K floats = K/2 LDS.128 (the first version kept two tables and
// read a row as 2 x K/2 LDS.64: the table loads were a third of the instructions of phase A and B1).
template <int K>
__device__ __forceinline__ void load_row2(const float* row, float2 (&rc)[K / 2], float2 (&ra)[K / 2]) {
  float4 q[K / 2];
#pragma unroll
  for (int j = 0; j < K / 2; ++j) q[j] = reinterpret_cast<const float4*>(row)[j];
#pragma unroll
  for (int j = 0; j < K / 2; ++j) {
    const float4 lo = q[j / 2], hi = q[(K / 2 + j) / 2];
    rc[j] = (j & 1) ? make_float2(lo.z, lo.w) : make_float2(lo.x, lo.y);
    ra[j] = ((K / 2 + j) & 1) ? make_float2(hi.z, hi.w) : make_float2(hi.x, hi.y);
  }
}

template <int K, bool FAST>
__device__ __forceinline__ void mpa_bwd_phase_a(const float* __restrict__ src, int c, int hw, int x, bool inb, int warp, int lane,
                                                float* tile, const float* tab, float2 (&accC)[K / 2], float2 (&accA)[K / 2]) {
  constexpr int H = K / 2;
  const float* gp = src + static_cast<long long>(warp) * hw + x;
  const long long gstep = static_cast<long long>(kBW) * hw;
  float* tp = tile + warp * kLd + lane;
  const float* rp = tab + warp * 2 * K;
  for (int ch0 = warp; ch0 < c; ch0 += kBW * kUn) {
    float vv[kUn];
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      if (FAST) {
        vv[u] = __ldg(gp + u * gstep);
      } else {
        vv[u] = (inb && ch0 + u * kBW < c) ? __ldg(gp + u * gstep) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      if (FAST || ch0 + u * kBW < c) {
        const float v = vv[u];
        tp[u * kBW * kLd] = v;
        float2 rc[H], ra[H];
        load_row2<K>(rp + u * kBW * 2 * K, rc, ra);
        const float2 v2 = make_float2(v, v);
#pragma unroll
        for (int j = 0; j < H; ++j) {
          accC[j] = ffma2(v2, rc[j], accC[j]);
          accA[j] = ffma2(v2, ra[j], accA[j]);
        }
      }
    }
    gp += kUn * gstep;
    tp += kUn * kBW * kLd;
    rp += kUn * kBW * 2 * K;
  }
}

template <int K, bool FAST>
__device__ __forceinline__ void mpa_bwd_phase_b1(float* __restrict__ dst, int c, int hw, bool inb, int warp, const float* tab,
                                                 const float2 (&a)[K / 2], const float2 (&d)[K / 2]) {
  constexpr int H = K / 2;
  const long long gstep = static_cast<long long>(kBW) * hw;
  dst += static_cast<long long>(warp) * hw;
  const float* rp = tab + warp * 2 * K;
#pragma unroll 4
  for (int ch = warp; ch < c; ch += kBW) {
    float2 rc[H], ra[H];
    load_row2<K>(rp, rc, ra);
    float2 v = make_float2(0.f, 0.f);                 // one accumulator pair: a single horizontal add per element
#pragma unroll
    for (int j = 0; j < H; ++j) {
      v = ffma2(a[j], ra[j], v);
      v = ffma2(d[j], rc[j], v);
    }
    const float val = v.x + v.y;
    if (FAST || inb) *dst = val;
    dst += gstep;
    rp += kBW * 2 * K;
  }
}

template <int K>
__global__ void __launch_bounds__(kBT, 2)
mpa_bwd_kernel(const float* __restrict__ fts, long long ep_stride, int S, const float* __restrict__ ctr,
               const float* __restrict__ coef, const float* __restrict__ beta, const float* __restrict__ fg,
               const float* __restrict__ bg, long long mask_stride, int c, int hw, int ntiles, float* __restrict__ dfts,
               long long d_ep_stride, float* __restrict__ part) {
  constexpr int P = K / 2, KP = Pad<K>::KP, H = Pad<K>::H, NA = 2 * KP;
  extern __shared__ __align__(16) float sm[];
  float* tab = sm;                         // [c][2 KP]  { centres | gradient coefficients of this image }
  float* av = tab + 2 * c * KP;            // [32][KP] a_k
  float* dv = av + 32 * KP;                // [32][KP] 2 dl_k
  float* tile = dv + 32 * KP;              // [c][33]
  float* red = tile + c * kLd;             // [kBW][NA][32]
  float* sum = red + kBW * NA * 32;        // [NA][32]
  float* konst = sum + NA * 32;            // [K] |ctr_k|^2, [K] beta
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.y, b = n / S, si = n - b * S;
  const float* src = fts + static_cast<long long>(b) * ep_stride + static_cast<long long>(si) * c * hw;
  for (int i = tid; i < c * KP; i += kBT) {
    const int ch = i / KP, k = i - ch * KP;
    // centre columns as differences to the first prototype of their group (exact in double, one rounding): the logits
    // l_k - l_g0 = 2 f.(ctr_k - ctr_g0) - (|ctr_k|^2 - |ctr_g0|^2) are O(10) instead of O(|f.ctr|) ~ 300, and since the
    // soft-max gradient sums to zero over a group, sum_k dl_k ctr_k = sum_k dl_k (ctr_k - ctr_g0): phase B1 needs no more
    tab[ch * 2 * KP + k] = k < K ? static_cast<float>(static_cast<double>(__ldg(ctr + ch * K + k)) -
                                                      static_cast<double>(__ldg(ctr + ch * K + (k / P) * P))) : 0.f;
    tab[ch * 2 * KP + KP + k] = k < K ? __ldg(coef + (static_cast<long long>(n) * c + ch) * K + k) : 0.f;
  }
  for (int i = tid; i < 32 * KP; i += kBT) {
    av[i] = 0.f;
    dv[i] = 0.f;
  }
  if (tid < K) {
    konst[tid] = __ldg(beta + n * 2 * K + K + tid);
    konst[K + tid] = __ldg(beta + n * 2 * K + tid);
  }
  float2 accB[kMaxCPT][H];
#pragma unroll
  for (int i = 0; i < kMaxCPT; ++i)
#pragma unroll
    for (int k = 0; k < H; ++k) accB[i][k] = make_float2(0.f, 0.f);
  float dsum = 0.f;
  const int t0 = static_cast<int>(static_cast<long long>(ntiles) * blockIdx.x / gridDim.x);
  const int t1 = static_cast<int>(static_cast<long long>(ntiles) * (blockIdx.x + 1) / gridDim.x);
  __syncthreads();
  for (int t = t0; t < t1; ++t) {
    const int x = t * 32 + lane;
    const bool inb = x < hw;
    float2 accC[H], accA[H];
#pragma unroll
    for (int j = 0; j < H; ++j) accC[j] = accA[j] = make_float2(0.f, 0.f);
    const bool fast = (t * 32 + 32 <= hw) && (c % (kBW * kUn) == 0);
    if (fast)
      mpa_bwd_phase_a<K, true>(src, c, hw, x, inb, warp, lane, tile, tab, accC, accA);
    else
      mpa_bwd_phase_a<K, false>(src, c, hw, x, inb, warp, lane, tile, tab, accC, accA);
#pragma unroll
    for (int j = 0; j < H; ++j) {
      red[(warp * NA + 2 * j) * 32 + lane] = accC[j].x;
      red[(warp * NA + 2 * j + 1) * 32 + lane] = accC[j].y;
      red[(warp * NA + KP + 2 * j) * 32 + lane] = accA[j].x;
      red[(warp * NA + KP + 2 * j + 1) * 32 + lane] = accA[j].y;
    }
    if (t + 1 < t1) prefetch_tile(src, c, hw, (t + 1) * 32);
    __syncthreads();
    if (tid < 64) {   // pixel step: warp g handles group g; every thread adds the eight warps' partials it needs itself
      {
        const int g = warp;
        const float m = inb ? __ldg((g == 0 ? fg : bg) + static_cast<long long>(n) * mask_stride + x) : 0.f;
        float sc[P], sa[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
          float u = 0.f, v = 0.f;
#pragma unroll
          for (int w = 0; w < kBW; ++w) {      // fixed order
            u += red[(w * NA + g * P + k) * 32 + lane];
            v += red[(w * NA + KP + g * P + k) * 32 + lane];
          }
          sc[k] = u;
          sa[k] = v;
        }
        float l[P], mx = -CUDART_INF_F;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          l[k] = fmaf(2.0f, sc[k], -konst[g * P + k]);
          mx = fmaxf(mx, l[k]);
        }
        float z = 0.f;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          l[k] = expf(l[k] - mx);
          z += l[k];
        }
        const float iz = 1.0f / z;
        float ds[P], dot = 0.f;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          l[k] *= iz;                                                                // sigma_k
          ds[k] = m * (sa[k] + konst[K + g * P + k]);                                // d sigma_k
          dot = fmaf(l[k], ds[k], dot);
        }
#pragma unroll
        for (int k = 0; k < P; ++k) {
          av[lane * KP + g * P + k] = m * l[k];
          dv[lane * KP + g * P + k] = 2.0f * l[k] * (ds[k] - dot);
        }
      }
    }
    __syncthreads();
    {   // B1: df tile = sum_k a_k A_k + 2 dl_k ctr_k
      float2 a[H], d[H];
      load_row<KP>(av + lane * KP, a);
      load_row<KP>(dv + lane * KP, d);
      float* dst = dfts + static_cast<long long>(b) * d_ep_stride + static_cast<long long>(si) * c * hw + x;
      if (fast)
        mpa_bwd_phase_b1<K, true>(dst, c, hw, inb, warp, tab, a, d);
      else
        mpa_bwd_phase_b1<K, false>(dst, c, hw, inb, warp, tab, a, d);
    }
    {   // B2: sum_x 2 dl_k f - the weight row of a pixel is loaded once for all channels of the thread
#pragma unroll 4
      for (int xx = 0; xx < 32; ++xx) {
        float2 r[H];
        load_row<KP>(dv + xx * KP, r);
#pragma unroll
        for (int i = 0; i < kMaxCPT; ++i) {
          const int ch = tid + i * kBT;
          if (ch < c) {
            const float v = tile[ch * kLd + xx];
            const float2 v2 = make_float2(v, v);
#pragma unroll
            for (int k = 0; k < H; ++k) accB[i][k] = ffma2(r[k], v2, accB[i][k]);
          }
        }
      }
    }
    if (tid < K) {
      for (int xx = 0; xx < 32; ++xx) dsum += dv[xx * KP + tid];
    }
    __syncthreads();
  }
  float* dstp = part + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * (c + 1) * K;
#pragma unroll
  for (int i = 0; i < kMaxCPT; ++i) {
    const int ch = tid + i * kBT;
    if (ch < c) {
#pragma unroll
      for (int k = 0; k < K; ++k) dstp[ch * K + k] = (k & 1) ? accB[i][k / 2].y : accB[i][k / 2].x;
    }
  }
  if (tid < K) dstp[c * K + tid] = dsum;
}

// d_ctr[ch][k] = sum over every (image, chunk) partial, in index order, minus ctr * sum of the 2 dl_k column.
// blockDim = (32 outputs, kFinRows): row r adds the partials p = r, r + kFinRows, ... (double), then the rows are added
// in order - deterministic, and 16 x more loads in flight than one thread per output.
constexpr int kFinRows = 32;     // (16 rows with 4 loads in flight: 29 us for 880 partials - the loop is a chain of DRAM round trips)
__global__ void __launch_bounds__(32 * kFinRows)
mpa_bwd_finalize_kernel(const float* __restrict__ part, long long nparts, int c, int K, const float* __restrict__ ctr,
                        float* __restrict__ d_ctr) {
  __shared__ double ss[kFinRows][32], sd[kFinRows][32];
  const int i = blockIdx.x * 32 + threadIdx.x, r = threadIdx.y;
  const bool live = i < c * K;
  const int k = live ? i % K : 0;
  double s = 0.0, d = 0.0;
  if (live) {
#pragma unroll 8
    for (long long p = r; p < nparts; p += kFinRows) {
      const float* row = part + p * (c + 1) * K;
      s += static_cast<double>(__ldg(row + i));
      d += static_cast<double>(__ldg(row + c * K + k));
    }
  }
  ss[r][threadIdx.x] = s;
  sd[r][threadIdx.x] = d;
  __syncthreads();
  if (r == 0 && live) {
    s = 0.0;
    d = 0.0;
    for (int j = 0; j < kFinRows; ++j) {
      s += ss[j][threadIdx.x];
      d += sd[j][threadIdx.x];
    }
    d_ctr[i] = static_cast<float>(s - static_cast<double>(__ldg(ctr + i)) * d);
  }
}

size_t cos_smem(int c, int K);
size_t mpa_smem(int c, int K);
struct BwdPlan {
  int chunks, ntiles;
};
// How many CTAs share one image.  The grid is N * chunks CTAs of ceil(ntiles / chunks) tiles each on `slots` resident CTAs
// (two per SM at the PEMP size): pick the split with the smallest  waves * (tiles per CTA + prologue)  - e.g. 80 images of
// 82 tiles: 8 chunks = 2.16 waves of 11 tiles (a third, nearly empty wave), 11 chunks = 2.97 waves of 8.
BwdPlan bwd_plan(int N, int hw, size_t smem, int ntiles = 0) {
  BwdPlan p;
  p.ntiles = ntiles > 0 ? ntiles : (hw + 31) / 32;
  long long per_sm = static_cast<long long>(227 * 1024) / static_cast<long long>(smem + 1024);
  const long long by_regs = 65536 / (128LL * kBT);
  if (per_sm > by_regs) per_sm = by_regs;
  if (per_sm < 1) per_sm = 1;
  const long long slots = 148 * per_sm;
  double best = 1e30;
  p.chunks = 1;
  const int cmax = p.ntiles < 64 ? p.ntiles : 64;
  for (int ch = 1; ch <= cmax; ++ch) {
    const long long waves = (static_cast<long long>(N) * ch + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * ((p.ntiles + ch - 1) / ch + 0.75);
    if (cost < best - 1e-9) {
      best = cost;
      p.chunks = ch;
    }
  }
  return p;
}
size_t cos_smem(int c, int K) {
  const size_t KP = K, NA = KP + 1;
  return (static_cast<size_t>(c) * kLd + static_cast<size_t>(c) * KP + kBW * NA * 32 + NA * 32 + 32 * KP + 64 + 32 + 64) * 4;
}
size_t mpa_smem(int c, int K) {
  const size_t KP = K, NA = 2 * KP;
  return (static_cast<size_t>(c) * kLd + 2 * static_cast<size_t>(c) * KP + kBW * NA * 32 + NA * 32 + 2 * 32 * KP + 2 * K) * 4;
}

template <int K, bool DENSE>
int launch_cos_bwd(const float* qry, long long ep, int Q, const float* pn, const float* g_pred, int N, int c, int hw,
                   const BwdPlan& pl, float scalar, float* dq, long long d_ep, float* part, cudaStream_t st) {
  const size_t smem = cos_smem(c, K);
  cudaError_t e = cudaFuncSetAttribute(cosine_bwd_kernel<K, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  cosine_bwd_kernel<K, DENSE><<<dim3(pl.chunks, N), kBT, smem, st>>>(qry, ep, Q, pn, g_pred, c, hw, pl.ntiles, scalar, dq, d_ep, part);
  return PEMP_OK;
}
template <int K>
int launch_mpa_bwd(const float* fts, long long ep, int S, const float* ctr, const float* coef, const float* beta, const float* fg,
                   const float* bg, long long mask_stride, int N, int c, int hw, const BwdPlan& pl, float* dfts, long long d_ep,
                   float* part, cudaStream_t st) {
  const size_t smem = mpa_smem(c, K);
  cudaError_t e = cudaFuncSetAttribute(mpa_bwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  mpa_bwd_kernel<K><<<dim3(pl.chunks, N), kBT, smem, st>>>(fts, ep, S, ctr, coef, beta, fg, bg, mask_stride, c, hw, pl.ntiles, dfts, d_ep, part);
  return PEMP_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ K1 backward
// proto_g[b, c] = mean_s sum_x f m_g / (sum_x m_g + eps)  =>  d f[b, s, c, x] = (g_fg[b, c] m_fg[x] / den_fg + g_bg[b, c] m_bg[x]
// / den_bg) / S: a rank-2 outer product per image - write only, the features are not needed.
namespace {
__global__ void __launch_bounds__(256)
map_pool_bwd_kernel(const float* __restrict__ fg, const float* __restrict__ bg, long long mask_stride, const float* __restrict__ g_fg,
                    const float* __restrict__ g_bg, int S, int c, int hw, int rows_per_cta, float eps, float* __restrict__ d_fts,
                    long long d_ep_stride) {
  extern __shared__ float wm[];      // [2][hw] masks divided by S * den
  __shared__ float scratch[2][8];
  const int n = blockIdx.y, b = n / S, si = n - b * S;
  const float* mf = fg + static_cast<long long>(n) * mask_stride;
  const float* mb = bg ? bg + static_cast<long long>(n) * mask_stride : nullptr;
  float sf = 0.f, sb = 0.f;
  for (int i = threadIdx.x; i < hw; i += 256) {
    const float a = __ldg(mf + i), q = mb ? __ldg(mb + i) : 0.f;
    wm[i] = a;
    wm[hw + i] = q;
    sf += a;
    sb += q;
  }
  sf = warp_sum(sf);
  sb = warp_sum(sb);
  if ((threadIdx.x & 31) == 0) {
    scratch[0][threadIdx.x >> 5] = sf;
    scratch[1][threadIdx.x >> 5] = sb;
  }
  __syncthreads();
  sf = sb = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    sf += scratch[0][w];
    sb += scratch[1][w];
  }
  const float kf = 1.0f / (static_cast<float>(S) * (sf + eps)), kb = 1.0f / (static_cast<float>(S) * (sb + eps));
  const int c0 = blockIdx.x * rows_per_cta, c1 = min(c, c0 + rows_per_cta);
  float* dst = d_fts + static_cast<long long>(b) * d_ep_stride + static_cast<long long>(si) * c * hw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ch = c0 + warp; ch < c1; ch += 8) {
    const float gf = __ldg(g_fg + static_cast<long long>(b) * c + ch) * kf;
    const float gb = g_bg ? __ldg(g_bg + static_cast<long long>(b) * c + ch) * kb : 0.f;
    float* row = dst + static_cast<long long>(ch) * hw;
    for (int i = lane; i < hw; i += 32) row[i] = fmaf(gf, wm[i], gb * wm[hw + i]);
  }
}
}  // namespace

extern "C" int pemp_map_pool_lowres_bwd(const float* fg, const float* bg, long long mask_stride, const float* g_fg, const float* g_bg,
                                        int B, int S, int c, int hw, float eps, float* d_fts, long long d_fts_episode_stride,
                                        pemp_stream_t stream) {
  PEMP_REQUIRE(fg && g_fg && d_fts, PEMP_E_NULL);
  PEMP_REQUIRE((bg == nullptr) == (g_bg == nullptr), PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && c > 0 && hw > 0 && static_cast<long long>(B) * S <= 65535, PEMP_E_SHAPE);
  const size_t smem = static_cast<size_t>(hw) * 2 * sizeof(float);
  PEMP_REQUIRE(smem <= 200 * 1024, PEMP_E_SHAPE);
  const int N = B * S;
  int chunks = (4 * 148 + N - 1) / N;
  if (chunks > c / 8) chunks = c / 8;
  if (chunks < 1) chunks = 1;
  int rows = (c + chunks - 1) / chunks;
  rows = (rows + 7) / 8 * 8;             // a whole number of rows for each of the 8 warps
  chunks = (c + rows - 1) / rows;
  cudaError_t e = cudaFuncSetAttribute(map_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long d_ep = d_fts_episode_stride ? d_fts_episode_stride : static_cast<long long>(S) * c * hw;
  map_pool_bwd_kernel<<<dim3(chunks, N), 256, smem, as_stream(stream)>>>(fg, bg, mask_stride, g_fg, g_bg, S, c, hw, rows, eps, d_fts, d_ep);
  return launch_status();
}

namespace {
// workspace of the K3 backward: pn [Bp][c][K] | nrm [Bp][K] | partials [N][chunks][c][K] | tables of the tensor-path kernel.
// `chunks` is the larger of the two kernels' splits (the diagnostic switch picks the kernel at launch time).
struct CosBwdWs {
  size_t off_nrm, off_part, off_tab, total;
};
CosBwdWs cos_bwd_ws(int N, int Bp, int c, int hw, int P) {
  const size_t K = 2 * P;
  const bool mma = pemp_cos_bwd_mma_shape(c, P, hw);
  int chunks = bwd_plan(N, hw, cos_smem(c, 2 * P)).chunks;
  if (mma) {
    const int m = bwd_plan(N, hw, pemp_cos_bwd_mma_smem(c), pemp_cos_bwd_mma_tiles(hw)).chunks;
    if (m > chunks) chunks = m;
  }
  CosBwdWs w;
  w.off_nrm = align_up(static_cast<size_t>(Bp) * c * K * 4, 256);
  w.off_part = w.off_nrm + align_up(static_cast<size_t>(Bp) * K * 4, 256);
  w.off_tab = w.off_part + align_up(static_cast<size_t>(N) * chunks * c * K * 4, 256);
  w.total = w.off_tab + (mma ? align_up(pemp_cos_bwd_mma_table_bytes(Bp, c), 256) : 0);
  return w;
}
}  // namespace

extern "C" size_t pemp_cosine_match_bwd_workspace_bytes(int N, int Bp, int c, int hw, int P) {
  if (N <= 0 || Bp <= 0 || c <= 0 || hw <= 0 || P < 1 || P > 4) return 0;
  return cos_bwd_ws(N, Bp, c, hw, P).total;
}

namespace {
int cosine_bwd_common(bool dense, const float* qry, long long qry_episode_stride, const float* fg_proto, const float* bg_proto,
                      const float* g, int N, int Bp, int c, int hw, int P, float scalar, float* d_qry, long long d_qry_episode_stride,
                      float* d_fg, float* d_bg, void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(qry && fg_proto && bg_proto && g && d_qry && d_fg && d_bg, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && N <= 65535 && Bp > 0 && c > 0 && hw > 0 && N % Bp == 0, PEMP_E_SHAPE);
  PEMP_REQUIRE(P >= 1 && P <= 4 && c <= kMaxCPT * kBT && cos_smem(c, 2 * P) <= 227 * 1024, PEMP_E_SHAPE);
  PEMP_REQUIRE(workspace && workspace_bytes >= pemp_cosine_match_bwd_workspace_bytes(N, Bp, c, hw, P), PEMP_E_WORKSPACE);
  cudaStream_t st = as_stream(stream);
  const BwdPlan pl = bwd_plan(N, hw, cos_smem(c, 2 * P));
  const int Q = N / Bp;
  char* ws = static_cast<char*>(workspace);
  const CosBwdWs wl = cos_bwd_ws(N, Bp, c, hw, P);
  float* pn = reinterpret_cast<float*>(ws);
  float* nrm = reinterpret_cast<float*>(ws + wl.off_nrm);
  float* part = reinterpret_cast<float*>(ws + wl.off_part);
  const long long ep = qry_episode_stride ? qry_episode_stride : static_cast<long long>(Q) * c * hw;
  const long long d_ep = d_qry_episode_stride ? d_qry_episode_stride : static_cast<long long>(Q) * c * hw;
  proto_norm_kernel<<<Bp, kBT, 0, st>>>(fg_proto, bg_proto, c, P, pn, nrm);
  int rc = PEMP_E_ALIGN, chunks = pl.chunks;
  if (g_bwd_path == 0 && pemp_cos_bwd_mma_shape(c, P, hw)) {   // PEMP_E_ALIGN: no tensor map - the CUDA-core kernel takes it
    const int m = bwd_plan(N, hw, pemp_cos_bwd_mma_smem(c), pemp_cos_bwd_mma_tiles(hw)).chunks;
    rc = pemp_cos_bwd_mma_launch(dense, qry, ep, Bp, Q, pn, g, c, hw, m, scalar, reinterpret_cast<float*>(ws + wl.off_tab), d_qry,
                                 d_ep, part, st);
    if (rc == PEMP_OK) chunks = m;
  }
  if (rc == PEMP_E_ALIGN) {
#define PEMP_COS_BWD(KK)                                                                                                        \
  rc = dense ? launch_cos_bwd<KK, true>(qry, ep, Q, pn, g, N, c, hw, pl, scalar, d_qry, d_ep, part, st)                        \
             : launch_cos_bwd<KK, false>(qry, ep, Q, pn, g, N, c, hw, pl, scalar, d_qry, d_ep, part, st)
  switch (P) {
    case 1: PEMP_COS_BWD(2); break;
    case 2: PEMP_COS_BWD(4); break;
    case 3: PEMP_COS_BWD(6); break;
    default: PEMP_COS_BWD(8); break;
  }
#undef PEMP_COS_BWD
  }
  if (rc != PEMP_OK) return rc;
  cosine_bwd_finalize_kernel<<<Bp, kBT, 0, st>>>(part, pn, nrm, Q, chunks, c, P, d_fg, d_bg);
  return launch_status();
}
}  // namespace

extern "C" int pemp_cosine_match_bwd(const float* qry, long long qry_episode_stride, const float* fg_proto, const float* bg_proto,
                                     const float* g_pred, int N, int Bp, int c, int hw, int P, float scalar, float* d_qry,
                                     long long d_qry_episode_stride, float* d_fg, float* d_bg, void* workspace,
                                     size_t workspace_bytes, pemp_stream_t stream) {
  return cosine_bwd_common(false, qry, qry_episode_stride, fg_proto, bg_proto, g_pred, N, Bp, c, hw, P, scalar, d_qry,
                           d_qry_episode_stride, d_fg, d_bg, workspace, workspace_bytes, stream);
}

extern "C" int pemp_cosine_sim_bwd(const float* qry, long long qry_episode_stride, const float* fg_proto, const float* bg_proto,
                                   const float* g_sim, int N, int Bp, int c, int hw, int P, float scalar, float* d_qry,
                                   long long d_qry_episode_stride, float* d_fg, float* d_bg, void* workspace,
                                   size_t workspace_bytes, pemp_stream_t stream) {
  return cosine_bwd_common(true, qry, qry_episode_stride, fg_proto, bg_proto, g_sim, N, Bp, c, hw, P, scalar, d_qry,
                           d_qry_episode_stride, d_fg, d_bg, workspace, workspace_bytes, stream);
}

extern "C" int pemp_debug_bwd_path(int mode) {
  const int old = g_bwd_path;
  if (mode == 0 || mode == 1) g_bwd_path = mode;
  return old;
}

namespace {
// workspace of the K2 backward: coef [N][c][K] | beta [N][2K] | partials [N][chunks][(c + 1) K] | image tables of the tensor-path
// kernel.  `chunks` is the larger of the two kernels' splits (the diagnostic switch picks the kernel at launch time).
struct MpaBwdWs {
  size_t off_beta, off_part, off_tab, off_img, total;
};
MpaBwdWs mpa_bwd_ws(int N, int c, int hw, int p) {
  const size_t n = static_cast<size_t>(N), K = 2 * p;
  const bool mma = pemp_mpa_bwd_mma_shape(c, p, hw);
  int chunks = bwd_plan(N, hw, mpa_smem(c, 2 * p)).chunks;
  if (mma) {
    const int m = bwd_plan(N, hw, pemp_mpa_bwd_mma_smem(c), pemp_mpa_bwd_mma_tiles(hw)).chunks;
    if (m > chunks) chunks = m;
  }
  MpaBwdWs w;
  w.off_beta = align_up(n * c * K * 4, 256);
  w.off_part = w.off_beta + align_up(n * 2 * K * 4, 256);
  w.off_tab = w.off_part + align_up(n * chunks * (c + 1) * K * 4, 256);
  w.off_img = w.off_tab + (mma ? align_up(pemp_mpa_bwd_mma_table_bytes(N, c), 256) : 0);      // per-image sums of the partials
  w.total = w.off_img + (mma ? align_up(n * (c + 1) * K * 4, 256) : 0);
  return w;
}
}  // namespace

extern "C" size_t pemp_meta_proto_attn_bwd_workspace_bytes(int B, int S, int c, int hw, int p) {
  if (B <= 0 || S <= 0 || c <= 0 || hw <= 0 || p < 1 || p > 4) return 0;
  return mpa_bwd_ws(B * S, c, hw, p).total;
}

extern "C" int pemp_meta_proto_attn_bwd(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg,
                                        const float* bg, long long mask_stride, const float* shot_centre, const float* shot_den,
                                        const float* g_fg, const float* g_bg, int B, int S, int c, int hw, int p, float* d_fts,
                                        long long d_fts_episode_stride, float* d_ctr, void* workspace, size_t workspace_bytes,
                                        pemp_stream_t stream) {
  PEMP_REQUIRE(fts && ctr && fg && bg && shot_centre && shot_den && g_fg && g_bg && d_fts && d_ctr, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && c > 0 && hw > 0 && static_cast<long long>(B) * S <= 65535, PEMP_E_SHAPE);
  PEMP_REQUIRE(p >= 1 && p <= 4 && c <= kMaxCPT * kBT && mpa_smem(c, 2 * p) <= 227 * 1024, PEMP_E_SHAPE);
  PEMP_REQUIRE(workspace && workspace_bytes >= pemp_meta_proto_attn_bwd_workspace_bytes(B, S, c, hw, p), PEMP_E_WORKSPACE);
  cudaStream_t st = as_stream(stream);
  const int N = B * S, K = 2 * p;
  const bool mma = g_bwd_path == 0 && pemp_mpa_bwd_mma_shape(c, p, hw);
  const BwdPlan pl = bwd_plan(N, hw, mpa_smem(c, 2 * p));
  char* ws = static_cast<char*>(workspace);
  const MpaBwdWs wl = mpa_bwd_ws(N, c, hw, p);
  float* coef = reinterpret_cast<float*>(ws);
  float* beta = reinterpret_cast<float*>(ws + wl.off_beta);
  float* part = reinterpret_cast<float*>(ws + wl.off_part);
  const long long ep = fts_episode_stride ? fts_episode_stride : static_cast<long long>(S) * c * hw;
  const long long d_ep = d_fts_episode_stride ? d_fts_episode_stride : static_cast<long long>(S) * c * hw;
  float* tabg = mma ? reinterpret_cast<float*>(ws + wl.off_tab) : nullptr;
  mpa_bwd_prepare_kernel<<<N + 1, kBT, 0, st>>>(g_fg, g_bg, shot_centre, shot_den, ctr, N, S, c, p, coef, beta, tabg,
                                                pemp_mpa_bwd_mma_table_ld(), pemp_mpa_bwd_mma_rows_per_warp(c));
  int rc = PEMP_E_ALIGN;
  long long nparts = static_cast<long long>(N) * pl.chunks;
  if (mma) {                                     // PEMP_E_ALIGN: the operand has no tensor map - the CUDA-core kernel takes it
    const BwdPlan pm = bwd_plan(N, hw, pemp_mpa_bwd_mma_smem(c), pemp_mpa_bwd_mma_tiles(hw));
    float* img_part = reinterpret_cast<float*>(ws + wl.off_img);
    rc = pemp_mpa_bwd_mma_launch(fts, ep, B, S, ctr, coef, beta, fg, bg, mask_stride, c, hw, pm.chunks, tabg, d_fts, d_ep, part,
                                 img_part, st);
    if (rc == PEMP_OK) {                         // the partials of an image have already been added
      nparts = N;
      part = img_part;
    }
  }
  if (rc == PEMP_E_ALIGN)
  switch (p) {
    case 1: rc = launch_mpa_bwd<2>(fts, ep, S, ctr, coef, beta, fg, bg, mask_stride, N, c, hw, pl, d_fts, d_ep, part, st); break;
    case 2: rc = launch_mpa_bwd<4>(fts, ep, S, ctr, coef, beta, fg, bg, mask_stride, N, c, hw, pl, d_fts, d_ep, part, st); break;
    case 3: rc = launch_mpa_bwd<6>(fts, ep, S, ctr, coef, beta, fg, bg, mask_stride, N, c, hw, pl, d_fts, d_ep, part, st); break;
    default: rc = launch_mpa_bwd<8>(fts, ep, S, ctr, coef, beta, fg, bg, mask_stride, N, c, hw, pl, d_fts, d_ep, part, st); break;
  }
  if (rc != PEMP_OK) return rc;
  mpa_bwd_finalize_kernel<<<(c * K + 31) / 32, dim3(32, kFinRows), 0, st>>>(part, nparts, c, K, ctr, d_ctr);
  return launch_status();
}
