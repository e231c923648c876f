// K1 / K8 fast path: masked average pooling as a persistent, TMA-fed kernel (c = 512 or 256).
//
// Same arithmetic as `pool_partial_kernel` (pool.cu; reference networks/pemp_stage1.py:223-227, pemp_stage2.py:196-200,
// canet.py:176-178, panet.py:181-186, Weighted_GAP networks/pfenet.py:15-20, pooling half of baseline.py:105-110):
//   num[c, g] = sum_x f[c, x] * m_g[x],   den[g] = sum_x m_g[x]      (g = foreground, background)
// It is K2's phase B with the masks as weights (mpa_tma.cu explains the data path): the feature maps are described to
// TMA as [c/4 groups][4*hw floats], a box of 32 floats x c/4 groups at the 16-byte aligned inner coordinate
// (e*hw + x_nom) & ~3 holds the channels 4g + e (column i = pixel x_nom + i - o_e), tiles advance by 28 pixels, one CTA
// per SM owns a flat range of tiles, warp 16 feeds a ring of boxes.  Consumer warp w = 4e + cp owns columns
// [8cp, 8cp+8) of box e: lane <-> rows {l, l+32, ...} (the 128-byte swizzle makes the column read conflict-free), two
// LDS.128 per row slot and tile, one FFMA2 per (row, column pair, group) into {even, odd} column sums.  There is no
// exchange between warps inside an image, so the only synchronisation per tile is the ring.  Weights of a tile come
// straight from the mask rows (L2), fetched one tile ahead by the 16 lanes that own a (column, group) and handed to the
// warp through 64 floats of shared memory.  At an image boundary the four warps of a box fold their sums through the
// box they just consumed and write one partial per (image, CTA); the finalize kernel adds an image's partials in CTA
// order, divides (optionally by caller-supplied denominators: K6) and averages the shots - deterministic.
#include "tma_common.cuh"

int pemp_pool_tma_launch(const float* fts, long long ep_stride, const float* fg, const float* bg, long long mask_stride, int B,
                         int S, int c, int hw, float eps, const float* den_override, float* fg_proto, float* bg_proto, char* ws,
                         size_t ws_bytes, cudaStream_t st);
size_t pemp_pool_tma_workspace_bytes(int B, int S, int c, int hw);

namespace {

constexpr int kTW = 32;
constexpr int kStep = 28;
constexpr int kCons = 16;
constexpr int kThreadsP = (kCons + 1) * 32;
constexpr int kMaxGrid = 148;

template <int C>
struct PoolCfg {
  static constexpr int kRows = C / 4;                       // rows of a box
  static constexpr int kBoxFloats = kRows * kTW;
  static constexpr uint32_t kBoxBytes = kBoxFloats * 4;
  static constexpr int kRL = kRows / 32;                    // row slots per lane
  static constexpr int kNB = C == 512 ? 12 : 16;            // ring slots (the consumers hold 4)
  // whole tiles only: a slot always carries the same channel class, so the warp that waits for use u+1 of a slot is the
  // one that consumed use u (otherwise a parity wait can pass on the phase before; see cosine_tma.cu)
  static_assert(kNB % 4 == 0, "ring slots must be a multiple of the 4 boxes of a tile");
};

template <int C>
struct PoolSmem {
  alignas(1024) float ring[PoolCfg<C>::kNB][PoolCfg<C>::kBoxFloats];
  alignas(16) float wts[kCons][16];                         // [warp][column pair][{fg_even, fg_odd, bg_even, bg_odd}]
  alignas(8) uint64_t full[PoolCfg<C>::kNB];
  alignas(8) uint64_t empty[PoolCfg<C>::kNB];
};

using namespace pemp_tma;


template <int C, bool kTwo>
__global__ void __launch_bounds__(kThreadsP, 1)
pool_tma_kernel(const __grid_constant__ CUtensorMap map, int S, int hw, int nt_img, long long T, int imgs,
                int* __restrict__ nparts, const float* __restrict__ fg, const float* __restrict__ bg, long long mask_stride,
                int maxp, float* __restrict__ part_num, float* __restrict__ part_den) {
  using Cfg = PoolCfg<C>;
  constexpr int kNB = Cfg::kNB, kRL = Cfg::kRL, kRows = Cfg::kRows;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  PoolSmem<C>& sm = *reinterpret_cast<PoolSmem<C>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const long long t0 = T * cta / G, t1 = T * (cta + 1) / G;
  const int ntl = static_cast<int>(t1 - t0);

  // partials an image ends up with = CTAs its tile range touches (read by the finalize kernel)
  for (int i = cta * kThreadsP + tid; i < imgs; i += G * kThreadsP) {
    const long long first = static_cast<long long>(i) * nt_img;
    nparts[i] = owner_of(first + nt_img - 1, T, G) - owner_of(first, T, G) + 1;
  }
  if (tid == 0) {
    for (int s = 0; s < kNB; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (ntl <= 0) return;

  if (warp == kCons) {
    // ============================ producer ============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
      int slot = 0;
      uint32_t par = 1;
      int img = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(img) * nt_img);
      int ep = img / S, s = img - ep * S;
      for (int k = 0; k < ntl; ++k) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = (e * hw + tl * kStep) & ~3;
          mbar_wait(&sm.empty[slot], par);
          mbar_expect_tx(&sm.full[slot], Cfg::kBoxBytes);
          tma_load_3d(&map, &sm.full[slot], sm.ring[slot], c0, s * kRows, ep);
          if (++slot == kNB) {
            slot = 0;
            par ^= 1;
          }
        }
        if (++tl == nt_img) {
          tl = 0;
          ++img;
          if (++s == S) {
            s = 0;
            ++ep;
          }
        }
      }
    }
    return;
  }

  // ============================ consumers ============================
  const int e = warp >> 2, cp = warp & 3;
  // weight roles: lanes 0..15 own (column 8cp + (lane & 7), group lane >> 3)
  const int col_w = cp * 8 + (lane & 7), g_w = (lane >> 3) & 1;
  const bool owner = lane < (kTwo ? 16 : 8);
  const float* mask_ptr = (g_w ? bg : fg) + col_w;
  int off_b = lane * kTW + (((2 * cp) ^ (lane & 7)) << 2);        // chunk 2cp; chunk 2cp+1: ^ 4; row step 32*i
  asm volatile("" : "+r"(off_b));
  float* const wts = sm.wts[warp];

  float2 acc[kRL][2];                                             // [row slot][group] = {even, odd column} sums
#pragma unroll
  for (int i = 0; i < kRL; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
  float den = 0.f;                                                // owner lanes: sum of their weights

  auto mask_of = [&](int img, int x_nom, int o) {
    const int p = col_w - o;
    const bool ok = owner && p >= 0 && p < kStep && x_nom + p < hw;
    return ok ? __ldg(mask_ptr + (img * mask_stride + (x_nom - o))) : 0.f;
  };

  int slot = e;
  uint32_t par = 0;
  int img = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(img) * nt_img);
  float m = mask_of(img, tl * kStep, (e * hw + tl * kStep) & 3);
  for (int k = 0; k < ntl; ++k) {
    const bool have_next = k + 1 < ntl;
    const bool last_of_img = (tl == nt_img - 1) || !have_next;
    int tl_n = tl + 1, img_n = img;
    if (tl_n == nt_img) {
      tl_n = 0;
      ++img_n;
    }
    // weights of this tile -> shared (column pair layout), next tile's mask value -> register
    den += m;
    if (owner) wts[((lane & 7) >> 1) * 4 + g_w * 2 + (lane & 1)] = m;
    const unsigned live = __ballot_sync(kFull, m != 0.f);
    float m_n = 0.f;
    if (have_next) m_n = mask_of(img_n, tl_n * kStep, (e * hw + tl_n * kStep) & 3);
    __syncwarp();
    const float* box = sm.ring[slot];
    mbar_wait(&sm.full[slot], par);
    if (live) {
#pragma unroll
      for (int ck = 0; ck < 2; ++ck) {
        float4 f[kRL];
#pragma unroll
        for (int i = 0; i < kRL; ++i) f[i] = *reinterpret_cast<const float4*>(box + (off_b ^ (ck << 2)) + 32 * i * kTW);
        const float4 wa = *reinterpret_cast<const float4*>(wts + (2 * ck) * 4);        // columns 4ck, 4ck+1
        const float4 wb = *reinterpret_cast<const float4*>(wts + (2 * ck + 1) * 4);    // columns 4ck+2, 4ck+3
#pragma unroll
        for (int i = 0; i < kRL; ++i) {
          const float2 f01 = make_float2(f[i].x, f[i].y), f23 = make_float2(f[i].z, f[i].w);
          acc[i][0] = ffma2(f01, make_float2(wa.x, wa.y), acc[i][0]);
          acc[i][0] = ffma2(f23, make_float2(wb.x, wb.y), acc[i][0]);
          if (kTwo) {
            acc[i][1] = ffma2(f01, make_float2(wa.z, wa.w), acc[i][1]);
            acc[i][1] = ffma2(f23, make_float2(wb.z, wb.w), acc[i][1]);
          }
        }
      }
    }

    if (!last_of_img) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
    } else {
      // ---------------- image boundary: fold the four column-owning warps of this box, write the partial ----------
      float* scratch = sm.ring[slot];                 // [3][2*kRL][32] numerators, then [4][2] denominators
      // denominators: lanes 0..7 hold foreground weights, 8..15 background
      den += __shfl_xor_sync(kFull, den, 1);
      den += __shfl_xor_sync(kFull, den, 2);
      den += __shfl_xor_sync(kFull, den, 4);
      named_bar(2 + e, 128);                          // all four warps are done reading the box
      if (cp > 0) {
#pragma unroll
        for (int i = 0; i < kRL; ++i)
#pragma unroll
          for (int g = 0; g < 2; ++g) scratch[((cp - 1) * 2 * kRL + i * 2 + g) * 32 + lane] = acc[i][g].x + acc[i][g].y;
      }
      if (lane == 0 || lane == 8) scratch[3 * 2 * kRL * 32 + cp * 2 + (lane >> 3)] = den;
      named_bar(2 + e, 128);
      if (cp == 0) {
        const int slot_idx = cta - owner_of(static_cast<long long>(img) * nt_img, T, G);
        float* out = part_num + (static_cast<long long>(img) * maxp + slot_idx) * (C * 2);
#pragma unroll
        for (int i = 0; i < kRL; ++i) {
          float v[2];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float a = acc[i][g].x + acc[i][g].y;
#pragma unroll
            for (int r = 0; r < 3; ++r) a += scratch[(r * 2 * kRL + i * 2 + g) * 32 + lane];
            v[g] = a;
          }
          *reinterpret_cast<float2*>(out + (e * kRows + lane + 32 * i) * 2) = make_float2(v[0], v[1]);
        }
        if (e == 0 && lane < 2) {
          float sden = 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r) sden += scratch[3 * 2 * kRL * 32 + r * 2 + lane];
          part_den[(static_cast<long long>(img) * maxp + slot_idx) * 2 + lane] = sden;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // scratch writes before the slot's next TMA fill
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
#pragma unroll
      for (int i = 0; i < kRL; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
      den = 0.f;
    }
    m = m_n;
    slot += 4;
    if (slot >= kNB) {
      slot -= kNB;
      par ^= 1;
    }
    tl = tl_n;
    img = img_n;
  }
}

// one thread per (b, channel): partials of every shot in CTA order, ratio per shot, mean over shots
template <int C>
__global__ void pool_tma_finalize_kernel(const float* __restrict__ part_num, const float* __restrict__ part_den,
                                         const int* __restrict__ nparts, const float* __restrict__ den_override, int B, int S,
                                         int maxp, float eps, float* __restrict__ fg_proto, float* __restrict__ bg_proto) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, ch = i - b * C;
  const int R = (ch & 3) * (C / 4) + (ch >> 2);
  float accf = 0.f, accb = 0.f;
  for (int s = 0; s < S; ++s) {
    const long long img = static_cast<long long>(b) * S + s;
    const int n = __ldg(nparts + img);
    float nf = 0.f, nb = 0.f, df = 0.f, db = 0.f;
    for (int sp = 0; sp < n; ++sp) {
      const float2 p = *reinterpret_cast<const float2*>(part_num + ((img * maxp + sp) * C + R) * 2);
      nf += p.x;
      nb += p.y;
      df += part_den[(img * maxp + sp) * 2 + 0];
      db += part_den[(img * maxp + sp) * 2 + 1];
    }
    if (den_override) {
      df = den_override[img * 2 + 0];
      db = den_override[img * 2 + 1];
    }
    accf += nf / (df + eps);
    accb += nb / (db + eps);
  }
  fg_proto[i] = accf / static_cast<float>(S);
  if (bg_proto) bg_proto[i] = accb / static_cast<float>(S);
}


struct PoolPlan {
  int G, nt_img, maxp;
  long long T;
  size_t off_nparts, off_num, off_den, total;
};
PoolPlan make_pool_plan(int B, int S, int c, int hw) {
  PoolPlan p;
  p.nt_img = (hw + kStep - 1) / kStep;
  p.T = static_cast<long long>(B) * S * p.nt_img;
  long long g = p.T / 4;
  p.G = static_cast<int>(g < 1 ? 1 : (g > kMaxGrid ? kMaxGrid : g));
  const long long per_cta = p.T / p.G;
  p.maxp = static_cast<int>((p.nt_img + per_cta - 1) / per_cta) + 1;
  const size_t imgs = static_cast<size_t>(B) * S;
  p.off_nparts = 0;
  p.off_num = align_up(imgs * sizeof(int), 256);
  p.off_den = p.off_num + align_up(imgs * p.maxp * c * 2 * sizeof(float), 256);
  p.total = p.off_den + align_up(imgs * p.maxp * 2 * sizeof(float), 256);
  return p;
}

template <int C>
int launch_pool(const CUtensorMap& map, const PoolPlan& pl, int B, int S, int hw, const float* fg, const float* bg,
                long long mask_stride, float eps, const float* den_override, float* fg_proto, float* bg_proto, char* ws,
                cudaStream_t st) {
  int* nparts = reinterpret_cast<int*>(ws + pl.off_nparts);
  float* num = reinterpret_cast<float*>(ws + pl.off_num);
  float* den = reinterpret_cast<float*>(ws + pl.off_den);
  const size_t smem = sizeof(PoolSmem<C>);
  cudaError_t e;
  if (bg) {
    e = cudaFuncSetAttribute(pool_tma_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    pool_tma_kernel<C, true><<<pl.G, kThreadsP, smem, st>>>(map, S, hw, pl.nt_img, pl.T, B * S, nparts, fg, bg, mask_stride, pl.maxp,
                                                           num, den);
  } else {
    e = cudaFuncSetAttribute(pool_tma_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    pool_tma_kernel<C, false><<<pl.G, kThreadsP, smem, st>>>(map, S, hw, pl.nt_img, pl.T, B * S, nparts, fg, fg, mask_stride, pl.maxp,
                                                            num, den);
  }
  const int total = B * C;
  pool_tma_finalize_kernel<C><<<(total + 255) / 256, 256, 0, st>>>(num, den, nparts, den_override, B, S, pl.maxp, eps, fg_proto,
                                                                  bg ? bg_proto : nullptr);
  return launch_status();
}

}  // namespace

size_t pemp_pool_tma_workspace_bytes(int B, int S, int c, int hw) {
  if ((c != 512 && c != 256) || hw < kTW) return 0;
  return make_pool_plan(B, S, c, hw).total;
}

// Returns PEMP_E_ALIGN (nothing launched) when the shape or the operand is not covered; the caller then uses the
// generic kernel.
int pemp_pool_tma_launch(const float* fts, long long ep_stride, const float* fg, const float* bg, long long mask_stride, int B,
                         int S, int c, int hw, float eps, const float* den_override, float* fg_proto, float* bg_proto, char* ws,
                         size_t ws_bytes, cudaStream_t st) {
  const long long eps_stride = ep_stride ? ep_stride : static_cast<long long>(S) * c * hw;
  if ((c != 512 && c != 256) || hw < kTW) return PEMP_E_ALIGN;
  const PoolPlan pl = make_pool_plan(B, S, c, hw);
  if (ws_bytes < pl.total) return PEMP_E_ALIGN;
  CUtensorMap map;
  if (!make_rows4_map(&map, fts, B, S, c, hw, eps_stride)) return PEMP_E_ALIGN;
  return c == 512 ? launch_pool<512>(map, pl, B, S, hw, fg, bg, mask_stride, eps, den_override, fg_proto, bg_proto, ws, st)
                  : launch_pool<256>(map, pl, B, S, hw, fg, bg, mask_stride, eps, den_override, fg_proto, bg_proto, ws, st);
}
