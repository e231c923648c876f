// K3 fast path: cosine matching as a persistent, TMA-fed, warp-specialised kernel (c = 512; P = 3 or 1).
//
// Same arithmetic as `cosine_match_kernel` (cosine.cu; reference networks/pemp_stage1.py:214-222,233-261,
// pemp_stage2.py:187-194,205-233, baseline.py:121-149, panet.py:122-156):
//   s[k, x] = (sum_c q[c, x] * p^[c, k]) * (1 / max(|q[:, x]|, 1e-8)) * scalar,   p^ = p / max(|p|, 1e-8)
//   pred[g, x] = max_j s[g*P + j, x]   (first maximum wins),  response = argmax index (+3 for foreground)
//
// Data path as in mpa_tma.cu: the query maps [.., c, hw] are described to TMA as [c/4 groups][4*hw floats]; a box of
// 32 floats x 128 groups at the 16-byte aligned inner coordinate (e*hw + x_nom) & ~3 holds the channels 4g + e, its
// column i is pixel x_nom + i - o_e; tiles advance by 28 pixels.  One CTA per SM owns a flat range of tiles; warp 16
// feeds an 8-slot ring of 16-KB boxes (full / empty mbarriers).  Consumer warp w = 4e + cp reads rows [32cp, 32cp+32)
// of box e: lane <-> (row mod 4, 16-byte chunk), one row-contiguous LDS.128 per 4 pixels of a channel, the channel's
// normalised prototypes (each value twice, so they are FFMA2 operands as loaded) from shared memory, 2 + 2K packed
// FFMA2 per load for |q|^2 and the K dots of 4 pixels.  The slot is released as soon as the warp has read it.  A
// halving butterfly over the 4 row groups leaves each lane with the 1 + K sums of one pixel; they go to
// part[buffer][warp][value][pixel] and the warp arrives on the buffer's mbarrier.  Two finishing warps (17, 18; 14
// pixels each, lane <-> (pixel, half of the partials)) wait for that barrier, add the 16 partials in a fixed order, hand
// the buffer back (second mbarrier) and finish the pixel: norm, scale, max / argmax, stores - so the 16 consumer warps
// never leave their load / FFMA2 loop (in the first version they finished the previous tile themselves: 0.70 of the
// HBM peak with four exchange buffers, 0.82 now).
//
// The prototype table depends on the episode: the consumers (re)build it in shared memory - normalised, in tile-row
// order - whenever the episode of the current image changes (at most twice per CTA at the bench shape).
#include "tma_common.cuh"

int pemp_cosine_tma_launch(const float* qry, long long ep_stride, const float* fg, const float* bg, int N, int Bp, int hw,
                           int P, float scalar, float* sim, float* pred, int64_t* response, cudaStream_t st);

namespace {

constexpr int kC = 512;
constexpr int kTW = 32;                          // floats per box row
constexpr int kStep = 28;                        // pixels per tile
constexpr int kBoxRows = kC / 4;
constexpr int kBoxFloats = kBoxRows * kTW;
constexpr uint32_t kBoxBytes = kBoxFloats * 4;
#ifndef PEMP_COS_NB
#define PEMP_COS_NB 8
#endif
constexpr int kNB = PEMP_COS_NB;                 // ring slots (the consumers hold 4, 4 in flight)
// The ring must be a whole number of tiles (4 boxes): then slot s always carries the same channel class and the warp that
// waits for use u+1 of a slot is the one that consumed use u.  With 10 or 11 slots a slot alternates between classes;
// a warp of the other class can reach its parity wait for use u+1 before use u has even landed (boxes complete out of
// order, and this kernel's consumers are usually waiting for data) - the parity test then passes on the phase BEFORE,
// the warp reads a stale box and releases a slot it does not own: measured as `unspecified launch failure` within a
// few hundred launches (8 slots: none; same throughput).
static_assert(kNB % 4 == 0, "ring slots must be a multiple of the 4 boxes of a tile");
constexpr int kCons = 16;
constexpr int kThreadsC = (kCons + 3) * 32;      // + producer warp + two finishing warps
#ifndef PEMP_COS_PB
#define PEMP_COS_PB 2
#endif
constexpr int kPB = PEMP_COS_PB;                 // exchange buffers (handed back by the finishing warps: free_bar)
constexpr int kPLd = 29;                         // pixel pitch of a value row in `part`
constexpr int kMaxGrid = 148;
constexpr float kCosEps = 1e-8f;

#ifndef PEMP_COS_DUP
#define PEMP_COS_DUP 0      // 1: table rows as {p0,p0,p1,p1,...} (FFMA2 operands as loaded, 3 LDS.128 per row at K = 6); 0: {p0,p1,...}
#endif                      // (LDS.128 + LDS.64 and one register copy per value: fewer shared-memory wavefronts - ncu: the pipe is the limit)
template <int K>
struct CosSmem {
  static constexpr int TL = PEMP_COS_DUP ? (2 * K + 3) / 4 * 4 : (K + 1) / 2 * 2;  // floats per table row
  alignas(1024) float ring[kNB][kBoxFloats];
  alignas(16) float table[kC * TL];
  alignas(16) float part[kPB][kCons][(1 + K) * kPLd]; // [buffer][warp][value][pixel], odd pitch: see finalize
  alignas(16) float red[kCons][8];
  alignas(8) uint64_t full[kNB];
  alignas(8) uint64_t empty[kNB];
  alignas(8) uint64_t part_bar[kPB];               // all 16 consumer warps have written part[b]
  alignas(8) uint64_t free_bar[kPB];               // both finishing warps have read part[b]
};

using namespace pemp_tma;

template <int K>
__global__ void __launch_bounds__(kThreadsC, 1)
cosine_tma_kernel(const __grid_constant__ CUtensorMap map, int Qper, int hw, int nt_img, long long T,
                  const float* __restrict__ fg_proto, const float* __restrict__ bg_proto, float scalar,
                  float* __restrict__ sim, float* __restrict__ pred, int64_t* __restrict__ response) {
  constexpr int P = K / 2, NV = 1 + K, TL = CosSmem<K>::TL;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  CosSmem<K>& sm = *reinterpret_cast<CosSmem<K>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  asm volatile("" : "+r"(tid), "+r"(lane), "+r"(warp));      // keep them in registers (no S2R re-reads in the loop)
  const int G = gridDim.x, cta = blockIdx.x;
  const long long t0 = T * cta / G, t1 = T * (cta + 1) / G;
  const int ntl = static_cast<int>(t1 - t0);

  if (tid == 0) {
    for (int s = 0; s < kNB; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 4);
    }
    for (int b = 0; b < kPB; ++b) {
      mbar_init(&sm.part_bar[b], kCons);
      mbar_init(&sm.free_bar[b], 2);
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
      int n = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(n) * nt_img);
      int ep = n / Qper, q = n - ep * Qper;
      for (int k = 0; k < ntl; ++k) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = (e * hw + tl * kStep) & ~3;
          mbar_wait(&sm.empty[slot], par);
          mbar_expect_tx(&sm.full[slot], kBoxBytes);
          tma_load_3d(&map, &sm.full[slot], sm.ring[slot], c0, q * kBoxRows, ep);
          if (++slot == kNB) {
            slot = 0;
            par ^= 1;
          }
        }
        if (++tl == nt_img) {
          tl = 0;
          if (++q == Qper) {
            q = 0;
            ++ep;
          }
        }
      }
    }
    return;
  }

  if (warp > kCons) {
    // ============================ finishing warps ============================
    // Warp 17 + hwp finishes pixels [14*hwp, 14*hwp + 14) of every tile: lane = pixel + 14*half adds the partial sums of
    // the consumer warps [8*half, 8*half + 8) in index order, the two halves are combined (lower + upper, fixed), then
    // lanes 0..13 do norm, scale, max / argmax over the P prototypes of each group (first maximum wins) and the stores.
    const int hwp = warp - kCons - 1;
    const int half = lane >= 14 ? 1 : 0, pix = 14 * hwp + (lane - 14 * half);
    const bool act = lane < 28;
    int n = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(n) * nt_img);
    for (int k = 0; k < ntl; ++k) {
      const int pb = k & (kPB - 1), x = tl * kStep + pix;
      mbar_wait(&sm.part_bar[pb], (k / kPB) & 1);
      float s[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) s[v] = 0.f;
      if (act) {
        const float* src = &sm.part[pb][half * 8][pix];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int v = 0; v < NV; ++v) s[v] += src[(i * NV + v) * kPLd];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.free_bar[pb]);          // part[pb] may be rewritten
#pragma unroll
      for (int v = 0; v < NV; ++v) s[v] += __shfl_down_sync(kFull, s[v], 14);   // lanes 0..13: lower + upper half
      if (lane < 14 && x < hw) {
        const float qinv = 1.0f / fmaxf(sqrtf(s[0]), kCosEps);
        float best[2];
        int arg[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          best[g] = -INFINITY;
          arg[g] = 0;
#pragma unroll
          for (int j = 0; j < P; ++j) {
            const float sv = s[1 + g * P + j] * qinv * scalar;
            if (sim) sim[((static_cast<long long>(n) * 2 + g) * P + j) * hw + x] = sv;
            if (sv > best[g]) {   // first maximum wins, as torch.max
              best[g] = sv;
              arg[g] = j;
            }
          }
        }
        if (pred) {
          pred[(static_cast<long long>(n) * 2 + 0) * hw + x] = best[0];
          pred[(static_cast<long long>(n) * 2 + 1) * hw + x] = best[1];
        }
        if (response) response[static_cast<long long>(n) * hw + x] = best[1] > best[0] ? arg[1] + 3 : arg[0];
      }
      if (++tl == nt_img) {
        tl = 0;
        ++n;
      }
    }
    return;
  }

  // ============================ consumers ============================
  const int e = warp >> 2, cp = warp & 3;
  const int rg = lane >> 3, jc = lane & 7;
  int off_a = (cp * 32 + rg) * kTW + ((jc ^ rg) << 2);      // even i; odd i: ^ 16; row step i*128
  int off_t = (e * kBoxRows + cp * 32 + rg) * TL;
  int col_a = 4 * jc + rg;                                  // box column this lane holds after the butterfly
  asm volatile("" : "+r"(off_a), "+r"(off_t), "+r"(col_a));
  int slot = e;
  uint32_t par = 0;
  int n = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(n) * nt_img);
  int b = n / Qper, q_in_b = n - b * Qper;                    // episode of image n, tracked without divisions
  int cur_b = -1;
  for (int k = 0; k < ntl; ++k) {
    if (b != cur_b) {
      // ---------------- (re)build the normalised prototype table of episode b ----------------
      cur_b = b;
      named_bar(1, kCons * 32);                              // nobody reads the old table any more
      const int R = tid, ch = 4 * (R & (kBoxRows - 1)) + (R >> 7);
      float raw[K];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        raw[j] = __ldg(bg_proto + (static_cast<long long>(b) * kC + ch) * P + j);
        raw[P + j] = __ldg(fg_proto + (static_cast<long long>(b) * kC + ch) * P + j);
      }
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        const float ss = warp_sum(raw[kk] * raw[kk]);
        if (lane == 0) sm.red[warp][kk] = ss;
      }
      named_bar(1, kCons * 32);
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        float ss = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kCons; ++w2) ss += sm.red[w2][kk];
        const float v = raw[kk] * (1.0f / fmaxf(sqrtf(ss), kCosEps));
        if (PEMP_COS_DUP) {
          sm.table[R * TL + 2 * kk] = v;
          sm.table[R * TL + 2 * kk + 1] = v;
        } else {
          sm.table[R * TL + kk] = v;
        }
      }
      named_bar(1, kCons * 32);
    }
    const int x_nom = tl * kStep;
    const int o = (e * hw + x_nom) & 3;
    const float* box = sm.ring[slot];
    mbar_wait(&sm.full[slot], par);

    float2 acc[2][NV];                                       // [column pair][|q|^2, K dots]
#pragma unroll
    for (int p2 = 0; p2 < 2; ++p2)
#pragma unroll
      for (int kk = 0; kk < NV; ++kk) acc[p2][kk] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 f = *reinterpret_cast<const float4*>(box + (off_a ^ ((i & 1) << 4)) + i * 4 * kTW);
      const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
      const float* trow = sm.table + off_t + i * 4 * TL;
      acc[0][0] = ffma2(f01, f01, acc[0][0]);
      acc[1][0] = ffma2(f23, f23, acc[1][0]);
#pragma unroll
      for (int kk = 0; kk < K; kk += 2) {
        float2 ta, tb;
        if (PEMP_COS_DUP) {
          const float4 t4 = *reinterpret_cast<const float4*>(trow + 2 * kk);
          ta = make_float2(t4.x, t4.y), tb = make_float2(t4.z, t4.w);
        } else {
          const float2 t2 = *reinterpret_cast<const float2*>(trow + kk);
          ta = make_float2(t2.x, t2.x), tb = make_float2(t2.y, t2.y);
        }
        acc[0][1 + kk] = ffma2(f01, ta, acc[0][1 + kk]);
        acc[1][1 + kk] = ffma2(f23, ta, acc[1][1 + kk]);
        acc[0][2 + kk] = ffma2(f01, tb, acc[0][2 + kk]);
        acc[1][2 + kk] = ffma2(f23, tb, acc[1][2 + kk]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[slot]);             // the box is not read again

    // the finishing warps must be done with the tile that used this exchange buffer last
    if (k >= kPB) mbar_wait(&sm.free_bar[k & (kPB - 1)], ((k / kPB) + 1) & 1);
    // halving butterfly over the row groups (lane bits 4 and 3): lane (rg, jc) ends with column 4*jc + rg
    {
      const bool hi = (lane & 16) != 0, lo = (lane & 8) != 0;
      const int p = col_a - o;
      float* dst = &sm.part[k & (kPB - 1)][warp][p];
#pragma unroll
      for (int kk = 0; kk < NV; ++kk) {
        const float2 keep = hi ? acc[1][kk] : acc[0][kk], send = hi ? acc[0][kk] : acc[1][kk];
        const float rx = keep.x + __shfl_xor_sync(kFull, send.x, 16);
        const float ry = keep.y + __shfl_xor_sync(kFull, send.y, 16);
        const float keep2 = lo ? ry : rx, send2 = lo ? rx : ry;
        const float q = keep2 + __shfl_xor_sync(kFull, send2, 8);
        if (p >= 0 && p < kStep) dst[kk * kPLd] = q;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.part_bar[k & (kPB - 1)]);

    slot += 4;
    if (slot >= kNB) {
      slot -= kNB;
      par ^= 1;
    }
    if (++tl == nt_img) {
      tl = 0;
      ++n;
      if (++q_in_b == Qper) {
        q_in_b = 0;
        ++b;
      }
    }
  }
}


template <int K>
int launch_tma(const CUtensorMap& map, int Qper, int hw, int nt_img, long long T, int G, const float* fg, const float* bg,
               float scalar, float* sim, float* pred, int64_t* response, cudaStream_t st) {
  const size_t smem = sizeof(CosSmem<K>);
  cudaError_t e = cudaFuncSetAttribute(cosine_tma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  cosine_tma_kernel<K><<<G, kThreadsC, smem, st>>>(map, Qper, hw, nt_img, T, fg, bg, scalar, sim, pred, response);
  return launch_status();
}

}  // namespace

// Returns PEMP_E_ALIGN (nothing launched) when the shape or the operand is not covered; the caller then uses the
// generic kernel.
int pemp_cosine_tma_launch(const float* qry, long long ep_stride, const float* fg, const float* bg, int N, int Bp, int hw,
                           int P, float scalar, float* sim, float* pred, int64_t* response, cudaStream_t st) {
  const int Qper = N / Bp;
  const long long eps_stride = ep_stride ? ep_stride : static_cast<long long>(Qper) * kC * hw;
  if ((P != 3 && P != 1) || hw < kTW) return PEMP_E_ALIGN;
  const int nt_img = (hw + kStep - 1) / kStep;
  const long long T = static_cast<long long>(N) * nt_img;
  long long g = T / 4;
  const int G = static_cast<int>(g < 1 ? 1 : (g > kMaxGrid ? kMaxGrid : g));
  CUtensorMap map;
  if (!make_rows4_map(&map, qry, Bp, Qper, kC, hw, eps_stride)) return PEMP_E_ALIGN;
  return P == 3 ? launch_tma<6>(map, Qper, hw, nt_img, T, G, fg, bg, scalar, sim, pred, response, st)
                : launch_tma<2>(map, Qper, hw, nt_img, T, G, fg, bg, scalar, sim, pred, response, st);
}
