// K2  meta-prototype attention: adaptive prototypes from support features, masks and the learnt centres.
//
// replaces the `self.ctr is not None` branch of mpm():
//   networks/pemp_stage1.py:202-213, networks/pemp_stage2.py:174-186
//     D[k,x]  = -sum_c (f[c,x] - ctr[c,k])^2                                  k = g*P + j, g=0 foreground
//     A[k,x]  = softmax_j(D[g*P+j, x]) * mask_g[x]
//     out[c,k]= sum_x f[c,x] A[k,x] / (sum_x A[k,x] + eps);   mean over the S shots
//
// The reference materialises [BS, c, 2P, hw] four times (32 MB / shot each).  Here every support feature is
// read from HBM exactly once:  algorithmic bytes per shot = (c*hw + 2*hw)*4.
//
// Arithmetic.  softmax over a group only needs differences  D[g*P+j] - D[g*P]  =  2 f.(ctr_j - ctr_0) -
// (|ctr_j|^2 - |ctr_0|^2), so phase A is 2(P-1) dot products per pixel against difference vectors that a
// tiny prologue kernel prepares once per launch (in double).  This is both cheaper than the 2P squared
// distances (|D| ~ 300 loses 1e-4 absolute in fp32; the differences are O(10)) and closer to the exact
// result.  Phase B is the weighted sum; pixels whose group mask is zero are skipped (warp-uniform pixel
// lists), which halves the work for complementary fg/bg masks.
//
// Mapping.  A *cluster* of CS CTAs (CS = 2 when the channel count allows) owns a tile of TW = 32*CS pixels
// of one image; CTA r of the cluster owns the channels [r*c/CS, (r+1)*c/CS).  Per tile:
//   phase A  lane <-> CS pixels (l, l+32), warp <-> a contiguous range of channel quads: four rows per step are
//            loaded with coalesced row segments (TW*4 contiguous bytes per row), multiplied against the
//            difference table (one 128-bit shared load per row, amortised over the lane's CS pixels) and
//            stored TRANSPOSED into shared memory as Ft[pixel][channel] (row stride = 4*odd floats =>
//            conflict-free 128-bit stores);
//   exchange + softmax: 2*TW threads, one per (class group, pixel), add the warps' partial dots in a fixed
//            order, push the sums into the peer CTA's shared memory with `st.async` (which completes a
//            transaction mbarrier there - no fence, so global loads in flight are not drained) and, once the
//            peer's sums have landed, turn the dots into the 2P weights of each pixel (both CTAs, redundantly);
//   phase B  thread <-> kQPT channel quads x one pixel group: conflict-free 128-bit shared loads of Ft feed
//            4*P FMAs per quad into register accumulators that live for the whole CTA.
// Partial numerators / denominators go to the workspace per (image, split); `mpa_finalize_kernel` adds the
// splits in index order, divides, and averages the shots - deterministic, no atomics.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

// TMA-fed persistent fast path for c = 512, P = 3 (mpa_tma.cu); PEMP_E_ALIGN = "operand not describable, use the
// generic kernel", nothing launched.
int pemp_mpa_tma_launch(const float* fts, long long ep_stride, const float* ctr, const float* fg, const float* bg,
                        long long mask_stride, int B, int S, int hw, float eps, float* fg_proto, float* bg_proto,
                        float* adaptive_p, float* shot_centre, float* shot_den, char* ws, size_t ws_bytes, cudaStream_t st);
size_t pemp_mpa_tma_workspace_bytes(int B, int S, int hw);
#ifndef PEMP_MPA_TMA
#define PEMP_MPA_TMA 1
#endif
static inline bool mpa_tma_shape(int c, int hw, int p) { return PEMP_MPA_TMA && c == 512 && p == 3 && hw >= 32; }
static int g_mpa_path = 0;   // diagnostic switch, see pemp_debug_mpa_path

namespace {

// ---- cluster handshake without fences ---------------------------------------------------------------------
// `cluster.sync()` compiles to MEMBAR + barrier + CCTL.IVALL: the fence waits for every outstanding global load of
// the thread, which serialises the register prefetch of the next tile behind the barrier.  The per-tile exchange of
// the partial dots therefore uses `st.async` into the peer's shared memory, completing a transaction mbarrier there.
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t map_to_peer(uint32_t local_addr, uint32_t peer_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(peer_rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init_local(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_local(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_local(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MPA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MPA_DONE;\n"
      "bra MPA_WAIT;\n"
      "MPA_DONE:\n"
      "}\n" ::"r"(smem_addr_u32(bar)),
      "r"(parity)
      : "memory");
}

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxChannels = 3072;     // per-CTA Ft[32*CS][c/CS + 4] must fit the 227 KB of shared memory

#ifndef PEMP_MPA_QPT
#define PEMP_MPA_QPT 2
#endif
#ifndef PEMP_MPA_LOADS
#define PEMP_MPA_LOADS 64              // independent row loads in flight per lane in phase A (specialised shapes)
#endif
#ifndef PEMP_MPA_LOADS_GENERIC
#define PEMP_MPA_LOADS_GENERIC 32      // ... run-time shapes need registers for address arithmetic
#endif
#ifndef PEMP_MPA_WAVES
#define PEMP_MPA_WAVES 4
#endif
#ifndef PEMP_MPA_CLUSTER
#define PEMP_MPA_CLUSTER 2
#endif
#ifndef PEMP_MPA_PREFETCH
#define PEMP_MPA_PREFETCH 0
#endif
constexpr int kQPT = PEMP_MPA_QPT;     // channel quads per phase-B thread
#ifdef PEMP_MPA_DEBUG_SKIP_B           // timing experiments only (results are wrong)
constexpr bool kDebugSkipB = true;
#else
constexpr bool kDebugSkipB = false;
#endif
#ifdef PEMP_MPA_DEBUG_SKIP_DOTS
constexpr bool kDebugSkipDots = true;
#else
constexpr bool kDebugSkipDots = false;
#endif

__host__ __device__ inline int nd_of(int P) { return 2 * (P - 1); }          // dot products per pixel
__host__ __device__ inline int ndp_of(int P) { return P <= 3 ? 4 : 8; }      // padded table row

// threads per phase-B pixel group for cc channels: power of two >= 32 (a warp never straddles groups)
__host__ __device__ inline int tpp_of(int cc) {
  int need = ((cc >> 2) + kQPT - 1) / kQPT, t = 32;
  while (t < need) t <<= 1;
  return t;
}

struct Smem {            // offsets in floats
  int ldf, ft, table, red, exin, wgt, total;
};
__host__ __device__ inline Smem smem_layout(int cc, int P, int CS) {
  const int TW = 32 * CS, NDP = ndp_of(P), ND = nd_of(P);
  Smem s;
  s.ldf = 4 * kQPT * tpp_of(cc) + 4;                           // >= cc, = 4 * odd
  int tile = TW * s.ldf, fold = kThreads * kQPT * 2 * 4 * P;   // the epilogue reuses Ft as fold[ng][kQPT*8*P][tpp]
  s.ft = 0;
  s.table = tile > fold ? tile : fold;
  s.red = s.table + (ND ? cc * NDP : 0);
  s.exin = s.red + (ND ? kWarps * NDP * TW : 0);             // the peer's partial dots land here (st.async), [2][NDP][TW]
  s.wgt = s.exin + ((ND && CS > 1) ? 2 * NDP * TW : 0);
  s.total = s.wgt + 2 * TW * 4;
  return s;
}

// ---- prologue: difference table ------------------------------------------------------------------
// table[c][ndp]: column (g*(P-1) + j-1) = 2*(ctr[c, g*P+j] - ctr[c, g*P]);  konst[g*(P-1)+j-1] =
// -(|ctr_{g*P+j}|^2 - |ctr_{g*P}|^2), accumulated in double.
__global__ void mpa_prepare_kernel(const float* __restrict__ ctr, int c, int P, float* __restrict__ table,
                                   float* __restrict__ konst) {
  const int nd = nd_of(P), ndp = ndp_of(P);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c * ndp; i += gridDim.x * blockDim.x) {
    int ch = i / ndp, d = i - ch * ndp;
    float v = 0.f;
    if (d < nd) {
      int g = d / (P - 1), j = d - g * (P - 1) + 1;
      v = 2.0f * (ctr[ch * 2 * P + g * P + j] - ctr[ch * 2 * P + g * P]);
    }
    table[i] = v;
  }
  if (blockIdx.x == 0) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nd) {
      int g = warp / (P - 1), j = warp - g * (P - 1) + 1;
      double s = 0.0;
      for (int ch = lane; ch < c; ch += 32) {
        double a = ctr[ch * 2 * P + g * P + j], b = ctr[ch * 2 * P + g * P];
        s += (a - b) * (a + b);
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
      if (lane == 0) konst[warp] = static_cast<float>(-s);
    }
  }
}

// ---- main kernel -------------------------------------------------------------------------------------
// HWT / CCT > 0 fix the pixel count and the per-CTA channel count at compile time (the PEMP shapes): every row
// offset, shared-memory stride and trip count becomes an immediate, which removes about half of the
// instructions of the generic version (address arithmetic).  SAFE = true clamps pixel indices per lane and is
// used when the image is narrower than one tile; otherwise the last tile is shifted left to end at hw and the
// pixels it shares with the previous tile get zero weight.
template <int P, int CS, int HWT, int CCT, bool SAFE, bool PF>
__global__ void __launch_bounds__(kThreads, 2)
mpa_kernel(const float* __restrict__ fts, long long ep_stride, int S, const float* __restrict__ table_g,
           const float* __restrict__ konst_g, const float* __restrict__ fg, const float* __restrict__ bg,
           long long mask_stride, int c_rt, int hw_rt, float* __restrict__ part_num, float* __restrict__ part_den) {
  constexpr int ND = 2 * (P - 1);
  constexpr int NDP = P <= 3 ? 4 : 8;
  constexpr int K = 2 * P;
  constexpr int TW = 32 * CS;                      // pixels per tile
  constexpr int U = (HWT > 0 ? PEMP_MPA_LOADS : PEMP_MPA_LOADS_GENERIC) / (4 * CS);   // quads per load batch
  extern __shared__ __align__(16) float smem[];
  const int hw = HWT > 0 ? HWT : hw_rt;
  const int cc = CCT > 0 ? CCT : c_rt / CS;        // channels owned by this CTA
  const int c = cc * CS;
  const Smem L = smem_layout(cc, P, CS);
  const int ldf = L.ldf;
  float* Ft = smem + L.ft;                         // [TW][ldf]
  float* table = smem + L.table;                   // [cc][NDP]      rows of this CTA's channels
  float* red = smem + L.red;                       // [kWarps][NDP][TW]
  float* exin = smem + L.exin;                     // [2][NDP][TW]   partial dots of the peer CTA
  float* wgt = smem + L.wgt;                       // [2][TW][4]     softmax * mask
  __shared__ float konst[8];
  __shared__ unsigned live_mask[2][CS];            // bit x set <=> group g has a non-zero weight at pixel x
  __shared__ float den_part[2][CS][4];
  __shared__ __align__(8) uint64_t xbar[2];        // transaction barriers of the two exchange buffers

  const int rank = CS > 1 ? static_cast<int>(cg::this_cluster().block_rank()) : 0;
  const int split = blockIdx.x / CS, nsplit = gridDim.x / CS, img = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = (hw + TW - 1) / TW;
  // cluster `split` of an image walks tiles split, split + nsplit, ...: the clusters of one image (adjacent
  // block ids, co-resident) touch adjacent pieces of every channel row at about the same time
  const int my_tiles = split < ntiles ? (ntiles - split + nsplit - 1) / nsplit : 0;
  const int ch0 = rank * cc;
  const int quads = cc >> 2;

  if (ND) {
    for (int i = tid; i < cc * NDP; i += kThreads) table[i] = __ldg(table_g + ch0 * NDP + i);
    if (tid < ND) konst[tid] = __ldg(konst_g + tid);
  }
  // channels beyond cc in the padded Ft rows are never written by phase A: clear them once
  for (int i = tid; i < TW * (ldf - cc); i += kThreads) Ft[(i / (ldf - cc)) * ldf + cc + i % (ldf - cc)] = 0.f;

  const float* img_base = fts + (img / S) * ep_stride + (static_cast<long long>(img % S) * c + ch0) * hw;
  const float* fgp = fg + img * mask_stride;
  const float* bgp = bg + img * mask_stride;
  uint32_t peer_exin = 0, peer_xbar = 0;
  if (CS > 1 && ND) {
    if (tid == 0) {
      mbar_init_local(&xbar[0], 1);
      mbar_init_local(&xbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    peer_exin = map_to_peer(smem_addr_u32(exin), rank ^ 1);
    peer_xbar = map_to_peer(smem_addr_u32(&xbar[0]), rank ^ 1);
    cg::this_cluster().sync();                     // barriers initialised before any remote store can arrive
  }

  // phase-A ownership: qw contiguous quads per warp
  const int qw = (quads + kWarps - 1) / kWarps;
  const int q_lo = warp * qw;
  const int nq = max(0, min(quads - q_lo, qw));    // == qw for every warp when kWarps divides quads
  // phase-B ownership (tpp is a power of two)
  const int tpp = tpp_of(cc);
  const int tpp_shift = 31 - __clz(tpp);
  const int ng = kThreads >> tpp_shift, ppg = TW / ng;
  const int qb = tid & (tpp - 1), grp = tid >> tpp_shift;
  float acc[kQPT][2][4][P];                                // [quad slot][group][channel][prototype]
#pragma unroll
  for (int a = 0; a < kQPT; ++a)
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < P; ++j) acc[a][g][e][j] = 0.f;
  float den[P];   // softmax threads: denominators of their group, summed over their pixels
#pragma unroll
  for (int j = 0; j < P; ++j) den[j] = 0.f;

  __syncthreads();

  // PF (specialised shapes, one load batch per tile): the row loads of tile it+1 are issued into registers as soon
  // as tile it's values have been consumed, so they are in flight during the exchange / softmax / phase B of tile it.
  float vpf[PF ? U : 1][4][CS];
  auto issue_loads = [&](int tile_it) {
    const int xn = (split + tile_it * nsplit) * TW;
    const int xw = min(xn, hw - TW);
    const float* pr = img_base + static_cast<long long>(q_lo * 4) * hw + xw + lane;
#pragma unroll
    for (int u = 0; u < (PF ? U : 0); ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int s2 = 0; s2 < CS; ++s2) vpf[u][e][s2] = __ldg(pr + (u * 4 + e) * hw + 32 * s2);
  };
  if (PF && my_tiles > 0) issue_loads(0);

  for (int it = 0; it < my_tiles; ++it) {
    const int x_nom = (split + it * nsplit) * TW;            // nominal first pixel of the tile
    const int x0 = SAFE ? x_nom : min(x_nom, hw - TW);       // window actually loaded
    const int buf = it & 1;

    // ---------------- phase A: load, dot with the difference table, transpose into Ft ---------------
    float pd[CS][NDP];
#pragma unroll
    for (int s = 0; s < CS; ++s)
#pragma unroll
      for (int d = 0; d < NDP; ++d) pd[s][d] = 0.f;
    int xoff[CS];                                            // SAFE: per-lane clamped pixel offsets
#pragma unroll
    for (int s = 0; s < CS; ++s) xoff[s] = SAFE ? min(x0 + lane + 32 * s, hw - 1) : x0 + lane + 32 * s;
    // consecutive channel rows are one `hw` stride apart: one running pointer, immediate offsets inside a batch
    const float* prow = img_base + static_cast<long long>(q_lo * 4) * hw + (SAFE ? 0 : xoff[0]);
    float* fst = Ft + lane * ldf + q_lo * 4;
    const float4* trow = reinterpret_cast<const float4*>(table + q_lo * 4 * NDP);
    for (int i = 0; i < nq; i += U) {
      float v[U][4][CS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (PF) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int s = 0; s < CS; ++s) v[u][e][s] = vpf[u][e][s];
        } else if (i + u < nq) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int s = 0; s < CS; ++s)
              v[u][e][s] = SAFE ? __ldg(prow + (u * 4 + e) * hw + xoff[s]) : __ldg(prow + (u * 4 + e) * hw + 32 * s);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i + u < nq) {
          if (ND && !kDebugSkipDots) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 ta = trow[(u * 4 + e) * (NDP / 4)];
              float4 tb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ND > 4) tb = trow[(u * 4 + e) * (NDP / 4) + 1];
#pragma unroll
              for (int s = 0; s < CS; ++s) {
                pd[s][0] = fmaf(v[u][e][s], ta.x, pd[s][0]);
                pd[s][1] = fmaf(v[u][e][s], ta.y, pd[s][1]);
                if (ND > 2) {
                  pd[s][2] = fmaf(v[u][e][s], ta.z, pd[s][2]);
                  pd[s][3] = fmaf(v[u][e][s], ta.w, pd[s][3]);
                }
                if (ND > 4) {
                  pd[s][4] = fmaf(v[u][e][s], tb.x, pd[s][4]);
                  pd[s][5] = fmaf(v[u][e][s], tb.y, pd[s][5]);
                }
              }
            }
          }
#pragma unroll
          for (int s = 0; s < CS; ++s)
            *reinterpret_cast<float4*>(fst + 32 * s * ldf + u * 4) =
                make_float4(v[u][0][s], v[u][1][s], v[u][2][s], v[u][3][s]);
        }
      }
      prow += static_cast<long long>(U * 4) * hw;
      fst += U * 4;
      trow += U * 4 * (NDP / 4);
    }
    if (ND) {
#pragma unroll
      for (int s = 0; s < CS; ++s)
#pragma unroll
        for (int d = 0; d < ND; ++d) red[(warp * NDP + d) * TW + lane + 32 * s] = pd[s][d];
    }
    if (PF && it + 1 < my_tiles) issue_loads(it + 1);
    __syncthreads();

    // ---------------- exchange + softmax: threads [0, TW) foreground group, [TW, 2*TW) background group ----
    // Thread (g, xl) owns the P-1 dots of group g at pixel xl: it adds the warps' partials (fixed order), sends the
    // sums to the peer CTA (st.async completes the peer's transaction barrier) and waits for the peer's sums.
    if (tid < 2 * TW) {
      const int g = tid / TW, xl = tid - g * TW, x = x0 + xl;
      float mine[P > 1 ? P - 1 : 1];
#pragma unroll
      for (int j = 1; j < P; ++j) {
        const int d = g * (P - 1) + j - 1;
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < kWarps; ++wv) s += red[(wv * NDP + d) * TW + xl];
        mine[j - 1] = s;
        if (CS > 1) st_async_f32(peer_exin + ((buf * NDP + d) * TW + xl) * 4, s, peer_xbar + buf * 8);
      }
      if (CS > 1 && ND) {
        if (tid == 0) mbar_expect_tx_local(&xbar[buf], ND * TW * 4);              // arm this tile's incoming transfer
        mbar_wait_local(&xbar[buf], (it >> 1) & 1);                               // the peer's dots have landed
      }
      // pixels before x_nom were already handled by the previous tile of the shifted last window
      const float m = (x < hw && x >= x_nom) ? __ldg((g ? bgp : fgp) + x) : 0.f;
      float e[P];
      e[0] = 0.f;
      float mx = 0.f;
#pragma unroll
      for (int j = 1; j < P; ++j) {
        const int d = g * (P - 1) + j - 1;
        float s = mine[j - 1];
        if (CS > 1) {   // fixed order rank 0 + rank 1 on both CTAs => identical weights
          const float theirs = exin[(buf * NDP + d) * TW + xl];
          s = rank == 0 ? s + theirs : theirs + s;
        }
        e[j] = s + konst[d];
        mx = fmaxf(mx, e[j]);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        e[j] = expf(e[j] - mx);
        sum += e[j];
      }
      float w4[4] = {0.f, 0.f, 0.f, 0.f};
      bool any = false;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        w4[j] = (e[j] / sum) * m;
        den[j] += w4[j];
        any |= w4[j] != 0.f;
      }
      *reinterpret_cast<float4*>(wgt + (g * TW + xl) * 4) = make_float4(w4[0], w4[1], w4[2], w4[3]);
      const unsigned bal = __ballot_sync(kFull, any);
      if (lane == 0) live_mask[g][xl >> 5] = bal;
    }
    __syncthreads();

    // ---------------- phase B: out[c, k] += f[c, x] * A[k, x] over the pixels whose weights are non-zero
    // (warp-uniform pixel lists from the bit masks; two pixels per step for load/FMA overlap)
    const float* fcol = Ft + qb * 4;
#pragma unroll
    for (int g = 0; g < (kDebugSkipB ? 0 : 2); ++g) {
      for (int pbase = grp * ppg; pbase < (grp + 1) * ppg; pbase += 32) {
        const int nbits = min(32, (grp + 1) * ppg - pbase);
        unsigned m = (live_mask[g][pbase >> 5] >> (pbase & 31)) & (nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1u));
        while (m) {
          const int i0 = __ffs(m) - 1;
          m &= m - 1;
          const bool two = m != 0;
          const int i1 = two ? __ffs(m) - 1 : i0;
          m &= m - 1;
          const int p0 = pbase + i0, p1 = pbase + i1;
          const float4 w0 = *reinterpret_cast<const float4*>(wgt + (g * TW + p0) * 4);
          float4 w1 = *reinterpret_cast<const float4*>(wgt + (g * TW + p1) * 4);
          if (!two) w1 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float wa[4] = {w0.x, w0.y, w0.z, w0.w}, wb[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int a = 0; a < kQPT; ++a) {
            const float4 f0 = *reinterpret_cast<const float4*>(fcol + p0 * ldf + a * tpp * 4);
            const float4 f1 = *reinterpret_cast<const float4*>(fcol + p1 * ldf + a * tpp * 4);
            const float fa[4] = {f0.x, f0.y, f0.z, f0.w}, fb[4] = {f1.x, f1.y, f1.z, f1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int j = 0; j < P; ++j) {
                acc[a][g][e][j] = fmaf(fa[e], wa[j], acc[a][g][e][j]);
                acc[a][g][e][j] = fmaf(fb[e], wb[j], acc[a][g][e][j]);
              }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---------------- epilogue: fold the pixel groups (fixed order), write the partials ------------------
  float* fold = Ft;   // reuse: [ng][kQPT*2*4*P][tpp]
  constexpr int kAcc = kQPT * 2 * 4 * P;
  {
    int o = 0;
#pragma unroll
    for (int a = 0; a < kQPT; ++a)
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int j = 0; j < P; ++j) fold[(grp * kAcc + (o++)) * tpp + qb] = acc[a][g][e][j];
  }
  if (tid < 2 * TW) {   // denominators: per-warp sums, combined below in warp order
    const int g = tid / TW, w_in_g = (tid - g * TW) >> 5;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float s = warp_sum(den[j]);
      if (lane == 0) den_part[g][w_in_g][j] = s;
    }
  }
  __syncthreads();
  if (grp == 0) {
    float* out = part_num + ((static_cast<long long>(img) * nsplit + split) * c + ch0) * K;
    int o = 0;
#pragma unroll
    for (int a = 0; a < kQPT; ++a) {
      const int q = qb + a * tpp;
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int j = 0; j < P; ++j) {
            float v = 0.f;
            for (int r = 0; r < ng; ++r) v += fold[(r * kAcc + o) * tpp + qb];
            ++o;
            if (q < quads) out[(q * 4 + e) * K + g * P + j] = v;
          }
    }
  }
  if (rank == 0 && tid < K) {
    const int g = tid / P, j = tid - g * P;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < CS; ++wv) s += den_part[g][wv][j];
    part_den[(static_cast<long long>(img) * nsplit + split) * K + tid] = s;
  }
  if (CS > 1) cg::this_cluster().sync();   // remote stores into the peer (and the peer's into us) have all landed
}

// one thread per (b, channel, k)
__global__ void mpa_finalize_kernel(const float* __restrict__ part_num, const float* __restrict__ part_den, int B, int S,
                                    int c, int P, int nsplit, float eps, float* __restrict__ fg_proto,
                                    float* __restrict__ bg_proto, float* __restrict__ adaptive_p,
                                    float* __restrict__ shot_centre, float* __restrict__ shot_den) {
  const int K = 2 * P;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * c * K) return;
  int k = static_cast<int>(i % K);
  long long t = i / K;
  int ch = static_cast<int>(t % c);
  int b = static_cast<int>(t / c);
  float accum = 0.f;
  for (int s = 0; s < S; ++s) {
    long long img = static_cast<long long>(b) * S + s;
    float num = 0.f, den = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) {
      num += part_num[((img * nsplit + sp) * c + ch) * K + k];
      den += part_den[(img * nsplit + sp) * K + k];
    }
    accum += num / (den + eps);
    if (shot_centre) {      // training forward: per-shot centres [BS, c, 2P] and denominators [BS, 2P] for the backward
      shot_centre[(img * c + ch) * K + k] = num / (den + eps);
      if (ch == 0) shot_den[img * K + k] = den + eps;
    }
  }
  float v = accum / static_cast<float>(S);
  int g = k / P, j = k - g * P;
  (g == 0 ? fg_proto : bg_proto)[(static_cast<long long>(b) * c + ch) * P + j] = v;
  if (adaptive_p) adaptive_p[(static_cast<long long>(b) * c + ch) * K + k] = v;
}

struct Plan {
  int cs, nsplit;
  size_t smem_bytes, off_table, off_konst, off_num, off_den, total;
};
Plan make_plan(int B, int S, int c, int hw, int P) {
  Plan p;
  const size_t imgs = static_cast<size_t>(B) * S;
  // CTA pairs need an even split of whole quads; otherwise one CTA owns all channels
  p.cs = (PEMP_MPA_CLUSTER == 2 && c % 8 == 0 && c >= 64) ? 2 : 1;
  const int ntiles = (hw + 32 * p.cs - 1) / (32 * p.cs);
  // aim for PEMP_MPA_WAVES waves of 2 CTAs/SM on 148 SMs; at least 2 tiles per split
  int want = static_cast<int>((PEMP_MPA_WAVES * 148 * 2 + imgs * p.cs - 1) / (imgs * p.cs));
  int cap = ntiles / 2 > 0 ? ntiles / 2 : 1;
  p.nsplit = want < cap ? want : cap;
  if (p.nsplit < 1) p.nsplit = 1;
  p.smem_bytes = static_cast<size_t>(smem_layout(c / p.cs, P, p.cs).total) * sizeof(float);
  p.off_table = 0;
  p.off_konst = align_up(static_cast<size_t>(c) * ndp_of(P) * sizeof(float), 256);
  p.off_num = p.off_konst + 256;
  p.off_den = p.off_num + align_up(imgs * p.nsplit * c * 2 * P * sizeof(float), 256);
  p.total = p.off_den + align_up(imgs * p.nsplit * 2 * P * sizeof(float), 256);
  return p;
}

template <int P, int CS, int HWT, int CCT, bool SAFE, bool PF>
int launch(const float* fts, long long ep_stride, const float* ctr, const float* fg, const float* bg,
           long long mask_stride, int B, int S, int c, int hw, float eps, float* fg_proto, float* bg_proto,
           float* adaptive_p, float* shot_centre, float* shot_den, char* ws, const Plan& pl, cudaStream_t st) {
  constexpr int ND = 2 * (P - 1);
  float* table = reinterpret_cast<float*>(ws + pl.off_table);
  float* konst = reinterpret_cast<float*>(ws + pl.off_konst);
  float* num = reinterpret_cast<float*>(ws + pl.off_num);
  float* den = reinterpret_cast<float*>(ws + pl.off_den);
  if (ND) mpa_prepare_kernel<<<8, 256, 0, st>>>(ctr, c, P, table, konst);
  if (pl.smem_bytes > 227 * 1024) return PEMP_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(mpa_kernel<P, CS, HWT, CCT, SAFE, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(pl.smem_bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(pl.nsplit) * CS, static_cast<unsigned>(B) * S);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const long long eps_stride = ep_stride ? ep_stride : static_cast<long long>(S) * c * hw;
  e = cudaLaunchKernelEx(&cfg, mpa_kernel<P, CS, HWT, CCT, SAFE, PF>, fts, eps_stride, S, static_cast<const float*>(table),
                         static_cast<const float*>(konst), fg, bg, mask_stride, c, hw, num, den);
  if (e != cudaSuccess) return static_cast<int>(e);
  long long total = static_cast<long long>(B) * c * 2 * P;
  mpa_finalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(num, den, B, S, c, P, pl.nsplit, eps,
                                                                                 fg_proto, bg_proto, adaptive_p, shot_centre, shot_den);
  return launch_status();
}

}  // namespace

extern "C" int pemp_debug_mpa_path(int mode) {
  const int old = g_mpa_path;
  if (mode == 0 || mode == 1) g_mpa_path = mode;
  return old;
}

extern "C" size_t pemp_meta_proto_attn_workspace_bytes(int B, int S, int c, int hw, int p) {
  if (B <= 0 || S <= 0 || c <= 0 || hw <= 0 || p < 1 || p > 4) return 0;
  size_t n = make_plan(B, S, c, hw, p).total;
  if (mpa_tma_shape(c, hw, p)) {
    const size_t t = pemp_mpa_tma_workspace_bytes(B, S, hw);
    if (t > n) n = t;
  }
  return n;
}

static int mpa_entry(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg,
                     const float* bg, long long mask_stride, int B, int S, int c, int hw, int p,
                     float eps, float* fg_proto, float* bg_proto, float* adaptive_p, float* shot_centre, float* shot_den,
                     void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(fts && ctr && fg && bg && fg_proto && bg_proto, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && c > 0 && hw > 0 && static_cast<long long>(B) * S <= 65535, PEMP_E_SHAPE);
  PEMP_REQUIRE(p >= 1 && p <= 4 && c % 4 == 0 && c <= kMaxChannels, PEMP_E_SHAPE);
  Plan pl = make_plan(B, S, c, hw, p);
  PEMP_REQUIRE(pl.smem_bytes <= 227 * 1024, PEMP_E_SHAPE);
  PEMP_REQUIRE(workspace && workspace_bytes >= pemp_meta_proto_attn_workspace_bytes(B, S, c, hw, p), PEMP_E_WORKSPACE);
  PEMP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, PEMP_E_ALIGN);
  char* ws = static_cast<char*>(workspace);
  cudaStream_t st = as_stream(stream);
  if (mpa_tma_shape(c, hw, p) && g_mpa_path != 1) {
    const int rc = pemp_mpa_tma_launch(fts, fts_episode_stride, ctr, fg, bg, mask_stride, B, S, hw, eps, fg_proto, bg_proto,
                                       adaptive_p, shot_centre, shot_den, ws, workspace_bytes, st);
    if (rc != PEMP_E_ALIGN) return rc;
  }
#define PEMP_MPA_ARGS \
  fts, fts_episode_stride, ctr, fg, bg, mask_stride, B, S, c, hw, eps, fg_proto, bg_proto, adaptive_p, shot_centre, shot_den, ws, pl, st
  const bool safe = hw < 32 * pl.cs;
  // fully specialised PEMP shape: c = 512, 51 x 51 features, 3 prototypes per class
  if (p == 3 && pl.cs == 2 && c == 512 && hw == 2601) return launch<3, 2, 2601, 256, false, PEMP_MPA_PREFETCH != 0>(PEMP_MPA_ARGS);
#define PEMP_MPA(PP)                                                                            \
  return pl.cs == 2 ? (safe ? launch<PP, 2, 0, 0, true, false>(PEMP_MPA_ARGS) : launch<PP, 2, 0, 0, false, false>(PEMP_MPA_ARGS)) \
                    : (safe ? launch<PP, 1, 0, 0, true, false>(PEMP_MPA_ARGS) : launch<PP, 1, 0, 0, false, false>(PEMP_MPA_ARGS))
  switch (p) {
    case 1: PEMP_MPA(1);
    case 2: PEMP_MPA(2);
    case 3: PEMP_MPA(3);
    default: PEMP_MPA(4);
  }
#undef PEMP_MPA_ARGS
#undef PEMP_MPA
}

extern "C" int pemp_meta_proto_attn(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg,
                                    const float* bg, long long mask_stride, int B, int S, int c, int hw, int p,
                                    float eps, float* fg_proto, float* bg_proto, float* adaptive_p, void* workspace,
                                    size_t workspace_bytes, pemp_stream_t stream) {
  return mpa_entry(fts, fts_episode_stride, ctr, fg, bg, mask_stride, B, S, c, hw, p, eps, fg_proto, bg_proto, adaptive_p,
                   nullptr, nullptr, workspace, workspace_bytes, stream);
}

// Training forward: the same kernels, and the per-shot centres [BS, c, 2p] (fg columns first) and denominators
// (sum of attention + eps) [BS, 2p] that pemp_meta_proto_attn_bwd needs.
extern "C" int pemp_meta_proto_attn_train(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg,
                                          const float* bg, long long mask_stride, int B, int S, int c, int hw, int p,
                                          float eps, float* fg_proto, float* bg_proto, float* shot_centre, float* shot_den,
                                          void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(shot_centre && shot_den, PEMP_E_NULL);
  return mpa_entry(fts, fts_episode_stride, ctr, fg, bg, mask_stride, B, S, c, hw, p, eps, fg_proto, bg_proto, nullptr,
                   shot_centre, shot_den, workspace, workspace_bytes, stream);
}
