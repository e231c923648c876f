// K12: backward of meta-prototype attention with the three thin products of a tile on the warp-level tensor
// path (`mma.sync.m16n8k8` tf32, SASS HMMA.1688), 3 x TF32 split = fp32-grade, the tile fed by TMA.  Same math, inputs, outputs
// and partial layout as `mpa_bwd_kernel` (train.cu; reference: what autograd records for networks/pemp_stage1.py:202-213),
// P = 3, c in {128, 256, 512, 1024}, hw >= 32, TMA-encodable operand; everything else takes the CUDA-core kernel.
//
// Why.  Per pixel tile the backward is three products with a 10-column table  T[ch] = { coef[ch][0..6) | ctr_k - ctr_g0 }:
//   phase A   dots[k][x]  = sum_ch T[ch][k] f[ch][x]            (10 x 32, contraction over the c channels)
//   phase B1  df[x][ch]   = sum_k  W[x][k]  T[ch][k]            (32 x c,  contraction over the 10 columns; W from the pixel step)
//   phase B2  dctr[ch][k] = sum_x  f[ch][x] dv[x][k]            (c x 6,   contraction over the pixels)
// On CUDA cores every FMA of A and B1 needs a table value that is uniform over the warp (lane = pixel), and a broadcast
// LDS.128 still costs four cycles of the 128 B/clk shared-memory crossbar: 12 cycles per channel and warp in A and again in B1
// = ~14 800 crossbar cycles per tile against the ~5 300 clocks the tile's 128 KB take at the HBM rate - the round-2 kernel sat
// at 0.31-0.33 of the HBM peak with the shared pipe 63 % busy.  With register fragments a table / weight value is read once per
// 8 x 8 block (~3 000 crossbar cycles per tile).  tcgen05 is not an option here: its operands live in shared memory behind
// descriptors (T as hi + lo K-major panels = 48-64 KB next to the 64-KB tile, twice per SM) and the three products contract
// over three different axes of the same tile; ~1 700 m16n8k8 per tile keep the legacy pipe (measured 2.0 clk per instruction
// and SM, tools/probes/mma_sync_probe.cu) a third busy.
//
// 3 x TF32: x = hi + lo with hi = x & 0xffffe000 (what the tensor core reads of an fp32 register) and lo = x - hi (exact);
// a.b ~ lo_a hi_b + hi_a lo_b + hi_a hi_b, small products first (see K9), error ~2^-21 per product.
//
// Structure (c = 512, the PEMP shape; the configurations are listed below): ONE CTA per SM with sixteen product warps + two
// pixel-step warps (96 registers); a CTA owns a run of 28-pixel tiles of one image (grid = images x chunks, train.cu plans the
// split).  The features come through the "four rows per group" tensor map of the forward kernels (tma_common.cuh): product
// warp w = e + 4 q owns rows [CW q, CW q + CW) of the class-e box of a tile (channels 4 g + e; CW = c / 16), i.e. its own
// 128-byte-swizzled quarter box, double buffered, each buffer with its own mbarrier - the tile is warp-private in every phase
// and the box of tile t+1 is requested before tile t is touched (the first version filled a single buffer with 64 four-byte
// cp.async per thread: a third of all stall samples sat in that loop; single-buffered TMA boxes still left the product warps
// waiting for the refill a quarter of the time).  Box column i of class e is pixel x_nom + i - o_e, o_e = (e hw + x_nom) & 3
// (aligned box origins).  Per tile a product warp runs
//     request box(t+1) | wait box(t) | A(t) -> dots -> arrive | B1(t-1) | wait weights(t) | B2(t)
// and the pixel-step warps (one per class group, lane = pixel) turn the sixteen partial dot sets into the weights of the tile
// (double buffered by tile parity) behind two mbarriers: B1 needs no features, so it hides the pixel step, and no warp ever
// waits at a block barrier in the tile loop (when two of the product warps did the pixel step between two __syncthreads,
// those were 16 % of all stall samples).  The single-buffered configurations split B1 in two halves around B2 instead: one
// hides the pixel step, the other the refill.
#include <math_constants.h>

#include "tma_common.cuh"

size_t pemp_mpa_bwd_mma_smem(int c);
bool pemp_mpa_bwd_mma_shape(int c, int p, int hw);
int pemp_mpa_bwd_mma_tiles(int hw);
size_t pemp_mpa_bwd_mma_table_bytes(int N, int c);
int pemp_mpa_bwd_mma_table_ld();
int pemp_mpa_bwd_mma_rows_per_warp(int c);
int pemp_mpa_bwd_mma_launch(const float* fts, long long ep, int B, int S, const float* ctr, const float* coef, const float* beta,
                            const float* fg, const float* bg, long long mask_stride, int c, int hw, int chunks, float* tabg,
                            float* dfts, long long d_ep, float* part, float* img_part, cudaStream_t st);

namespace {

using namespace pemp_tma;

// A configuration is (MB, KW, NBUF): KW product warps (+ two pixel-step warps, one per class group), warp w = e + 4 q owns rows
// [CW q, CW q + CW) of the class-e box with CW = 16 MB = c / KW, in NBUF buffers:
//   c = 512 / 256: KW = 16, NBUF = 2, one CTA per SM: the box of tile t+1 is requested before tile t is touched
//   c = 128 / 1024: KW = 8, NBUF = 1 (two CTAs per SM at c = 128): the box is refilled piecewise as soon as B2 is done with it
constexpr int kP = 3, kK = 2 * kP;               // prototypes per group, coefficient columns
constexpr int kND = 2 * (kP - 1);                // centre-difference columns (the first prototype of a group has none)
constexpr int kNK = kK + kND;                    // 10 table columns: [0, 6) coef, [6, 8) fg differences, [8, 10) bg differences
constexpr int kTLd = 12;                         // table row pitch: conflict-free for the fragments of phase A and of B1
constexpr int kStep = 28;                        // pixels a tile advances (32 box columns cover them for every class)
constexpr int kRedLd = 33;                       // pixel pitch of a dot row in `red`: a fragment store touches banks g + 8 tg
constexpr int kWtLd = 12;                        // W[x][0..10) + two zero columns (conflict-free fragment reads)
constexpr int kStgLd = 36;                       // pixel pitch of a staged gradient row (conflict-free fragment stores)
constexpr int kDvRows = 32 + 3;                  // dv rows for pixels -3 .. 31 of a tile (rows outside [0, 28) stay zero)
static_assert(kNK > 8 && kNK <= kTLd, "two 8-wide column blocks, the second one inside the row padding");

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// fragments: A 16 x 8 (a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)), B 8 x 8 (b0 (t, g) b1 (t+4, g)),
// C 16 x 8 (c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)) with g = lane >> 2, t = lane & 3.  The contraction index may
// be permuted freely as long as both operands agree: phases A and B1 use it that way to stay bank-conflict free.
struct FragA {
  uint32_t hi[4], lo[4];
};
struct FragB {
  uint32_t hi[2], lo[2];
};
__device__ __forceinline__ void mma3(float (&d)[4], const FragA& a, const FragB& b) {
  mma_tf32(d, a.lo[0], a.lo[1], a.lo[2], a.lo[3], b.hi[0], b.hi[1]);
  mma_tf32(d, a.hi[0], a.hi[1], a.hi[2], a.hi[3], b.lo[0], b.lo[1]);
  mma_tf32(d, a.hi[0], a.hi[1], a.hi[2], a.hi[3], b.hi[0], b.hi[1]);
}
// (a 128-byte-swizzled box: 16-byte chunk j of row r sits at chunk j ^ (r & 7); rows are 32 floats)

template <int MB, int KW, int NBUF>              // MB: 16-row blocks per warp
struct Smem {
  static constexpr int CW = 16 * MB, c = CW * KW;
  alignas(1024) float tile[NBUF][KW][CW * 32];   // warp w: rows [CW q, CW q + CW) of the class-e box, e = w & 3, q = w >> 2
  alignas(16) float tab[c * kTLd];               // row R = w CW + r  <->  channel 4 (CW half + r) + e
  alignas(16) float red[KW][kNK * kRedLd];       // partial dots of the warps: [k][pixel of the tile]
  alignas(16) float wt[2][32 * kWtLd];           // [tile parity][pixel]{ a_k (6) | 2 dl_k of the non-first prototypes (4) | 0 0 }
  alignas(16) float dv[kDvRows * 8];             // [pixel + 3]{ 2 dl_k (6) | 0 0 }
  alignas(16) float stage[KW][8 * kStgLd];       // per warp: one 8-row block of the gradient tile on its way out
  float konst[2 * kK];                           // |ctr_k|^2 - |ctr_g0|^2, beta
  alignas(8) uint64_t full[NBUF][KW];            // per product warp and buffer: the box has landed
  alignas(8) uint64_t part_bar;                  // all product warps have written their dots of a tile
  alignas(8) uint64_t wts_bar;                   // both pixel-step warps have written the weights of a tile
};

static_assert(sizeof(Smem<2, 16, 2>) <= 227 * 1024, "c = 512, double buffered, one CTA per SM");

// HW: the map size as a compile-time constant (0 = take the argument).  The gradient rows of a block are channels 4 apart, i.e.
// 16 hw bytes: with hw known the eight stores of a block use immediate offsets instead of a 64-bit pointer bump each (the
// bumps and the block addressing were 11 % of all instructions, ncu); instantiated for the PEMP map, 51 x 51.
template <int MB, int KW, int NBUF, int HW>
__global__ void __launch_bounds__((KW + 2) * 32, (MB == 1 && KW == 8) ? 2 : 1)
mpa_bwd_mma_kernel(const __grid_constant__ CUtensorMap map, int S, const float* __restrict__ tabg,
                   const float* __restrict__ beta, const float* __restrict__ fg, const float* __restrict__ bg,
                   long long mask_stride, int hw_arg, int ntiles, float* __restrict__ dfts, long long d_ep_stride,
                   float* __restrict__ part) {
  constexpr int CW = 16 * MB, c = CW * KW, kW = KW, kT = (KW + 2) * 32;
  const int hw = HW > 0 ? HW : hw_arg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem<MB, KW, NBUF>& sm = *reinterpret_cast<Smem<MB, KW, NBUF>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
  const int e = warp & 3, half = warp >> 2;      // channel class and row range (q in the comments) of this product warp
  const int n = blockIdx.y, b = n / S, si = n - b * S;
  float* dst = dfts + static_cast<long long>(b) * d_ep_stride + static_cast<long long>(si) * c * hw;
  const int tb = static_cast<int>(static_cast<long long>(ntiles) * blockIdx.x / gridDim.x);
  const int te = static_cast<int>(static_cast<long long>(ntiles) * (blockIdx.x + 1) / gridDim.x);
  const bool pixel_warp = warp >= kW;            // the last two warps: pixel step of the foreground / background group
  float* box = sm.tile[0][pixel_warp ? 0 : warp];
  const float* trow = sm.tab + (pixel_warp ? 0 : warp) * CW * kTLd;  // this warp's table rows

  // NBUF = 1: the box is refilled in NH pieces, each as soon as B2 is done with its rows
  constexpr int NH = (NBUF == 1 && MB >= 2) ? 2 : 1;
  if (tid == 0) {
    for (int w = 0; w < NBUF * kW; ++w) mbar_init(&sm.full[0][0] + w, NH);
    mbar_init(&sm.part_bar, kW);
    mbar_init(&sm.wts_bar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // NH bulk-tensor loads per warp and tile: CW / NH rows x 32 floats of class e each, into buffer `buf`
  auto fill = [&](int t, int buf, int h) {
    if (lane == 0) {
      mbar_expect_tx(&sm.full[buf][warp], CW / NH * 32 * 4);
      tma_load_3d(&map, &sm.full[buf][warp], sm.tile[buf][warp] + h * (CW / NH) * 32, (e * hw + t * kStep) & ~3,
                  si * (c / 4) + CW * half + h * (CW / NH), b);
    }
  };
  if (tb < te && !pixel_warp) {
#pragma unroll
    for (int h = 0; h < NH; ++h) fill(tb, 0, h);
  }

  // table of this image, laid out by mpa_bwd_prepare_kernel (train.cu): a straight 16-byte copy (the first tile is on its way meanwhile;
  // when every CTA built its table from coef / ctr itself the prologue was 15 % of all stall samples at ~8 tiles per CTA)
  {
    const float4* tsrc = reinterpret_cast<const float4*>(tabg + static_cast<long long>(n) * c * kTLd);
    float4* tdst = reinterpret_cast<float4*>(sm.tab);
#pragma unroll
    for (int i = 0; i < (c * kTLd / 4 + kT - 1) / kT; ++i)
      if (tid + i * kT < c * kTLd / 4) tdst[tid + i * kT] = __ldg(tsrc + tid + i * kT);
  }
  for (int i = tid; i < 2 * 32 * kWtLd; i += kT) (&sm.wt[0][0])[i] = 0.f;
  for (int i = tid; i < kDvRows * 8; i += kT) sm.dv[i] = 0.f;
  if (tid < kK) {
    sm.konst[tid] = __ldg(beta + n * 2 * kK + kK + tid);
    sm.konst[kK + tid] = __ldg(beta + n * 2 * kK + tid);
  }
  float accB[MB][4];                             // dctr partial: (row 16 mb + g (+8) of the warp, k = 2 tg (+1))
#pragma unroll
  for (int i = 0; i < MB; ++i) accB[i][0] = accB[i][1] = accB[i][2] = accB[i][3] = 0.f;
  __syncthreads();
  float* dstp = part + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * (c + 1) * kK;

  if (pixel_warp) {
    // ============================ pixel-step warps: lane = pixel of the tile ============================
    // wait for the KW partial dot sets of a tile, add them in a fixed order, soft-max and its derivative -> the weights of
    // the tile (double buffered by tile parity: the product warps use them into the next tile) and dv; they own the
    // sums of 2 dl_k.  The product warps never wait at a block barrier: as warps of the same CTA did this step, everybody
    // else sat at two barriers per tile (16 % of all stall samples, ncu).
    const int grp = warp - kW;
    float dsum[kP] = {0.f, 0.f, 0.f};
    for (int t = tb; t < te; ++t) {
      const int x0 = t * kStep, par = (t - tb) & 1;
      const bool live = lane < kStep && x0 + lane < hw;
      const int pl = lane < kStep ? lane : kStep - 1;
      const float m = live ? __ldg((grp == 0 ? fg : bg) + static_cast<long long>(n) * mask_stride + x0 + lane) : 0.f;
      mbar_wait(&sm.part_bar, par);
      float sa[kP], sc[kP];
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        float u0 = 0.f, u1 = 0.f, v0 = 0.f, v1 = 0.f;   // fixed order, two chains (this sum is on the tile's critical path)
#pragma unroll
        for (int w = 0; w < kW; w += 2) {
          u0 += sm.red[w][(grp * kP + k) * kRedLd + pl];
          u1 += sm.red[w + 1][(grp * kP + k) * kRedLd + pl];
          if (k > 0) {
            v0 += sm.red[w][(kK + grp * (kP - 1) + k - 1) * kRedLd + pl];
            v1 += sm.red[w + 1][(kK + grp * (kP - 1) + k - 1) * kRedLd + pl];
          }
        }
        sa[k] = u0 + u1;
        sc[k] = v0 + v1;
      }
      float l[kP], mx = -CUDART_INF_F;
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        l[k] = fmaf(2.0f, sc[k], -sm.konst[grp * kP + k]);
        mx = fmaxf(mx, l[k]);
      }
      float z = 0.f;
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        l[k] = expf(l[k] - mx);
        z += l[k];
      }
      const float iz = 1.0f / z;
      float ds[kP], dot = 0.f, d2[kP];
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        l[k] *= iz;                                                              // sigma_k
        ds[k] = m * (sa[k] + sm.konst[kK + grp * kP + k]);                      // d sigma_k
        dot = fmaf(l[k], ds[k], dot);
      }
      float* wtp = sm.wt[par];
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        d2[k] = live ? 2.0f * l[k] * (ds[k] - dot) : 0.f;
        if (lane < kStep) {
          wtp[lane * kWtLd + grp * kP + k] = m * l[k];
          sm.dv[(lane + 3) * 8 + grp * kP + k] = d2[k];
          if (k > 0) wtp[lane * kWtLd + kK + grp * (kP - 1) + k - 1] = d2[k];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.wts_bar);
#pragma unroll
      for (int k = 0; k < kP; ++k) {               // off the critical path: after the hand-over
        const float tot = warp_sum(d2[k]);
        if (lane == 0) dsum[k] += tot;
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kP; ++k) dstp[c * kK + grp * kP + k] = dsum[k];
    }
    return;
  }
  // ============================ product warps ============================

  // ---- phase B1: df^T [pixel 16] x [row 8] per block, contraction over the 10 columns (8 + 2) ----
  // The columns 0..7 take the three products of the split as three MMAs; the columns 8, 9 would need three more that are
  // three quarters zeros, so their three products are stacked along the contraction axis of ONE MMA instead: slot t < 2 holds
  // (W_hi, T_hi) of column 8 + t, slot t >= 2 (W_lo, T_hi) and slot t + 4 < 6 (W_hi, T_lo) - 4 MMAs per block instead of 6.
  // W fragments of a tile (A operand: pixel 16 mb + g (+8), column tg (+4) for w0; the stacked columns 8 + (tg & 1) for w1)
  struct FragS {
    uint32_t a[4];
  };
  auto load_w = [&](const float* wt, FragA (&w0)[2], FragS (&w1)[2]) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const float* p0 = wt + (mb * 16 + g) * kWtLd + tg;
      const float* p1 = p0 + 8 * kWtLd;
      split_tf32(p0[0], w0[mb].hi[0], w0[mb].lo[0]);
      split_tf32(p1[0], w0[mb].hi[1], w0[mb].lo[1]);
      split_tf32(p0[4], w0[mb].hi[2], w0[mb].lo[2]);
      split_tf32(p1[4], w0[mb].hi[3], w0[mb].lo[3]);
      uint32_t h0, l0, h1, l1;
      split_tf32(p0[8 - tg + (tg & 1)], h0, l0);
      split_tf32(p1[8 - tg + (tg & 1)], h1, l1);
      w1[mb].a[0] = tg < 2 ? h0 : l0;
      w1[mb].a[1] = tg < 2 ? h1 : l1;
      w1[mb].a[2] = tg < 2 ? h0 : 0u;
      w1[mb].a[3] = tg < 2 ? h1 : 0u;
    }
  };
  // One block of 8 table rows (tb8 = its first row) x 32 pixels.  C fragment: (pixel 16 mb + g (+8), rows 2 tg / 2 tg + 1 of the
  // block).  Stored straight from the fragments a warp instruction wrote 4 channel rows x 32 bytes (4-8 L1 requests, and the
  // next block's MMAs waited on the stores' data registers: 14 % of all stall samples, ncu); the block goes through 1 KB of
  // shared memory instead and leaves as 8 rows of 112 contiguous bytes.  orow = row 0 of the block at this lane's pixel; the
  // rows of a block are channels 4 apart.
  float* stg = sm.stage[warp];
  const long long row_step = 4LL * hw;             // the rows of a block are channels 4 apart
  auto b1_block = [&](const FragA (&w0)[2], const FragS (&w1)[2], const float* tb8, float* orow, int rem) {
    const float* tb0 = tb8 + g * kTLd + tg;
    FragB b0;
    split_tf32(tb0[0], b0.hi[0], b0.lo[0]);
    split_tf32(tb0[4], b0.hi[1], b0.lo[1]);
    uint32_t th, tl;
    split_tf32(tb0[8 - tg + (tg & 1)], th, tl);           // stacked columns 8 + (tg & 1): slot tg = T_hi, slot tg + 4 = T_lo (tg < 2)
    tl = tg < 2 ? tl : 0u;
    float d[2][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) d[mb][0] = d[mb][1] = d[mb][2] = d[mb][3] = 0.f;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) mma_tf32(d[mb], w1[mb].a[0], w1[mb].a[1], w1[mb].a[2], w1[mb].a[3], th, tl);
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) mma3(d[mb], w0[mb], b0);
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      float* sp = stg + 2 * tg * kStgLd + mb * 16 + g;
      sp[0] = d[mb][0];
      sp[kStgLd] = d[mb][1];
      sp[8] = d[mb][2];
      sp[kStgLd + 8] = d[mb][3];
    }
    __syncwarp();
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = stg[j * kStgLd + lane];
    __syncwarp();
    if (lane < rem) {
      float* op = orow;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *op = v[j];
        op += row_step;
      }
    }
  };
  // B1 of a tile in two halves of the warp's row blocks: the first one right after the refill (covers the TMA latency), the
  // second one after the dots of the NEXT tile are handed over (covers the pixel step) - hence the double-buffered weights
  auto b1_range = [&](int tp, int nb0, int nb1) {
    FragA w0[2];
    FragS w1[2];
    load_w(sm.wt[(tp - tb) & 1], w0, w1);
    const int xp = tp * kStep, rem = min(kStep, hw - xp);
    float* orow = dst + static_cast<long long>(4 * (CW * half) + e) * hw + xp + lane;
#pragma unroll 2
    for (int nb = nb0; nb < nb1; ++nb) b1_block(w0, w1, trow + nb * 8 * kTLd, orow + 32LL * nb * hw, rem);
  };
  // row blocks in the first half.  Two buffers: the next box is already on its way, so all of B1 goes under the pixel step
  constexpr int kH = NBUF == 2 ? 0 : CW / 16;

  for (int t = tb; t < te; ++t) {
    const int x0 = t * kStep;
    const int o = (e * hw + x0) & 3;              // box column i is pixel x0 + i - o
    const int buf = NBUF == 2 ? (t - tb) & 1 : 0;
    // two buffers: the other one was last read by B2(t-1), so the box of tile t+1 is requested a whole tile ahead
    if (NBUF == 2 && t + 1 < te) fill(t + 1, buf ^ 1, 0);
    box = sm.tile[buf][warp];
    mbar_wait(&sm.full[buf][warp], NBUF == 2 ? ((t - tb) >> 1) & 1 : (t - tb) & 1);
    // ---------------- phase A: dots^T [k 16 (10 used)] x [column 8] per column block, contraction over the warp's rows.
    // Contraction slots tg / tg + 4 of a row block are its rows 2 tg / 2 tg + 1, and column slot n of column block j is box
    // column 4 n + j: a lane's four B elements of a row are then ONE 16-byte chunk (chunk g of the row, conflict-free
    // against the box swizzle) instead of four 4-byte loads with their own swizzled addresses.
    float dacc[4][4];
#pragma unroll
    for (int pb = 0; pb < 4; ++pb) dacc[pb][0] = dacc[pb][1] = dacc[pb][2] = dacc[pb][3] = 0.f;
    {
      const float* fa = box + (2 * tg) * 32 + ((g ^ (2 * tg)) << 2);          // row 2 tg of a block, chunk g
      const float* fb = box + (2 * tg + 1) * 32 + ((g ^ (2 * tg + 1)) << 2);  // row 2 tg + 1
      const float* ta = trow + (2 * tg) * kTLd + g;
#pragma unroll 2
      for (int cb = 0; cb < CW / 8; ++cb) {
        FragA a;
        split_tf32(ta[0], a.hi[0], a.lo[0]);
        split_tf32(g < kNK - 8 ? ta[8] : 0.f, a.hi[1], a.lo[1]);
        split_tf32(ta[kTLd], a.hi[2], a.lo[2]);
        split_tf32(g < kNK - 8 ? ta[kTLd + 8] : 0.f, a.hi[3], a.lo[3]);
        const float4 va = *reinterpret_cast<const float4*>(fa), vb = *reinterpret_cast<const float4*>(fb);
        FragB f[4];
        split_tf32(va.x, f[0].hi[0], f[0].lo[0]);
        split_tf32(va.y, f[1].hi[0], f[1].lo[0]);
        split_tf32(va.z, f[2].hi[0], f[2].lo[0]);
        split_tf32(va.w, f[3].hi[0], f[3].lo[0]);
        split_tf32(vb.x, f[0].hi[1], f[0].lo[1]);
        split_tf32(vb.y, f[1].hi[1], f[1].lo[1]);
        split_tf32(vb.z, f[2].hi[1], f[2].lo[1]);
        split_tf32(vb.w, f[3].hi[1], f[3].lo[1]);
        // the three products of a block as three rounds over the four independent accumulators (no back-to-back dependence)
#pragma unroll
        for (int pb = 0; pb < 4; ++pb) mma_tf32(dacc[pb], a.lo[0], a.lo[1], a.lo[2], a.lo[3], f[pb].hi[0], f[pb].hi[1]);
#pragma unroll
        for (int pb = 0; pb < 4; ++pb) mma_tf32(dacc[pb], a.hi[0], a.hi[1], a.hi[2], a.hi[3], f[pb].lo[0], f[pb].lo[1]);
#pragma unroll
        for (int pb = 0; pb < 4; ++pb) mma_tf32(dacc[pb], a.hi[0], a.hi[1], a.hi[2], a.hi[3], f[pb].hi[0], f[pb].hi[1]);
        fa += 8 * 32;
        fb += 8 * 32;
        ta += 8 * kTLd;
      }
    }
    {
      // C fragment of column block j: (k = g (+8), column slots 2 tg / 2 tg + 1) = box columns 8 tg + j / 8 tg + 4 + j
      float* rw = sm.red[warp];
#pragma unroll
      for (int pb = 0; pb < 4; ++pb) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int p = 8 * tg + 4 * j + pb - o;
          if (p >= 0 && p < kStep) {
            rw[g * kRedLd + p] = dacc[pb][j];
            if (g < kNK - 8) rw[(g + 8) * kRedLd + p] = dacc[pb][2 + j];
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.part_bar);     // dots of tile t handed to the pixel-step warps
    if (t > tb) b1_range(t - 1, kH, CW / 8);      // second half of the previous tile's gradient rows meanwhile
    mbar_wait(&sm.wts_bar, (t - tb) & 1);         // weights / dv of tile t
    // ---------------- phase B2: dctr [row 16] x [k 8 (6 used)] per row block, contraction over the 32 box columns.
    // Contraction slots tg / tg + 4 of column block kb are the box columns 8 tg + 2 kb / 8 tg + 2 kb + 1: a lane's eight A
    // elements of a row (all four column blocks) are the two chunks 2 tg, 2 tg + 1 of the row.
    {
      FragB d[4];
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const float* dp = sm.dv + (8 * tg + 2 * kb - o + 3) * 8 + g;
        split_tf32(dp[0], d[kb].hi[0], d[kb].lo[0]);
        split_tf32(dp[8], d[kb].hi[1], d[kb].lo[1]);
      }
      const float* fr = box + g * 32;                      // row g of a 16-row block; its swizzle phase is g
      const int c0 = ((2 * tg) ^ g) << 2, c1 = ((2 * tg + 1) ^ g) << 2;
      constexpr int MG = MB / NH < 2 ? MB / NH : 2;        // row blocks in flight: independent accumulators, products in rounds
#pragma unroll
      for (int m0 = 0; m0 < MB; m0 += MG) {
        float4 u[MG][4];                                   // [block]{row g: chunks 2 tg, 2 tg + 1; row g + 8: the same}
#pragma unroll
        for (int j = 0; j < MG; ++j) {
          const float* r = fr + (m0 + j) * 16 * 32;
          u[j][0] = *reinterpret_cast<const float4*>(r + c0);
          u[j][1] = *reinterpret_cast<const float4*>(r + c1);
          u[j][2] = *reinterpret_cast<const float4*>(r + 8 * 32 + c0);
          u[j][3] = *reinterpret_cast<const float4*>(r + 8 * 32 + c1);
        }
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          FragA a[MG];
#pragma unroll
          for (int j = 0; j < MG; ++j) {
            // columns 8 tg + 2 kb, + 1: elements (2 kb, 2 kb + 1) of the 8-column run = chunk kb >> 1, halves x y / z w
            const float4 lo4 = u[j][kb >> 1], hi4 = u[j][2 + (kb >> 1)];
            const float e0 = (kb & 1) ? lo4.z : lo4.x, e1 = (kb & 1) ? lo4.w : lo4.y;
            const float e2 = (kb & 1) ? hi4.z : hi4.x, e3 = (kb & 1) ? hi4.w : hi4.y;
            split_tf32(e0, a[j].hi[0], a[j].lo[0]);       // (row g, slot tg)
            split_tf32(e2, a[j].hi[1], a[j].lo[1]);       // (row g + 8, slot tg)
            split_tf32(e1, a[j].hi[2], a[j].lo[2]);       // (row g, slot tg + 4)
            split_tf32(e3, a[j].hi[3], a[j].lo[3]);       // (row g + 8, slot tg + 4)
          }
#pragma unroll
          for (int j = 0; j < MG; ++j)
            mma_tf32(accB[m0 + j], a[j].lo[0], a[j].lo[1], a[j].lo[2], a[j].lo[3], d[kb].hi[0], d[kb].hi[1]);
#pragma unroll
          for (int j = 0; j < MG; ++j)
            mma_tf32(accB[m0 + j], a[j].hi[0], a[j].hi[1], a[j].hi[2], a[j].hi[3], d[kb].lo[0], d[kb].lo[1]);
#pragma unroll
          for (int j = 0; j < MG; ++j)
            mma_tf32(accB[m0 + j], a[j].hi[0], a[j].hi[1], a[j].hi[2], a[j].hi[3], d[kb].hi[0], d[kb].hi[1]);
        }
        // the warp is done with these rows of its half box: fetch them for the next tile (under the rest of B2 and B1)
        if (NH > 1 && m0 + MG == MB / NH) {
          __syncwarp();
          if (t + 1 < te) fill(t + 1, 0, 0);
        }
      }
    }
    __syncwarp();
    if (NBUF == 1 && t + 1 < te) fill(t + 1, 0, NH - 1);
    // ---------------- phase B1 of this tile, first half of the warp's row blocks (under the refill)
    if (kH > 0) b1_range(t, 0, kH);
  }
  if (tb < te) b1_range(te - 1, kH, CW / 8);
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    if (tg < kK / 2) {
      const int ch = 4 * (CW * half + mb * 16 + g) + e;
      *reinterpret_cast<float2*>(dstp + ch * kK + 2 * tg) = make_float2(accB[mb][0], accB[mb][1]);
      *reinterpret_cast<float2*>(dstp + (ch + 32) * kK + 2 * tg) = make_float2(accB[mb][2], accB[mb][3]);
    }
  }
}

// img_part[n] = sum of the image's per-CTA partials, in chunk order: the finalize kernel then reads one partial per image
// instead of one per CTA (880 x 12 KB took it 22 us at 80 images).  A separate launch: letting the last CTA of an image do
// it inside the main kernel needs a __threadfence per CTA, which waits for all of the CTA's gradient stores (9 % of the
// kernel's stall samples, ncu).
__global__ void __launch_bounds__(256)
mpa_bwd_image_sum_kernel(const float* __restrict__ part, int chunks, int M, float* __restrict__ img_part) {
  const int n = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= M) return;
  const float* pp = part + static_cast<long long>(n) * chunks * M + i;
  double sum = 0.0;
#pragma unroll 4
  for (int ch = 0; ch < chunks; ++ch) sum += static_cast<double>(__ldg(pp + static_cast<long long>(ch) * M));
  img_part[static_cast<long long>(n) * M + i] = static_cast<float>(sum);
}

template <int MB, int KW, int NBUF, int HW>
int launch_hw(const CUtensorMap& map, int S, const float* tabg, const float* beta, const float* fg, const float* bg,
              long long mask_stride, int N, int hw, int chunks, float* dfts, long long d_ep, float* part, float* img_part,
              cudaStream_t st) {
  const size_t smem = sizeof(Smem<MB, KW, NBUF>);
  cudaError_t err = cudaFuncSetAttribute(mpa_bwd_mma_kernel<MB, KW, NBUF, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
  if (err != cudaSuccess) return static_cast<int>(err);
  mpa_bwd_mma_kernel<MB, KW, NBUF, HW><<<dim3(chunks, N), (KW + 2) * 32, smem, st>>>(
      map, S, tabg, beta, fg, bg, mask_stride, hw, pemp_mpa_bwd_mma_tiles(hw), dfts, d_ep, part);
  const int M = (16 * MB * KW + 1) * kK;
  mpa_bwd_image_sum_kernel<<<dim3((M + 255) / 256, N), 256, 0, st>>>(part, chunks, M, img_part);
  return PEMP_OK;
}

constexpr int kHwPemp = 51 * 51;                 // the PEMP feature map (417 x 417 crops at stride 8)
template <int MB, int KW, int NBUF>
int launch(const CUtensorMap& map, int S, const float* tabg, const float* beta, const float* fg, const float* bg,
           long long mask_stride, int N, int hw, int chunks, float* dfts, long long d_ep, float* part, float* img_part,
           cudaStream_t st) {
  if (hw == kHwPemp)
    return launch_hw<MB, KW, NBUF, kHwPemp>(map, S, tabg, beta, fg, bg, mask_stride, N, hw, chunks, dfts, d_ep, part, img_part, st);
  return launch_hw<MB, KW, NBUF, 0>(map, S, tabg, beta, fg, bg, mask_stride, N, hw, chunks, dfts, d_ep, part, img_part, st);
}

}  // namespace

bool pemp_mpa_bwd_mma_shape(int c, int p, int hw) { return p == kP && (c == 128 || c == 256 || c == 512 || c == 1024) && hw >= 32; }
int pemp_mpa_bwd_mma_tiles(int hw) { return (hw + kStep - 1) / kStep; }
// rows of a class box per product warp (the table of an image is laid out in that order by mpa_bwd_prepare_kernel)
int pemp_mpa_bwd_mma_rows_per_warp(int c) { return c == 512 ? 32 : c == 1024 ? 128 : 16; }

size_t pemp_mpa_bwd_mma_smem(int c) {
  switch (c) {
    case 128: return sizeof(Smem<1, 8, 1>);
    case 256: return sizeof(Smem<1, 16, 2>);
    case 512: return sizeof(Smem<2, 16, 2>);
    default: return sizeof(Smem<8, 8, 1>);
  }
}

// Returns PEMP_E_ALIGN (nothing launched) when the operand cannot be described by a tensor map; the caller then uses the
// CUDA-core kernel.
int pemp_mpa_bwd_mma_launch(const float* fts, long long ep, int B, int S, const float* ctr, const float* coef, const float* beta,
                            const float* fg, const float* bg, long long mask_stride, int c, int hw, int chunks, float* tabg,
                            float* dfts, long long d_ep, float* part, float* img_part, cudaStream_t st) {
  CUtensorMap map;
  // no L2 promotion beyond the 128-byte row piece: a CTA comes back for the neighbouring piece ~10 us later, by which time the
  // write stream has pushed it out of L2 (with 256-byte promotion the kernel read 1.6 x its algorithmic bytes from DRAM, ncu)
  const int box_rows = c == 1024 ? 64 : pemp_mpa_bwd_mma_rows_per_warp(c);      // c = 1024: two pieces per box
  if (!make_rows4_map(&map, fts, B, S, c, hw, ep, box_rows, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return PEMP_E_ALIGN;
  const int N = B * S;
  switch (c) {
    case 128: return launch<1, 8, 1>(map, S, tabg, beta, fg, bg, mask_stride, N, hw, chunks, dfts, d_ep, part, img_part, st);
    case 256: return launch<1, 16, 2>(map, S, tabg, beta, fg, bg, mask_stride, N, hw, chunks, dfts, d_ep, part, img_part, st);
    case 512: return launch<2, 16, 2>(map, S, tabg, beta, fg, bg, mask_stride, N, hw, chunks, dfts, d_ep, part, img_part, st);
    case 1024: return launch<8, 8, 1>(map, S, tabg, beta, fg, bg, mask_stride, N, hw, chunks, dfts, d_ep, part, img_part, st);
    default: return PEMP_E_SHAPE;
  }
}
// tabg [N][c][kTLd], written by mpa_bwd_prepare_kernel (train.cu) in the row order of the main kernel
size_t pemp_mpa_bwd_mma_table_bytes(int N, int c) { return static_cast<size_t>(N) * c * kTLd * sizeof(float); }
int pemp_mpa_bwd_mma_table_ld() { return kTLd; }
