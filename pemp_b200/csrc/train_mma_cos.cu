// K12: backward of cosine matching (K3) on the warp-level tensor path - the structure of train_mma.cu (TMA-fed,
// warp-private double-buffered quarter boxes, 3 x TF32 `mma.sync` fragments, one CTA per SM with sixteen product warps) applied
// to what autograd records for networks/pemp_stage1.py:214-215, 233-261.  Same math, inputs, outputs and partial layout as
// `cosine_bwd_kernel` (train.cu); P = 3, c in {256, 512}, hw >= 32, TMA-encodable query maps; everything else takes that kernel.
//
// Per pixel tile, with the 6-column table T[ch] = normalised prototypes (k < 3 background, k >= 3 foreground):
//   phase A   s[k][x]   = sum_ch T[ch][k] q[ch][x],  |q|^2[x] = sum_ch q[ch][x]^2     (the norm on the CUDA cores, from the same
//                                                                                       fragment registers)
//   pixel step           arg-max per class group (first maximum wins, recomputed) or, DENSE, every column:
//                        W[x][k] = g[x] scalar / |q|  for the selected columns, t[x] = sum_k W[x][k] s[k][x] / |q|^2
//   phase B1  dq[x][ch] = sum_k W[x][k] T[ch][k] - t[x] q[ch][x]
//   phase B2  dT[ch][k] = sum_x q[ch][x] W[x][k]
// Differences to the K2 kernel: one 8-wide column block (no stacked columns); ONE pixel-step warp (the two class groups share
// |q| and t); B1 needs the tile itself, so the deferred B1(t-1) reads the box of tile t-1 and the box of tile t+1 is requested
// only after it (a shorter prefetch distance than K2's whole tile: B1(t-1) hides the pixel step, weights wait + B2 hide the load).
#include <math_constants.h>

#include "tma_common.cuh"

bool pemp_cos_bwd_mma_shape(int c, int P, int hw);
size_t pemp_cos_bwd_mma_smem(int c);
size_t pemp_cos_bwd_mma_table_bytes(int Bp, int c);
int pemp_cos_bwd_mma_tiles(int hw);
int pemp_cos_bwd_mma_launch(bool dense, const float* qry, long long ep, int Bp, int Q, const float* pn, const float* g, int c, int hw,
                            int chunks, float scalar, float* tabg, float* dq, long long d_ep, float* part, cudaStream_t st);

namespace {

using namespace pemp_tma;

constexpr int kP = 3, kK = 2 * kP;               // prototypes per group, table columns
constexpr int kKW = 16;                          // product warps (+ one pixel-step warp)
constexpr int kTLd = 12;                         // table row pitch (conflict-free fragments, see tests/test_fragment_maps.py)
constexpr int kStep = 28;                        // pixels a tile advances
constexpr int kRedRows = kK + 1;                 // 6 dots + |q|^2
constexpr int kRedLd = 33;
constexpr int kWtLd = 12;                        // W[x][0..6) + zeros
constexpr int kStgLd = 36;
constexpr int kDvRows = 32 + 3;
constexpr float kCosEps = 1e-8f;                 // F.cosine_similarity's eps

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
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

template <int MB>                                // 16-row blocks per product warp: c = 256 MB
struct Smem {
  static constexpr int CW = 16 * MB, c = CW * kKW;
  alignas(1024) float tile[2][kKW][CW * 32];     // warp w: rows [CW q, CW q + CW) of the class-e box, e = w & 3, q = w >> 2
  alignas(16) float tab[c * kTLd];               // row R = w CW + r  <->  channel 4 (CW q + r) + e
  alignas(16) float red[kKW][kRedRows * kRedLd]; // partial dots and squared norms of the warps: [k][pixel of the tile]
  alignas(16) float wt[2][32 * kWtLd];           // [tile parity][pixel]{ W (6) | 0 .. }
  alignas(16) float dv[kDvRows * 8];             // [pixel + 3]{ W (6) | 0 0 }
  alignas(16) float tq[2][32];                   // [tile parity][pixel] t
  alignas(16) float stage[kKW][8 * kStgLd];
  alignas(8) uint64_t full[2][kKW];
  alignas(8) uint64_t part_bar;
  alignas(8) uint64_t wts_bar;
};
static_assert(sizeof(Smem<2>) <= 227 * 1024, "c = 512, double buffered, one CTA per SM");

template <int MB, int HW, bool DENSE>
__global__ void __launch_bounds__((kKW + 1) * 32, 1)
cos_bwd_mma_kernel(const __grid_constant__ CUtensorMap map, int Q, const float* __restrict__ tabg, const float* __restrict__ gin,
                   float scalar, int hw_arg, int ntiles, float* __restrict__ dq, long long d_ep_stride, float* __restrict__ part) {
  constexpr int CW = 16 * MB, c = CW * kKW, kT = (kKW + 1) * 32;
  const int hw = HW > 0 ? HW : hw_arg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem<MB>& sm = *reinterpret_cast<Smem<MB>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
  const int e = warp & 3, qr = warp >> 2;        // channel class and row range of this product warp
  const int n = blockIdx.y, b = n / Q, qi = n - b * Q;
  float* dst = dq + static_cast<long long>(b) * d_ep_stride + static_cast<long long>(qi) * c * hw;
  const int tb = static_cast<int>(static_cast<long long>(ntiles) * blockIdx.x / gridDim.x);
  const int te = static_cast<int>(static_cast<long long>(ntiles) * (blockIdx.x + 1) / gridDim.x);
  const bool pixel_warp = warp >= kKW;
  const float* trow = sm.tab + (pixel_warp ? 0 : warp) * CW * kTLd;

  if (tid == 0) {
    for (int w = 0; w < 2 * kKW; ++w) mbar_init(&sm.full[0][0] + w, 1);
    mbar_init(&sm.part_bar, kKW);
    mbar_init(&sm.wts_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto fill = [&](int t, int buf) {
    if (lane == 0) {
      mbar_expect_tx(&sm.full[buf][warp], CW * 32 * 4);
      tma_load_3d(&map, &sm.full[buf][warp], sm.tile[buf][warp], (e * hw + t * kStep) & ~3, qi * (c / 4) + CW * qr, b);
    }
  };
  if (tb < te && !pixel_warp) fill(tb, 0);

  {  // table of this prototype set, laid out by cos_bwd_table_kernel: a straight 16-byte copy
    const float4* tsrc = reinterpret_cast<const float4*>(tabg + static_cast<long long>(b) * c * kTLd);
    float4* tdst = reinterpret_cast<float4*>(sm.tab);
    for (int i = tid; i < c * kTLd / 4; i += kT) tdst[i] = __ldg(tsrc + i);
  }
  for (int i = tid; i < 2 * 32 * kWtLd; i += kT) (&sm.wt[0][0])[i] = 0.f;
  for (int i = tid; i < kDvRows * 8; i += kT) sm.dv[i] = 0.f;
  for (int i = tid; i < 2 * 32; i += kT) (&sm.tq[0][0])[i] = 0.f;
  __syncthreads();

  if (pixel_warp) {
    // ============================ pixel-step warp: lane = pixel of the tile ============================
    for (int t = tb; t < te; ++t) {
      const int x0 = t * kStep, par = (t - tb) & 1, x = x0 + lane;
      const bool live = lane < kStep && x < hw;
      const int pl = lane < kStep ? lane : kStep - 1;
      float gk[DENSE ? kK : 2];                  // upstream gradient of this pixel, times scalar
#pragma unroll
      for (int k = 0; k < (DENSE ? kK : 2); ++k)
        gk[k] = live ? __ldg(gin + (static_cast<long long>(n) * (DENSE ? kK : 2) + k) * hw + x) * scalar : 0.f;
      mbar_wait(&sm.part_bar, par);
      float s[kRedRows];
#pragma unroll
      for (int k = 0; k < kRedRows; ++k) {
        float u0 = 0.f, u1 = 0.f;                // fixed order, two chains
#pragma unroll
        for (int w = 0; w < kKW; w += 2) {
          u0 += sm.red[w][k * kRedLd + pl];
          u1 += sm.red[w + 1][k * kRedLd + pl];
        }
        s[k] = u0 + u1;
      }
      const float nq = sqrtf(s[kK]);
      const float invq = 1.0f / fmaxf(nq, kCosEps);
      float w6[kK], tsum = 0.f;
      if (DENSE) {
#pragma unroll
        for (int k = 0; k < kK; ++k) {
          w6[k] = gk[k] * invq;
          tsum = fmaf(gk[k], s[k], tsum);
        }
      } else {
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
          int best = 0;
          float bv = s[grp * kP];
#pragma unroll
          for (int k = 1; k < kP; ++k) {
            if (s[grp * kP + k] > bv) {
              bv = s[grp * kP + k];
              best = k;
            }
          }
          const float co = gk[grp] * invq;
          tsum = fmaf(gk[grp], bv, tsum);
#pragma unroll
          for (int k = 0; k < kP; ++k) w6[grp * kP + k] = (k == best) ? co : 0.f;
        }
      }
      const float tv = (nq < kCosEps) ? 0.f : tsum * invq * invq * invq;
      if (lane < kStep) {
        float* wtp = sm.wt[par];
#pragma unroll
        for (int k = 0; k < kK; ++k) {
          wtp[lane * kWtLd + k] = w6[k];
          sm.dv[(lane + 3) * 8 + k] = w6[k];
        }
        sm.tq[par][lane] = tv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.wts_bar);
    }
    return;
  }
  // ============================ product warps ============================
  float accB[MB][4];                             // dT partial: (row 16 mb + g (+8) of the warp, k = 2 tg (+1))
#pragma unroll
  for (int i = 0; i < MB; ++i) accB[i][0] = accB[i][1] = accB[i][2] = accB[i][3] = 0.f;
  float* stg = sm.stage[warp];
  const long long row_step = 4LL * hw;           // the rows of a block are channels 4 apart
  const int o = (e * hw) & 3;                    // box column i is pixel x0 + i - o (x0 is a multiple of 4: o does not depend on the tile)

  // B1 of tile tp from its weights, its t and ITS box: df^T [pixel 16] x [row 8] per block, one 8-wide column block
  auto b1_tile = [&](int tp, const float* bx) {
    const int par = (tp - tb) & 1;
    FragA w0[2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const float* p0 = sm.wt[par] + (mb * 16 + g) * kWtLd + tg;
      const float* p1 = p0 + 8 * kWtLd;
      split_tf32(p0[0], w0[mb].hi[0], w0[mb].lo[0]);
      split_tf32(p1[0], w0[mb].hi[1], w0[mb].lo[1]);
      split_tf32(p0[4], w0[mb].hi[2], w0[mb].lo[2]);
      split_tf32(p1[4], w0[mb].hi[3], w0[mb].lo[3]);
    }
    const int xp = tp * kStep, rem = min(kStep, hw - xp);
    const float tl = sm.tq[par][lane];           // t of this lane's pixel (0 for lanes >= 28: zero-initialised)
    const int col = lane + o;                    // box column of this lane's pixel (<= 30 for lanes < 28)
    float* orow = dst + static_cast<long long>(4 * (CW * qr) + e) * hw + xp + lane;
#pragma unroll 2
    for (int nb = 0; nb < CW / 8; ++nb) {
      const float* tb0 = trow + (nb * 8 + g) * kTLd + tg;
      FragB b0;
      split_tf32(tb0[0], b0.hi[0], b0.lo[0]);
      split_tf32(tb0[4], b0.hi[1], b0.lo[1]);
      float d[2][4];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) d[mb][0] = d[mb][1] = d[mb][2] = d[mb][3] = 0.f;
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
      for (int j = 0; j < 8; ++j) {
        const int row = nb * 8 + j;              // - t q: the feature of (row, this lane's pixel) from the swizzled box
        const float qv = lane < kStep ? bx[row * 32 + ((((col >> 2) ^ row) & 7) << 2) + (col & 3)] : 0.f;
        v[j] = fmaf(-tl, qv, stg[j * kStgLd + lane]);
      }
      __syncwarp();
      if (lane < rem) {
        float* op = orow;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *op = v[j];
          op += row_step;
        }
      }
      orow += 8 * row_step;
    }
  };

  for (int t = tb; t < te; ++t) {
    const int buf = (t - tb) & 1;
    const float* box = sm.tile[buf][warp];
    mbar_wait(&sm.full[buf][warp], ((t - tb) >> 1) & 1);
    // ---------------- phase A: dots^T [k 16 (6 used)] x [column 8] per column block + squared norms of the columns
    float dacc[4][4], nacc[4];
#pragma unroll
    for (int pb = 0; pb < 4; ++pb) {
      dacc[pb][0] = dacc[pb][1] = dacc[pb][2] = dacc[pb][3] = 0.f;
      nacc[pb] = 0.f;
    }
    {
      const float* fa = box + (2 * tg) * 32 + ((g ^ (2 * tg)) << 2);          // row 2 tg of a block, chunk g
      const float* fb = box + (2 * tg + 1) * 32 + ((g ^ (2 * tg + 1)) << 2);  // row 2 tg + 1
      const float* ta = trow + (2 * tg) * kTLd + g;
#pragma unroll 2
      for (int cb = 0; cb < CW / 8; ++cb) {
        FragA a;
        split_tf32(ta[0], a.hi[0], a.lo[0]);                 // columns 6, 7 of the table hold zeros
        split_tf32(ta[kTLd], a.hi[2], a.lo[2]);
        a.hi[1] = a.lo[1] = a.hi[3] = a.lo[3] = 0u;           // rows 8..15 of the k axis: unused
        const float4 va = *reinterpret_cast<const float4*>(fa), vb = *reinterpret_cast<const float4*>(fb);
        nacc[0] = fmaf(va.x, va.x, fmaf(vb.x, vb.x, nacc[0]));
        nacc[1] = fmaf(va.y, va.y, fmaf(vb.y, vb.y, nacc[1]));
        nacc[2] = fmaf(va.z, va.z, fmaf(vb.z, vb.z, nacc[2]));
        nacc[3] = fmaf(va.w, va.w, fmaf(vb.w, vb.w, nacc[3]));
        FragB f[4];
        split_tf32(va.x, f[0].hi[0], f[0].lo[0]);
        split_tf32(va.y, f[1].hi[0], f[1].lo[0]);
        split_tf32(va.z, f[2].hi[0], f[2].lo[0]);
        split_tf32(va.w, f[3].hi[0], f[3].lo[0]);
        split_tf32(vb.x, f[0].hi[1], f[0].lo[1]);
        split_tf32(vb.y, f[1].hi[1], f[1].lo[1]);
        split_tf32(vb.z, f[2].hi[1], f[2].lo[1]);
        split_tf32(vb.w, f[3].hi[1], f[3].lo[1]);
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
      // C fragment of column block j: (k = g, column slots 2 tg / 2 tg + 1) = box columns 8 tg + j / 8 tg + 4 + j
      float* rw = sm.red[warp];
#pragma unroll
      for (int pb = 0; pb < 4; ++pb) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int p = 8 * tg + 4 * j + pb - o;
          if (g < kK && p >= 0 && p < kStep) rw[g * kRedLd + p] = dacc[pb][j];
        }
        // squared norm of box column 4 g + pb: this lane holds the rows {2 tg, 2 tg + 1} of every block
        float nn = nacc[pb];
        nn += __shfl_xor_sync(kFull, nn, 1);
        nn += __shfl_xor_sync(kFull, nn, 2);
        const int pn = 4 * g + pb - o;
        if (tg == 0 && pn >= 0 && pn < kStep) rw[kK * kRedLd + pn] = nn;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.part_bar);     // dots of tile t handed to the pixel-step warp
    if (t > tb) b1_tile(t - 1, sm.tile[buf ^ 1][warp]);   // the previous tile's gradient rows meanwhile (its box is still there)
    __syncwarp();
    if (t + 1 < te) fill(t + 1, buf ^ 1);         // only now may the other buffer be refilled
    mbar_wait(&sm.wts_bar, (t - tb) & 1);         // weights of tile t
    // ---------------- phase B2: dT [row 16] x [k 8 (6 used)] per row block, contraction over the 32 box columns
    {
      FragB d[4];
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const float* dp = sm.dv + (8 * tg + 2 * kb - o + 3) * 8 + g;
        split_tf32(dp[0], d[kb].hi[0], d[kb].lo[0]);
        split_tf32(dp[8], d[kb].hi[1], d[kb].lo[1]);
      }
      const float* fr = box + g * 32;
      const int c0 = ((2 * tg) ^ g) << 2, c1 = ((2 * tg + 1) ^ g) << 2;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const float* r = fr + mb * 16 * 32;
        float4 u[4];
        u[0] = *reinterpret_cast<const float4*>(r + c0);
        u[1] = *reinterpret_cast<const float4*>(r + c1);
        u[2] = *reinterpret_cast<const float4*>(r + 8 * 32 + c0);
        u[3] = *reinterpret_cast<const float4*>(r + 8 * 32 + c1);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const float4 lo4 = u[kb >> 1], hi4 = u[2 + (kb >> 1)];
          const float e0 = (kb & 1) ? lo4.z : lo4.x, e1 = (kb & 1) ? lo4.w : lo4.y;
          const float e2 = (kb & 1) ? hi4.z : hi4.x, e3 = (kb & 1) ? hi4.w : hi4.y;
          FragA a;
          split_tf32(e0, a.hi[0], a.lo[0]);
          split_tf32(e2, a.hi[1], a.lo[1]);
          split_tf32(e1, a.hi[2], a.lo[2]);
          split_tf32(e3, a.hi[3], a.lo[3]);
          mma3(accB[mb], a, d[kb]);
        }
      }
    }
  }
  if (tb < te) b1_tile(te - 1, sm.tile[(te - 1 - tb) & 1][warp]);
  float* dstp = part + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * c * kK;
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    if (tg < kK / 2) {
      const int ch = 4 * (CW * qr + mb * 16 + g) + e;
      *reinterpret_cast<float2*>(dstp + ch * kK + 2 * tg) = make_float2(accB[mb][0], accB[mb][1]);
      *reinterpret_cast<float2*>(dstp + (ch + 32) * kK + 2 * tg) = make_float2(accB[mb][2], accB[mb][3]);
    }
  }
}

// tabg [Bp][c][kTLd] in the row order of the main kernel: row R = w CW + r <-> channel 4 (CW (w >> 2) + r) + (w & 3)
__global__ void __launch_bounds__(256)
cos_bwd_table_kernel(const float* __restrict__ pn, int c, int CW, float* __restrict__ tabg) {
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= c * kTLd) return;
  const int R = i / kTLd, k = i - R * kTLd;
  const int w = R / CW, r = R - w * CW, ch = 4 * (CW * (w >> 2) + r) + (w & 3);
  tabg[static_cast<long long>(b) * c * kTLd + i] = k < kK ? __ldg(pn + (static_cast<long long>(b) * c + ch) * kK + k) : 0.f;
}

constexpr int kHwPemp = 51 * 51;
template <int MB, int HW>
int launch_hw(bool dense, const CUtensorMap& map, int Q, const float* tabg, const float* g, float scalar, int N, int hw, int chunks,
              float* dq, long long d_ep, float* part, cudaStream_t st) {
  const size_t smem = sizeof(Smem<MB>);
  const int nt = pemp_cos_bwd_mma_tiles(hw);
  cudaError_t err;
  if (dense) {
    err = cudaFuncSetAttribute(cos_bwd_mma_kernel<MB, HW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return static_cast<int>(err);
    cos_bwd_mma_kernel<MB, HW, true><<<dim3(chunks, N), (kKW + 1) * 32, smem, st>>>(map, Q, tabg, g, scalar, hw, nt, dq, d_ep, part);
  } else {
    err = cudaFuncSetAttribute(cos_bwd_mma_kernel<MB, HW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return static_cast<int>(err);
    cos_bwd_mma_kernel<MB, HW, false><<<dim3(chunks, N), (kKW + 1) * 32, smem, st>>>(map, Q, tabg, g, scalar, hw, nt, dq, d_ep, part);
  }
  return PEMP_OK;
}
template <int MB>
int launch(bool dense, const CUtensorMap& map, int Q, const float* tabg, const float* g, float scalar, int N, int hw, int chunks,
           float* dq, long long d_ep, float* part, cudaStream_t st) {
  if (hw == kHwPemp) return launch_hw<MB, kHwPemp>(dense, map, Q, tabg, g, scalar, N, hw, chunks, dq, d_ep, part, st);
  return launch_hw<MB, 0>(dense, map, Q, tabg, g, scalar, N, hw, chunks, dq, d_ep, part, st);
}

}  // namespace

bool pemp_cos_bwd_mma_shape(int c, int P, int hw) { return P == kP && (c == 256 || c == 512) && hw >= 32; }
size_t pemp_cos_bwd_mma_smem(int c) { return c == 512 ? sizeof(Smem<2>) : sizeof(Smem<1>); }
size_t pemp_cos_bwd_mma_table_bytes(int Bp, int c) { return static_cast<size_t>(Bp) * c * kTLd * sizeof(float); }
int pemp_cos_bwd_mma_tiles(int hw) { return (hw + kStep - 1) / kStep; }

// Returns PEMP_E_ALIGN (nothing launched) when the query maps cannot be described by a tensor map; the caller then uses the
// CUDA-core kernel.  pn [Bp][c][6] normalised prototypes (proto_norm_kernel, train.cu); part [N][chunks][c][6].
int pemp_cos_bwd_mma_launch(bool dense, const float* qry, long long ep, int Bp, int Q, const float* pn, const float* g, int c, int hw,
                            int chunks, float scalar, float* tabg, float* dq, long long d_ep, float* part, cudaStream_t st) {
  CUtensorMap map;
  const int CW = c / kKW;
  if (!make_rows4_map(&map, qry, Bp, Q, c, hw, ep, CW, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return PEMP_E_ALIGN;
  cos_bwd_table_kernel<<<dim3((c * kTLd + 255) / 256, Bp), 256, 0, st>>>(pn, c, CW, tabg);
  const int N = Bp * Q;
  if (c == 512) return launch<2>(dense, map, Q, tabg, g, scalar, N, hw, chunks, dq, d_ep, part, st);
  return launch<1>(dense, map, Q, tabg, g, scalar, N, hw, chunks, dq, d_ep, part, st);
}
