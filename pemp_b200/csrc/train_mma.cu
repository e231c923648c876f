// K12: backward of meta-prototype attention with the three thin products of a tile on the warp-level tensor
// path (`mma.sync.m16n8k8` tf32, SASS HMMA.1688), 3 x TF32 split = fp32-grade.  Same math, inputs, outputs and partial layout
// as `mpa_bwd_kernel` (train.cu; reference: what autograd records for networks/pemp_stage1.py:202-213), P = 3, c % 128 == 0.
//
// Why.  Per 32-pixel tile the backward is three products with a 10-column table  T[ch] = { coef[ch][0..6) | ctr_k - ctr_g0 }:
//   phase A   dots[k][x]  = sum_ch T[ch][k] f[ch][x]            (10 x 32, contraction over the c channels)
//   phase B1  df[x][ch]   = sum_k  W[x][k]  T[ch][k]            (32 x c,  contraction over the 10 columns; W from the pixel step)
//   phase B2  dctr[ch][k] = sum_x  f[ch][x] dv[x][k]            (c x 6,   contraction over the 32 pixels)
// On CUDA cores every FMA of A and B1 needs a table value that is uniform over the warp (lane = pixel), and a broadcast
// LDS.128 still costs four cycles of the 128 B/clk shared-memory crossbar: 12 cycles per channel and warp in A and again in B1
// = ~14 800 crossbar cycles per tile against the ~5 300 clocks the tile's 128 KB take at the HBM rate - the round-2 kernel sat
// at 0.31-0.33 of the HBM peak with the shared pipe 63 % busy.  With register fragments a table / weight value is read once per
// 8 x 8 block: ~2 400 crossbar cycles and ~1 000 issue slots per warp and tile (2 800 before).  tcgen05 is not an option here:
// its operands live in shared memory behind descriptors (T as hi + lo K-major panels = 48-64 KB next to the 64-KB tile, twice
// per SM) and the three products contract over three different axes of the same tile; 240 m16n8k8 per warp and tile keep the
// legacy pipe (measured 2.0 clk per instruction and SM, tools/probes/mma_sync_probe.cu) ~45 % busy at 0.6 of the HBM peak.
//
// 3 x TF32: x = hi + lo with hi = x & 0xffffe000 (what the tensor core reads of an fp32 register) and lo = x - hi (exact);
// a.b ~ lo_a hi_b + hi_a lo_b + hi_a hi_b, small products first (see K9), error ~2^-21 per product.
//
// Structure: 256 threads, two CTAs per SM; a CTA owns a run of 32-pixel tiles of one image.  Warp w owns the channels
// [w c/8, (w+1) c/8) in every phase, so the tile buffer is warp-private: it is filled by 4-byte `cp.async` (rows start at
// arbitrary 4-byte phases: hw is odd) one phase ahead - the loads of tile t+1 are issued after B2(t), run under B1(t), which
// needs no features - and only two block barriers per tile remain (dots -> pixel step -> weights).
//   tile [ch][32] with column x stored at x ^ s(ch), s = (ch & 3) << 3 | ((ch >> 2) & 1) << 2: conflict-free for the
//   row-contiguous fill, the B fragments of phase A (4 channels x 8 pixels) and the A fragments of B2 (8 channels x 4 pixels).
#include <math_constants.h>

#include "common.cuh"

size_t pemp_mpa_bwd_mma_smem(int c);
bool pemp_mpa_bwd_mma_shape(int c, int p);
int pemp_mpa_bwd_mma_launch(const float* fts, long long ep, int S, const float* ctr, const float* coef, const float* beta,
                            const float* fg, const float* bg, long long mask_stride, int N, int c, int hw, int chunks, int ntiles,
                            float* dfts, long long d_ep, float* part, cudaStream_t st);

namespace {

constexpr int kT = 256, kW = 8;                  // threads / warps per CTA
constexpr int kP = 3, kK = 2 * kP;               // prototypes per group, coefficient columns
constexpr int kND = 2 * (kP - 1);                // centre-difference columns (the first prototype of a group has none)
constexpr int kNK = kK + kND;                    // 10 table columns: [0, 6) coef, [6, 8) fg differences, [8, 10) bg differences
constexpr int kN1 = kNK - 8;                     // columns of the second 8-wide block
constexpr int kRedLd = 40;                       // pixel pitch of a dot row in `red` (conflict-free 64-bit fragment stores)
constexpr int kWtLd = 12;                        // W[x][0..10) + two zero columns (conflict-free fragment reads)
static_assert(kN1 == 2, "the second column block is laid out for two columns");

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
// C 16 x 8 (c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)) with g = lane >> 2, t = lane & 3
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src, uint32_t nbytes) {   // nbytes 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}

template <int MB>                                // 16-channel blocks per warp: c = 128 MB
__global__ void __launch_bounds__(kT, MB <= 4 ? 2 : 1)
mpa_bwd_mma_kernel(const float* __restrict__ fts, long long ep_stride, int S, const float* __restrict__ ctr,
                   const float* __restrict__ coef, const float* __restrict__ beta, const float* __restrict__ fg,
                   const float* __restrict__ bg, long long mask_stride, int hw, int ntiles, float* __restrict__ dfts,
                   long long d_ep_stride, float* __restrict__ part) {
  constexpr int c = 128 * MB, CW = 16 * MB;      // channels; channels per warp
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                              // [c][32] swizzled
  float* t0 = tile + c * 32;                     // [c][8]   table columns 0..7, column k stored at k ^ (((ch >> 2) & 1) << 2)
  float* t1 = t0 + c * 8;                        // [c][2]   table columns 8, 9
  float* red = t1 + c * kN1;                     // [kW][kNK][kRedLd] partial dots of the warps
  float* wt = red + kW * kNK * kRedLd;           // [32][kWtLd] { a_k (6) | 2 dl_k of the non-first prototypes (4) | 0 0 }
  float* dv = wt + 32 * kWtLd;                   // [32][8]  { 2 dl_k (6) | 0 0 }
  float* konst = dv + 32 * 8;                    // [kK] |ctr_k|^2 - |ctr_g0|^2, [kK] beta
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
  const int n = blockIdx.y, b = n / S, si = n - b * S;
  const int wch = warp * CW;                     // this warp's channels [wch, wch + CW)
  const float* src = fts + static_cast<long long>(b) * ep_stride + static_cast<long long>(si) * c * hw;
  float* dst = dfts + static_cast<long long>(b) * d_ep_stride + static_cast<long long>(si) * c * hw;
  const int tb = static_cast<int>(static_cast<long long>(ntiles) * blockIdx.x / gridDim.x);
  const int te = static_cast<int>(static_cast<long long>(ntiles) * (blockIdx.x + 1) / gridDim.x);

  // fill of the warp's rows of one tile: lane = pixel, 128 contiguous bytes per row, zero fill past the end of the row
  const uint32_t tile_s = smem_u32(tile);
  auto fill = [&](int t) {
    const int x = t * 32 + lane;
    const uint32_t nb = x < hw ? 4u : 0u;
    const float* gp = src + static_cast<long long>(wch) * hw + (x < hw ? x : hw - 1);
#pragma unroll 8
    for (int r = 0; r < CW; ++r) {               // CW is a multiple of 16, wch of 16: s(ch) only depends on r & 7
      const int sw = ((r & 3) << 3) | (((r >> 2) & 1) << 2);
      cp_async4(tile_s + 4u * static_cast<uint32_t>((wch + r) * 32 + (lane ^ sw)), gp, nb);
      gp += hw;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (tb < te) fill(tb);

  // tables of this image (the first tile is on its way meanwhile)
  for (int i = tid; i < c * kNK; i += kT) {
    const int ch = i / kNK, k = i - ch * kNK;
    float v;
    if (k < kK) {
      v = __ldg(coef + (static_cast<long long>(n) * c + ch) * kK + k);
    } else {
      // centre columns as differences to the first prototype of their group (exact in double, one rounding), see train.cu
      const int j = k - kK, grp = j / (kP - 1), kk = grp * kP + 1 + (j - grp * (kP - 1));
      v = static_cast<float>(static_cast<double>(__ldg(ctr + ch * kK + kk)) - static_cast<double>(__ldg(ctr + ch * kK + grp * kP)));
    }
    if (k < 8)
      t0[ch * 8 + (k ^ (((ch >> 2) & 1) << 2))] = v;
    else
      t1[ch * kN1 + (k - 8)] = v;
  }
  for (int i = tid; i < 32 * kWtLd; i += kT) wt[i] = 0.f;
  for (int i = tid; i < 32 * 8; i += kT) dv[i] = 0.f;
  if (tid < kK) {
    konst[tid] = __ldg(beta + n * 2 * kK + kK + tid);
    konst[kK + tid] = __ldg(beta + n * 2 * kK + tid);
  }
  float accB[MB][4];                             // dctr partial: (ch = wch + 16 mb + g (+8), k = 2 tg (+1))
#pragma unroll
  for (int i = 0; i < MB; ++i) accB[i][0] = accB[i][1] = accB[i][2] = accB[i][3] = 0.f;
  float dsum[kP] = {0.f, 0.f, 0.f};              // lane 0 of warps 0 / 1: sum_x 2 dl_k of its group
  __syncthreads();

  for (int t = tb; t < te; ++t) {
    const int x0 = t * 32;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---------------- phase A: dots^T [k 16 (10 used)] x [px 8] per pixel block, contraction over the warp's channels
    float dacc[4][4];
#pragma unroll
    for (int pb = 0; pb < 4; ++pb) dacc[pb][0] = dacc[pb][1] = dacc[pb][2] = dacc[pb][3] = 0.f;
#pragma unroll 2
    for (int cb = 0; cb < CW / 8; ++cb) {
      const int ch0 = wch + cb * 8 + tg;         // channels of a0 / b0; a2 / b1: + 4 (bit 2 set: swizzles flip)
      FragA a;
      split_tf32(t0[ch0 * 8 + g], a.hi[0], a.lo[0]);
      split_tf32(t0[(ch0 + 4) * 8 + (g ^ 4)], a.hi[2], a.lo[2]);
      const float u1 = g < kN1 ? t1[ch0 * kN1 + g] : 0.f, u3 = g < kN1 ? t1[(ch0 + 4) * kN1 + g] : 0.f;
      split_tf32(u1, a.hi[1], a.lo[1]);
      split_tf32(u3, a.hi[3], a.lo[3]);
      const float* r0 = tile + ch0 * 32;
      const float* r1 = tile + (ch0 + 4) * 32;
#pragma unroll
      for (int pb = 0; pb < 4; ++pb) {
        const int px = pb * 8 + g;
        FragB f;
        split_tf32(r0[px ^ (tg << 3)], f.hi[0], f.lo[0]);
        split_tf32(r1[px ^ ((tg << 3) | 4)], f.hi[1], f.lo[1]);
        mma3(dacc[pb], a, f);
      }
    }
    {
      float* rw = red + warp * (kNK * kRedLd);
#pragma unroll
      for (int pb = 0; pb < 4; ++pb) {
        *reinterpret_cast<float2*>(rw + g * kRedLd + pb * 8 + 2 * tg) = make_float2(dacc[pb][0], dacc[pb][1]);
        if (g < kN1) *reinterpret_cast<float2*>(rw + (g + 8) * kRedLd + pb * 8 + 2 * tg) = make_float2(dacc[pb][2], dacc[pb][3]);
      }
    }
    __syncthreads();
    // ---------------- pixel step: warp 0 = foreground group, warp 1 = background group, lane = pixel
    if (warp < 2) {
      const int grp = warp, x = x0 + lane;
      const float m = x < hw ? __ldg((grp == 0 ? fg : bg) + static_cast<long long>(n) * mask_stride + x) : 0.f;
      float sa[kP], sc[kP];
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        float u = 0.f, v = 0.f;
#pragma unroll
        for (int w = 0; w < kW; ++w) {            // fixed order
          u += red[(w * kNK + grp * kP + k) * kRedLd + lane];
          if (k > 0) v += red[(w * kNK + kK + grp * (kP - 1) + k - 1) * kRedLd + lane];
        }
        sa[k] = u;
        sc[k] = v;
      }
      float l[kP], mx = -CUDART_INF_F;
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        l[k] = fmaf(2.0f, sc[k], -konst[grp * kP + k]);
        mx = fmaxf(mx, l[k]);
      }
      float z = 0.f;
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        l[k] = expf(l[k] - mx);
        z += l[k];
      }
      const float iz = 1.0f / z;
      float ds[kP], dot = 0.f;
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        l[k] *= iz;                                                              // sigma_k
        ds[k] = m * (sa[k] + konst[kK + grp * kP + k]);                         // d sigma_k
        dot = fmaf(l[k], ds[k], dot);
      }
#pragma unroll
      for (int k = 0; k < kP; ++k) {
        const float d2 = 2.0f * l[k] * (ds[k] - dot);
        wt[lane * kWtLd + grp * kP + k] = m * l[k];
        dv[lane * 8 + grp * kP + k] = d2;
        if (k > 0) wt[lane * kWtLd + kK + grp * (kP - 1) + k - 1] = d2;
        const float tot = warp_sum(d2);
        if (lane == 0) dsum[k] += tot;
      }
    }
    __syncthreads();
    // ---------------- phase B2: dctr [ch 16] x [k 8 (6 used)] per channel block, contraction over the 32 pixels
    {
      FragB d[4];
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        split_tf32(dv[(kb * 8 + tg) * 8 + g], d[kb].hi[0], d[kb].lo[0]);
        split_tf32(dv[(kb * 8 + tg + 4) * 8 + g], d[kb].hi[1], d[kb].lo[1]);
      }
      const int sw = ((g & 3) << 3) | (((g >> 2) & 1) << 2);
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const float* r0 = tile + (wch + mb * 16 + g) * 32;
        const float* r1 = r0 + 8 * 32;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const int px = kb * 8 + tg;
          FragA a;
          split_tf32(r0[px ^ sw], a.hi[0], a.lo[0]);
          split_tf32(r1[px ^ sw], a.hi[1], a.lo[1]);
          split_tf32(r0[(px + 4) ^ sw], a.hi[2], a.lo[2]);
          split_tf32(r1[(px + 4) ^ sw], a.hi[3], a.lo[3]);
          mma3(accB[mb], a, d[kb]);
        }
      }
    }
    __syncwarp();                                 // the warp is done with its rows of the tile: fetch the next one under B1
    if (t + 1 < te) fill(t + 1);
    // ---------------- phase B1: df^T [px 16] x [ch 8] per block, contraction over the 10 columns (8 + 2)
    {
      FragA w0[2], w1[2];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const float* p0 = wt + (mb * 16 + g) * kWtLd + tg;
        const float* p1 = p0 + 8 * kWtLd;
        split_tf32(p0[0], w0[mb].hi[0], w0[mb].lo[0]);
        split_tf32(p1[0], w0[mb].hi[1], w0[mb].lo[1]);
        split_tf32(p0[4], w0[mb].hi[2], w0[mb].lo[2]);
        split_tf32(p1[4], w0[mb].hi[3], w0[mb].lo[3]);
        split_tf32(p0[8], w1[mb].hi[0], w1[mb].lo[0]);      // columns 8 + tg: 10, 11 hold zeros
        split_tf32(p1[8], w1[mb].hi[1], w1[mb].lo[1]);
        w1[mb].hi[2] = w1[mb].hi[3] = w1[mb].lo[2] = w1[mb].lo[3] = 0u;
      }
      const int rem = hw - x0;                    // valid pixels of this tile (>= 32 except for the last one)
      const int sw = ((g >> 2) & 1) << 2;
      float* orow = dst + static_cast<long long>(wch + 2 * tg) * hw + x0 + g;
#pragma unroll 2
      for (int nb = 0; nb < CW / 8; ++nb) {
        const int ch = wch + nb * 8 + g;
        FragB b0, b1;
        split_tf32(t0[ch * 8 + (tg ^ sw)], b0.hi[0], b0.lo[0]);
        split_tf32(t0[ch * 8 + ((tg + 4) ^ sw)], b0.hi[1], b0.lo[1]);
        split_tf32(tg < kN1 ? t1[ch * kN1 + tg] : 0.f, b1.hi[0], b1.lo[0]);
        b1.hi[1] = b1.lo[1] = 0u;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          float d[4] = {0.f, 0.f, 0.f, 0.f};
          mma3(d, w1[mb], b1);
          mma3(d, w0[mb], b0);
          float* o = orow + mb * 16;
          if (mb * 16 + g < rem) {
            o[0] = d[0];
            o[hw] = d[1];
          }
          if (mb * 16 + g + 8 < rem) {
            o[8] = d[2];
            o[hw + 8] = d[3];
          }
        }
        orow += 8LL * hw;
      }
    }
  }
  float* dstp = part + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * (c + 1) * kK;
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    if (tg < kK / 2) {
      const int ch = wch + mb * 16 + g;
      *reinterpret_cast<float2*>(dstp + ch * kK + 2 * tg) = make_float2(accB[mb][0], accB[mb][1]);
      *reinterpret_cast<float2*>(dstp + (ch + 8) * kK + 2 * tg) = make_float2(accB[mb][2], accB[mb][3]);
    }
  }
  if (warp < 2 && lane == 0) {
#pragma unroll
    for (int k = 0; k < kP; ++k) dstp[c * kK + warp * kP + k] = dsum[k];
  }
}

}  // namespace

bool pemp_mpa_bwd_mma_shape(int c, int p) { return p == kP && (c == 128 || c == 256 || c == 512 || c == 1024); }

size_t pemp_mpa_bwd_mma_smem(int c) {
  return (static_cast<size_t>(c) * (32 + 8 + kN1) + kW * kNK * kRedLd + 32 * kWtLd + 32 * 8 + 2 * kK) * sizeof(float);
}

int pemp_mpa_bwd_mma_launch(const float* fts, long long ep, int S, const float* ctr, const float* coef, const float* beta,
                            const float* fg, const float* bg, long long mask_stride, int N, int c, int hw, int chunks, int ntiles,
                            float* dfts, long long d_ep, float* part, cudaStream_t st) {
  const size_t smem = pemp_mpa_bwd_mma_smem(c);
#define PEMP_BWD_MMA(MBV)                                                                                                        \
  do {                                                                                                                           \
    cudaError_t e = cudaFuncSetAttribute(mpa_bwd_mma_kernel<MBV>, cudaFuncAttributeMaxDynamicSharedMemorySize,                  \
                                         static_cast<int>(smem));                                                               \
    if (e != cudaSuccess) return static_cast<int>(e);                                                                            \
    mpa_bwd_mma_kernel<MBV><<<dim3(chunks, N), kT, smem, st>>>(fts, ep, S, ctr, coef, beta, fg, bg, mask_stride, hw, ntiles,    \
                                                               dfts, d_ep, part);                                               \
  } while (0)
  switch (c / 128) {
    case 1: PEMP_BWD_MMA(1); break;
    case 2: PEMP_BWD_MMA(2); break;
    case 4: PEMP_BWD_MMA(4); break;
    case 8: PEMP_BWD_MMA(8); break;
    default: return PEMP_E_SHAPE;
  }
#undef PEMP_BWD_MMA
  return PEMP_OK;
}
