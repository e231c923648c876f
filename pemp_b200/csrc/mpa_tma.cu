// K2 fast path: meta-prototype attention as a persistent, TMA-fed, warp-specialised kernel (c = 512, P = 3).
//
// Same arithmetic as `mpa_kernel` (mpa.cu; reference networks/pemp_stage1.py:202-213, pemp_stage2.py:174-186):
//   dots   d[x, 0..3] = sum_c f[c, x] * table[c, 0..3]          (differences of squared distances, see mpa.cu)
//   A[k,x] = softmax_j(dots of class group g) * mask_g[x]       k = g*3 + j
//   num[c,k] += f[c,x] A[k,x],  den[k] += A[k,x]
// but the features never pass through registers on their way in:
//
//  * The support maps are [.., c, hw] fp32 with hw = 2601, so a channel row starts at an arbitrary 4-byte phase and
//    neither 128-bit loads nor a [c, hw] tensor map (row pitch 10404 B) are legal.  FOUR rows, however, are 16*hw
//    bytes: the tensor map describes the episode as  [S*c/4 groups][4*hw floats]  (pitch 16*hw B) and a box of
//    32 floats x 128 groups starting at inner coordinate  e*hw + x  holds pixels x.. of the channels 4g+e,
//    g = 0..127.  A box origin must itself be 16-byte aligned (tools/probes/tma_unaligned_probe.cu: any other inner
//    coordinate traps), so the origin is rounded down to a multiple of 4 floats: box column i of class e holds pixel
//    x_nom + i - o_e with o_e = (e*hw + x_nom) & 3, and a tile advances by 28 pixels so that the 28 nominal pixels
//    are inside the 32 columns of every class (+14 % shared-memory traffic, no extra DRAM traffic: the 4-column
//    overlap is an L2 hit).  Four boxes (e = 0..3) are one 512-channel tile (64 KB); they land in shared memory as
//    128-byte rows with the 128-byte swizzle.  Columns outside the image row (neighbouring rows' data, or zero fill
//    past the tensor) get zero weight.
//  * One CTA per SM, persistent: CTA b owns the flat tile range [T*b/G, T*(b+1)/G) of the launch (tiles of one image
//    are consecutive), so it touches 2-3 images and the pipeline never drains between them.  Warp 16 (one elected
//    lane) is the producer: a 12-slot ring of 16-KB boxes with a full / empty mbarrier per slot, i.e. up to 192 KB
//    of loads in flight per SM without a single load instruction or staging register in the consumers.
//  * Warps 0-15 are consumers (112 registers each via setmaxnreg; the producer warpgroup gives registers back).
//    Warp w = e + 4*cp (one channel class per warp scheduler) only ever touches box e of a tile:
//      phase A  rows [32cp, 32cp+32) of the box, lane <-> (row mod 4, 16-byte chunk): one conflict-free LDS.128 gives
//               4 pixels of a channel, the (pre-duplicated) table row comes with two more, 8 packed FFMA2 accumulate
//               4 pixels x 4 dots; a halving butterfly over the 4 row groups leaves each lane with the 4 dots of one
//               pixel, stored to part[tile parity][warp][dot][pixel]; the warp arrives on the buffer's mbarrier;
//      phase B  lane <-> rows {l, l+32, l+64, l+96} of box e (the swizzle makes the column read conflict-free):
//               for each live class group 12 FFMA2 per column pair accumulate {even, odd} column sums of
//               4 channels x 3 prototypes - 48 accumulator registers that live across the whole image.
//    Per iteration a consumer runs A(k+1), then B(k): the partial dots of tile k+1 are on their way to the softmax
//    warps while it accumulates tile k, so nobody waits for the slowest warp.
//  * Warps 17 and 18 (the producer's warpgroup) are the softmax warps, one per class group, lane <-> pixel of the tile:
//    wait for the 16 partial dot sets, add them in a fixed order, exp2 / reciprocal (the table carries log2 e), x mask,
//    and publish the 3 weights of every pixel in both column-pair alignments plus a bit mask of live pixels; they also
//    own the group's denominators.  Consumers therefore never compute a weight (in the first version every consumer
//    warp recomputed the weights of its 8 columns: 4 x redundant, 65 of 474 instructions per warp and tile).
//    At an image boundary the four warps of a box fold their accumulators through the box they just consumed
//    (fixed order cp = 0..3) and write one partial per (image, CTA); `mpa_tma_finalize_kernel` adds the partials of an
//    image in CTA order, divides and averages the shots - deterministic, no float atomics.
#include "tma_common.cuh"

int pemp_mpa_tma_launch(const float* fts, long long ep_stride, const float* ctr, const float* fg, const float* bg,
                        long long mask_stride, int B, int S, int hw, float eps, float* fg_proto, float* bg_proto,
                        float* adaptive_p, float* shot_centre, float* shot_den, char* ws, size_t ws_bytes, cudaStream_t st);
size_t pemp_mpa_tma_workspace_bytes(int B, int S, int hw);

namespace {

constexpr int kC = 512, kP = 3, kK = 6;
constexpr int kTW = 32;                          // floats per box row (one 128-byte swizzle row)
#ifndef PEMP_MPA_TMA_STEP
#define PEMP_MPA_TMA_STEP 28                     // 29 would still be covered by every class (o <= 3) but measured 6 % slower
#endif
constexpr int kStep = PEMP_MPA_TMA_STEP;         // pixels a tile advances
constexpr int kBoxRows = kC / 4;                 // 128 channel groups
constexpr int kBoxFloats = kBoxRows * kTW;       // 4096
constexpr uint32_t kBoxBytes = kBoxFloats * 4;   // 16 KB
#ifndef PEMP_MPA_TMA_DUP
#define PEMP_MPA_TMA_DUP 0                       // 1: table rows as {t0,t0,t1,t1,t2,t2,t3,t3} (FFMA2 operands as loaded, two LDS.128 per row);
                                                 // 0: {t0,t1,t2,t3}, one LDS.128 + four register copies.  Round 1 measured 1 as +6 %; with
                                                 // the softmax warps and the 12-slot ring the shared-memory pipe became the limit
                                                 // (ncu: 76 % busy, table loads 2/3 of phase A's wavefronts) and 0 is +3.5 % (r02)
#endif
#ifndef PEMP_MPA_TMA_SLOTS
#define PEMP_MPA_TMA_SLOTS 12                     // 3 tiles: two held by the consumers, one in flight (11 slots: -10 %)
#endif
constexpr int kTD = PEMP_MPA_TMA_DUP ? 8 : 4;    // floats per table row
constexpr int kNB = PEMP_MPA_TMA_SLOTS;
// whole tiles only: a slot must always carry the same channel class, or a parity wait can pass on the phase before
// (see cosine_tma.cu; the 11-slot builds of this kernel had that latent hazard)
static_assert(kNB % 4 == 0, "ring slots must be a multiple of the 4 boxes of a tile");
constexpr int kCons = 16;                        // consumer warps
constexpr int kThreadsT = (kCons + 4) * 32;      // + one producer warpgroup (setmaxnreg works on whole warpgroups)
constexpr int kRegsCons = 112, kRegsProd = 32;   // the CTA pool is the launch allocation (640 x 96) = 512 x 112 + 128 x 32
#ifndef PEMP_MPA_TMA_WPAIRS
#define PEMP_MPA_TMA_WPAIRS 18
#endif
constexpr int kWPairs = PEMP_MPA_TMA_WPAIRS;                      // column pairs per weight array: pixels -4 .. 31 of a tile (zero padded)
constexpr int kMaxGrid = 148;
#ifndef PEMP_MPA_TMA_PARTLD
#define PEMP_MPA_TMA_PARTLD 28                    // = kStep: every byte counts, the 12th ring slot needs the room
#endif
constexpr int kPartLd = PEMP_MPA_TMA_PARTLD;     // pixel pitch of a dot row in `part` (>= 28; any pitch is conflict-free: a
                                                 // warp instruction touches one dot row at 32 distinct pixels)

struct TmaSmem {
  alignas(1024) float ring[kNB][kBoxFloats];
  alignas(16) float table[kC * kTD];             // tile-row order: coefficients of channel 4g + e at row e*128 + g
  alignas(16) float part[2][kCons][4 * kPartLd];  // [tile parity][warp][dot][pixel]
  // weights of a tile, written by the two softmax warps: [tile parity][pairing][group][pixel pair]
  // [{w0a,w0b,w1a,w1b,w2a,w2b,-,-}]; pairing 0 pairs pixels (2q-4, 2q-3), pairing 1 pixels (2q-3, 2q-2), so a consumer
  // finds its column pair (columns 2i, 2i+1 = pixels 2i-o, 2i+1-o) as one aligned 32-byte record whatever o is
  alignas(16) float wpx[2][2][2][kWPairs * 8];
  unsigned wlive[2][2];                          // bit (pixel + 4) set <=> the pixel has a non-zero weight in the group
  alignas(8) uint64_t full[kNB];
  alignas(8) uint64_t empty[kNB];
  alignas(8) uint64_t part_bar[2];               // all 16 warps have written part[b]
  alignas(8) uint64_t wts_bar[2];                // both softmax warps have written wpx[b] / wlive[b]
  float konst[4];
};

using namespace pemp_tma;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreadsT, 1)
mpa_tma_kernel(const __grid_constant__ CUtensorMap map, int S, int hw, int nt_img, long long T, int imgs,
               const float* __restrict__ ctr, int* __restrict__ nparts, const float* __restrict__ fg,
               const float* __restrict__ bg, long long mask_stride, int maxp, float* __restrict__ part_num,
               float* __restrict__ part_den) {
  // the dynamic shared window starts 1024-byte aligned (declared alignment; checked once below), so every address in
  // `sm` is a compile-time offset and nothing has to be re-derived inside the tile loop
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TmaSmem& sm = *reinterpret_cast<TmaSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const long long t0 = T * cta / G, t1 = T * (cta + 1) / G;

  // The mbarriers come first so that the producer can start filling the ring while the other 19 warps build the table:
  // the first boxes are on their way during the ~3 us of prologue (matters for short launches: 64 one-shot images give
  // every CTA only 40 tiles).
  if (tid == 0) {
    for (int s = 0; s < kNB; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 4);
    }
    mbar_init(&sm.part_bar[0], kCons);
    mbar_init(&sm.part_bar[1], kCons);
    mbar_init(&sm.wts_bar[0], 2);
    mbar_init(&sm.wts_bar[1], 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // register re-allocation first (whole warpgroups), so that the producer can leave at once
  if (warp >= kCons) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProd));
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCons));
  }
  constexpr int kBuild = kCons * 32;                            // the 16 consumer warps build the table
  constexpr int kAfterBuild = kBuild + 64;                      // ... and the two softmax warps wait for it with them
  if (warp < kCons) {
    const int tix = tid;
    for (int i = tix; i < kC * 4; i += kBuild) {
      const int R = i >> 2, d = i & 3;
      const int ch = 4 * (R & (kBoxRows - 1)) + (R >> 7);
      const int g = d >> 1, j = (d & 1) + 1;
      // exact difference, then one rounding of the product with 2 log2(e): the dots come out in log2 units
      const float v = static_cast<float>(2.8853900817779268 *
                                         (static_cast<double>(__ldg(ctr + ch * kK + g * kP + j)) - __ldg(ctr + ch * kK + g * kP)));
      if (kTD == 8) {
        sm.table[R * 8 + 2 * d] = v;
        sm.table[R * 8 + 2 * d + 1] = v;
      } else {
        sm.table[R * 4 + d] = v;
      }
    }
    if (warp < 4) {
      const int g = warp >> 1, j = (warp & 1) + 1;
      double s = 0.0;
      for (int ch = lane; ch < kC; ch += 32) {
        const double a = __ldg(ctr + ch * kK + g * kP + j), b = __ldg(ctr + ch * kK + g * kP);
        s += (a - b) * (a + b);
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
      if (lane == 0) sm.konst[warp] = static_cast<float>(-s * 1.4426950408889634);
    }
    // partials an image ends up with = CTAs its tile range touches (64-bit divisions: once per image, for the finalize)
    for (int i = cta * kBuild + tix; i < imgs; i += G * kBuild) {
      const long long first = static_cast<long long>(i) * nt_img;
      nparts[i] = owner_of(first + nt_img - 1, T, G) - owner_of(first, T, G) + 1;
    }
    for (int i = tix; i < 2 * 2 * 2 * kWPairs * 8; i += kBuild) (&sm.wpx[0][0][0][0])[i] = 0.f;   // padding stays 0
    named_bar(1, kAfterBuild);                                // table / constants / zero padding visible to consumers + softmax warps
  }

  if (warp >= kCons) {
    // ============================ producer: one lane feeds the ring ============================
    if (warp == kCons && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
      int slot = 0;
      uint32_t par = 1;                               // a fresh barrier's "previous phase" counts as complete
      int img = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(img) * nt_img);
      int ep = img / S, s = img - ep * S;
      for (long long t = t0; t < t1; ++t) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = (e * hw + tl * kStep) & ~3;   // box origins must be 16-byte aligned
          mbar_wait(&sm.empty[slot], par);
          mbar_expect_tx(&sm.full[slot], kBoxBytes);
          tma_load_3d(&map, &sm.full[slot], sm.ring[slot], c0, s * kBoxRows, ep);
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
    } else if (warp == kCons + 1 || warp == kCons + 2) {
      // ============================ softmax warps: one per class group ============================
      // lane <-> pixel of the tile.  The warp waits for the 16 partial dot sets of a tile, adds them in a fixed order,
      // turns them into the 3 weights of its group (x mask) and publishes them in both pair alignments, together with
      // the live-pixel bit mask; it also owns the denominators of its group.  The consumers never compute a weight.
      const int g = warp - kCons - 1;
      named_bar(1, kAfterBuild);                              // konst / zeroed weight padding written by the consumer warps
      const float k0 = sm.konst[g * 2], k1 = sm.konst[g * 2 + 1];
      const float* mrow = g ? bg : fg;
      const int pl = lane < kStep ? lane : kStep - 1;             // clamped pixel for addressing
      const int P = lane + 4;
      float den[kP] = {0.f, 0.f, 0.f};
      int img = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(img) * nt_img);
      const int ntl = static_cast<int>(t1 - t0);
      for (int k = 0; k < ntl; ++k) {
        const int buf = k & 1, x_nom = tl * kStep;
        const bool valid = lane < kStep && x_nom + lane < hw;
        const float m = valid ? __ldg(mrow + (img * mask_stride + x_nom + lane)) : 0.f;
        mbar_wait(&sm.part_bar[buf], (k >> 1) & 1);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const float* pr = &sm.part[buf][0][2 * g * kPartLd + pl];
#pragma unroll
        for (int w2 = 0; w2 < kCons; w2 += 2) {
          s0 += pr[w2 * 4 * kPartLd];
          s1 += pr[(w2 + 1) * 4 * kPartLd];
          s2 += pr[w2 * 4 * kPartLd + kPartLd];
          s3 += pr[(w2 + 1) * 4 * kPartLd + kPartLd];
        }
        // dots are in log2 units (the table carries log2 e); m = 0 kills pixels outside the tile / image
        const float e1 = (s0 + s1) + k0, e2 = (s2 + s3) + k1;
        const float mx = fmaxf(0.f, fmaxf(e1, e2));
        // ex2.approx: 2 ulp, arguments <= 0; the sum is in [1, 3], rcp.approx is 1 ulp
        const float x0e = ex2_approx(0.f - mx), x1e = ex2_approx(e1 - mx), x2e = ex2_approx(e2 - mx);
        const float r = m * rcp_approx(x0e + x1e + x2e);
        const float w[kP] = {x0e * r, x1e * r, x2e * r};
        if (lane < kStep) {
          float* wa = &sm.wpx[buf][0][g][(P >> 1) * 8 + (P & 1)];
          float* wb = &sm.wpx[buf][1][g][((P - 1) >> 1) * 8 + ((P - 1) & 1)];
#pragma unroll
          for (int j = 0; j < kP; ++j) {
            den[j] += w[j];
            wa[2 * j] = w[j];
            wb[2 * j] = w[j];
          }
        }
        const unsigned lv = __ballot_sync(kFull, lane < kStep && ((w[0] != 0.f) | (w[1] != 0.f) | (w[2] != 0.f)));
        if (lane == 0) sm.wlive[buf][g] = lv << 4;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.wts_bar[buf]);
        if (tl == nt_img - 1 || k + 1 == ntl) {                    // image boundary: this CTA's denominators of the image
          const int slot_idx = cta - owner_of(static_cast<long long>(img) * nt_img, T, G);
#pragma unroll
          for (int j = 0; j < kP; ++j) {
            const float d = warp_sum(den[j]);
            if (lane == 0) part_den[(static_cast<long long>(img) * maxp + slot_idx) * 8 + g * kP + j] = d;
            den[j] = 0.f;
          }
        }
        if (++tl == nt_img) {
          tl = 0;
          ++img;
        }
      }
    }
    return;
  }

  // ============================ consumers ============================
  // a scheduler (warp & 3) hosts one channel class with all four column groups: with e = warp >> 2 it hosted one
  // column group of all classes, and the group whose last columns are outside the tile had less to do (+1.3 %)
  const int e = warp & 3, cp = warp >> 2;
  // phase-A lane roles
  const int rg = lane >> 3, jc = lane & 7;
  // Per-lane offsets that never change.  They pass through an empty `asm volatile` so the compiler keeps them in
  // registers instead of re-deriving them from %tid in every phase of every tile (measured: ~60 integer
  // instructions per warp and tile).
  int off_a = (cp * 32 + rg) * kTW + ((jc ^ rg) << 2);            // phase A, even i; odd i: ^ 16; row step i*128
  int off_t = (e * kBoxRows + cp * 32 + rg) * kTD;                // table row of (i = 0)
  int off_b = lane * kTW + (((2 * cp) ^ (lane & 7)) << 2);        // phase B, chunk 2cp; chunk 2cp+1: ^ 4; row step 32*i
  int col_a = 4 * jc + rg;                                        // box column this lane holds after the butterfly
  asm volatile("" : "+r"(off_a), "+r"(off_t), "+r"(off_b), "+r"(col_a));

  float2 acc[4][2][kP];                               // [row slot][group][prototype] = {even, odd column} sums
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int j = 0; j < kP; ++j) acc[i][g][j] = make_float2(0.f, 0.f);

  // ---- phase A of one tile: dots of the 32 box columns over this warp's 32 rows -> part[buf][warp] ----
  auto phase_a = [&](const float* box, int o, int buf) {
    float2 pa[2][4];                                  // [column pair][dot]
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int d = 0; d < 4; ++d) pa[p][d] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // row cp*32 + 4i + rg of the box; its swizzle phase is ((i & 1) << 2) | rg
      const float4 f = *reinterpret_cast<const float4*>(box + (off_a ^ ((i & 1) << 4)) + i * 4 * kTW);
      const float4* trow = reinterpret_cast<const float4*>(sm.table + off_t + i * 4 * kTD);
      const float4 ta = trow[0];
      const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
      float2 td[4];
      if (kTD == 8) {
        const float4 tb = trow[1];
        td[0] = make_float2(ta.x, ta.y), td[1] = make_float2(ta.z, ta.w);
        td[2] = make_float2(tb.x, tb.y), td[3] = make_float2(tb.z, tb.w);
      } else {
        td[0] = make_float2(ta.x, ta.x), td[1] = make_float2(ta.y, ta.y);
        td[2] = make_float2(ta.z, ta.z), td[3] = make_float2(ta.w, ta.w);
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        pa[0][d] = ffma2(f01, td[d], pa[0][d]);
        pa[1][d] = ffma2(f23, td[d], pa[1][d]);
      }
#ifdef PEMP_TMA_DEBUG_SHORT_A                        // timing experiments only (results are wrong)
      if (i == 0) break;
#endif
    }
    // halving butterfly over the row groups (lane bits 4 and 3): lane (rg, jc) ends with column 4*jc + rg
    const bool hi = (lane & 16) != 0, lo = (lane & 8) != 0;
    const int p = col_a - o;                          // pixel (relative to x_nom) of that column
    float* dst = &sm.part[buf][warp][p];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float2 keep = hi ? pa[1][d] : pa[0][d], send = hi ? pa[0][d] : pa[1][d];
      const float rx = keep.x + __shfl_xor_sync(kFull, send.x, 16);
      const float ry = keep.y + __shfl_xor_sync(kFull, send.y, 16);
      const float keep2 = lo ? ry : rx, send2 = lo ? rx : ry;
      const float q = keep2 + __shfl_xor_sync(kFull, send2, 8);
      if (p >= 0 && p < kStep) dst[d * kPartLd] = q;  // [dot][pixel], row pitch 40: conflict-free for writer and reader
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.part_bar[buf]);
  };

  // ---- phase B: rows {l, l+32, l+64, l+96} of the box x this warp's 8 columns ----
  // A class group is skipped when none of the 8 columns has a non-zero weight in it (one warp-uniform test per
  // group: with complementary masks most 8-pixel runs are all-foreground or all-background).
  auto phase_b = [&](const float* box, int o, int buf) {
#ifdef PEMP_TMA_DEBUG_SKIP_B                         // timing experiments only (results are wrong)
    return;
#endif
    // this warp's columns 8cp .. 8cp+7 are the pixels 8cp - o .. of the tile: bit (pixel + 4) of wlive, and the
    // column pair 2i, 2i+1 is record ((8cp + 2i - o + 4) - (o & 1)) / 2 of pairing o & 1
    const int sh = 8 * cp - o + 4;
    float4 f[2][4];
#pragma unroll
    for (int ck = 0; ck < 2; ++ck)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        f[ck][i] = *reinterpret_cast<const float4*>(box + (off_b ^ (ck << 2)) + 32 * i * kTW);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (((sm.wlive[buf][g] >> sh) & 0xffu) == 0) continue;
      const float* wts = &sm.wpx[buf][o & 1][g][((sh - (o & 1)) >> 1) * 8];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const float4 wa = *reinterpret_cast<const float4*>(wts + pp * 8);
        const float2 wb = *reinterpret_cast<const float2*>(wts + pp * 8 + 4);
        const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 fv = f[pp >> 1][i];
          const float2 fp = (pp & 1) ? make_float2(fv.z, fv.w) : make_float2(fv.x, fv.y);
          acc[i][g][0] = ffma2(fp, w0, acc[i][g][0]);
          acc[i][g][1] = ffma2(fp, w1, acc[i][g][1]);
          acc[i][g][2] = ffma2(fp, wb, acc[i][g][2]);
        }
      }
    }
  };

  // Software pipeline: per iteration  A(k+1), B(k).  The partial dots of tile k+1 are handed to the softmax warps
  // (mbarrier arrive) before B(k); the weights of tile k were requested one B and one A earlier, so the wait on them is
  // normally free.  Buffer reuse: the softmax warps write wpx[b] for tile k+2 only after all 16 arrivals of tile k+2,
  // i.e. after every consumer finished B(k), the last reader of wpx[b]; a consumer rewrites part[b] in A(k+2), after its
  // B(k), which waited for the weights of tile k and hence for the softmax warps' last read of part[b].
  const int ntl = static_cast<int>(t1 - t0);
  if (ntl <= 0) return;
  int slot = e;
  uint32_t par = 0;
  int img = static_cast<int>(t0 / nt_img), tl = static_cast<int>(t0 - static_cast<long long>(img) * nt_img);
  int o = (e * hw + tl * kStep) & 3;
  mbar_wait(&sm.full[slot], par);
  phase_a(sm.ring[slot], o, 0);
  for (int k = 0; k < ntl; ++k) {
    const bool have_next = k + 1 < ntl;
    const bool last_of_img = (tl == nt_img - 1) || !have_next;
    // next tile
    int slot_n = slot + 4;
    uint32_t par_n = par;
    if (slot_n >= kNB) {
      slot_n -= kNB;
      par_n ^= 1;
    }
    int tl_n = tl + 1, img_n = img;
    if (tl_n == nt_img) {
      tl_n = 0;
      ++img_n;
    }
    const int o_n = (e * hw + tl_n * kStep) & 3;
    if (have_next) {
      mbar_wait(&sm.full[slot_n], par_n);
      phase_a(sm.ring[slot_n], o_n, (k + 1) & 1);
    }
    mbar_wait(&sm.wts_bar[k & 1], (k >> 1) & 1);
    phase_b(sm.ring[slot], o, k & 1);

    if (!last_of_img) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
    } else {
      // ---------------- image boundary: fold the four column-owning warps of this box, write the partial ----------
      float* scratch = sm.ring[slot];                 // [3][24][32] numerators
      named_bar(2 + e, 128);                          // all four warps are done reading the box
      if (cp > 0) {
        int o2 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int j = 0; j < kP; ++j) scratch[((cp - 1) * 24 + (o2++)) * 32 + lane] = acc[i][g][j].x + acc[i][g][j].y;
      }
      named_bar(2 + e, 128);
      if (cp == 0) {
        const int slot_idx = cta - owner_of(static_cast<long long>(img) * nt_img, T, G);
        float* out = part_num + (static_cast<long long>(img) * maxp + slot_idx) * (kC * kK);
        int o2 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[kK];
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int j = 0; j < kP; ++j) {
              float a2 = acc[i][g][j].x + acc[i][g][j].y;
#pragma unroll
              for (int r = 0; r < 3; ++r) a2 += scratch[(r * 24 + o2) * 32 + lane];
              ++o2;
              v[g * kP + j] = a2;
            }
          float2* dst = reinterpret_cast<float2*>(out + (e * kBoxRows + lane + 32 * i) * kK);
          dst[0] = make_float2(v[0], v[1]);
          dst[1] = make_float2(v[2], v[3]);
          dst[2] = make_float2(v[4], v[5]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // scratch writes before the slot's next TMA fill
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int j = 0; j < kP; ++j) acc[i][g][j] = make_float2(0.f, 0.f);
    }

    o = o_n;
    slot = slot_n;
    par = par_n;
    tl = tl_n;
    img = img_n;
  }
}

// one thread per (b, channel, k): add the partials of every shot in CTA order, divide, average the shots
__global__ void mpa_tma_finalize_kernel(const float* __restrict__ part_num, const float* __restrict__ part_den,
                                        const int* __restrict__ nparts, int B, int S, int maxp, float eps,
                                        float* __restrict__ fg_proto, float* __restrict__ bg_proto,
                                        float* __restrict__ adaptive_p, float* __restrict__ shot_centre,
                                        float* __restrict__ shot_den) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * kC * kK) return;
  // thread <-> (b, tile row R, k): consecutive threads read consecutive floats of a partial (the first version walked the
  // channels in natural order, i.e. 24-byte pieces 3 KB apart); the few output stores are the strided side now
  const int k = static_cast<int>(i % kK);
  const long long t = i / kK;
  const int R = static_cast<int>(t % kC), b = static_cast<int>(t / kC);
  const int ch = 4 * (R & (kBoxRows - 1)) + (R >> 7);
  float accum = 0.f;
  for (int s = 0; s < S; ++s) {
    const long long img = static_cast<long long>(b) * S + s;
    const int n = __ldg(nparts + img);
    float num = 0.f, den = 0.f;
    // four partials per batch with all eight loads in flight (an image has 2-3 partials at the bench shape); the sums still run
    // in CTA order, and a missing partial adds an exact 0
    for (int sp0 = 0; sp0 < n; sp0 += 4) {
      float nv[4], dv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool on = sp0 + j < n;
        nv[j] = on ? part_num[((img * maxp + sp0 + j) * kC + R) * kK + k] : 0.f;
        dv[j] = on ? part_den[(img * maxp + sp0 + j) * 8 + k] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        num += nv[j];
        den += dv[j];
      }
    }
    accum += num / (den + eps);
    if (shot_centre) {      // training forward (see mpa.cu)
      shot_centre[(img * kC + ch) * kK + k] = num / (den + eps);
      if (ch == 0) shot_den[img * kK + k] = den + eps;
    }
  }
  const float v = accum / static_cast<float>(S);
  const int g = k / kP, j = k - g * kP;
  (g == 0 ? fg_proto : bg_proto)[(static_cast<long long>(b) * kC + ch) * kP + j] = v;
  if (adaptive_p) adaptive_p[(static_cast<long long>(b) * kC + ch) * kK + k] = v;
}


struct TmaPlan {
  int G, nt_img, maxp;
  long long T;
  size_t off_table, off_konst, off_nparts, off_num, off_den, total;
};
TmaPlan make_tma_plan(int B, int S, int hw) {
  TmaPlan p;
  p.nt_img = (hw + kStep - 1) / kStep;
  p.T = static_cast<long long>(B) * S * p.nt_img;
  // one CTA per SM; a CTA wants at least a few tiles to amortise its prologue
  long long g = p.T / 4;
  p.G = static_cast<int>(g < 1 ? 1 : (g > kMaxGrid ? kMaxGrid : g));
  // partial slots per image: an image's nt_img tiles are spread over at most this many consecutive CTAs
  const long long per_cta = p.T / p.G;                // every CTA owns per_cta or per_cta + 1 tiles
  p.maxp = static_cast<int>((p.nt_img + per_cta - 1) / per_cta) + 1;
  const size_t imgs = static_cast<size_t>(B) * S;
  p.off_table = 0;
  p.off_konst = kC * kTD * sizeof(float);
  p.off_nparts = p.off_konst + 256;
  p.off_num = p.off_nparts + align_up(imgs * sizeof(int), 256);
  p.off_den = p.off_num + align_up(imgs * p.maxp * kC * kK * sizeof(float), 256);
  p.total = p.off_den + align_up(imgs * p.maxp * 8 * sizeof(float), 256);
  return p;
}

}  // namespace

size_t pemp_mpa_tma_workspace_bytes(int B, int S, int hw) { return make_tma_plan(B, S, hw).total; }

// Returns PEMP_E_ALIGN (nothing launched) when the operand cannot be described by a tensor map; the caller then
// uses the generic kernel.
int pemp_mpa_tma_launch(const float* fts, long long ep_stride, const float* ctr, const float* fg, const float* bg,
                        long long mask_stride, int B, int S, int hw, float eps, float* fg_proto, float* bg_proto,
                        float* adaptive_p, float* shot_centre, float* shot_den, char* ws, size_t ws_bytes, cudaStream_t st) {
  const long long eps_stride = ep_stride ? ep_stride : static_cast<long long>(S) * kC * hw;
  if (hw < kTW) return PEMP_E_ALIGN;
  const TmaPlan pl = make_tma_plan(B, S, hw);
  PEMP_REQUIRE(ws_bytes >= pl.total, PEMP_E_WORKSPACE);
  CUtensorMap map;
  if (!make_rows4_map(&map, fts, B, S, kC, hw, eps_stride)) return PEMP_E_ALIGN;

  float* num = reinterpret_cast<float*>(ws + pl.off_num);
  float* den = reinterpret_cast<float*>(ws + pl.off_den);
  int* nparts = reinterpret_cast<int*>(ws + pl.off_nparts);
  const size_t smem = sizeof(TmaSmem);
  cudaError_t e = cudaFuncSetAttribute(mpa_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  mpa_tma_kernel<<<pl.G, kThreadsT, smem, st>>>(map, S, hw, pl.nt_img, pl.T, B * S, ctr, nparts, fg, bg, mask_stride, pl.maxp,
                                               num, den);
  const long long total = static_cast<long long>(B) * kC * kK;
  mpa_tma_finalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      num, den, nparts, B, S, pl.maxp, eps, fg_proto, bg_proto, adaptive_p, shot_centre, shot_den);
  return launch_status();
}
