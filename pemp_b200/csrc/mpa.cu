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
// result.  Phase B is the weighted sum; for a pixel whose group mask is zero the P products are skipped
// (warp-uniform test), which halves the work for complementary fg/bg masks.
//
// Mapping.  grid = (pixel split, image).  A CTA walks tiles of 32 pixels.
//   phase A  lane <-> pixel, warp <-> a quarter-of-channels stripe: four channel rows per step are loaded with
//            coalesced 128-byte row segments, multiplied against the difference table (one broadcast
//            128-bit smem load per row) and stored TRANSPOSED into shared memory as Ft[pixel][channel]
//            (row stride c+4 floats => conflict-free 128-bit stores);
//   softmax  64 threads turn the reduced dots into the 2P weights of each pixel;
//   phase B  thread <-> 4 consecutive channels x half of the tile's pixels: one conflict-free 128-bit smem
//            load per pixel feeds 4*P (or 8*P) FMAs into register accumulators that live for the whole CTA.
// Partial numerators / denominators go to the workspace per (image, split); `mpa_finalize_kernel` adds the
// splits in index order, divides, and averages the shots - deterministic, no atomics.
#include "common.cuh"

namespace {

constexpr int kTW = 32;          // pixels per tile
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxChannels = 1536;      // Ft[32][c+4] must fit the 227 KB of shared memory

// floats reserved for the transposed tile Ft[kTW][4*kQPT*tpp + 4]; the epilogue reuses it as fold[ng][kQPT*8*P][tpp]
#ifndef PEMP_MPA_QPT
#define PEMP_MPA_QPT 1
#endif
#ifndef PEMP_MPA_U
#define PEMP_MPA_U 8
#endif
#ifndef PEMP_MPA_PF
#define PEMP_MPA_PF 0
#endif
#ifndef PEMP_MPA_INTERLEAVE
#define PEMP_MPA_INTERLEAVE 1
#endif
#ifndef PEMP_MPA_WAVES
#define PEMP_MPA_WAVES 4
#endif
constexpr int kQPT = PEMP_MPA_QPT;     // channel quads per phase-B thread
__host__ __device__ inline int ft_floats(int c, int P) {
  int need = ((c >> 2) + kQPT - 1) / kQPT, tpp = 32;
  while (tpp < need) tpp <<= 1;
  int tile = kTW * (4 * kQPT * tpp + 4), fold = kThreads * kQPT * 2 * 4 * P;
  return tile > fold ? tile : fold;
}
__host__ __device__ inline int nd_of(int P) { return 2 * (P - 1); }          // dot products per pixel
__host__ __device__ inline int ndp_of(int P) { return P <= 3 ? 4 : 8; }      // padded table row

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
// Phase-B thread layout for c channels (quads = c/4): each thread owns kQPT = 2 channel quads
// {q, q + tpp} of one pixel group; tpp = threads per pixel group (power of two, >= 32 so a warp never
// straddles groups), ng = kThreads / tpp pixel groups of ppg = kTW / ng pixels.
__host__ __device__ inline int tpp_of(int c) {
  int need = ((c >> 2) + kQPT - 1) / kQPT, t = 32;
  while (t < need) t <<= 1;
  return t;
}

template <int P>
__global__ void __launch_bounds__(kThreads, 2)
mpa_kernel(const float* __restrict__ fts, long long ep_stride, int S, const float* __restrict__ table_g,
           const float* __restrict__ konst_g, const float* __restrict__ fg, const float* __restrict__ bg,
           long long mask_stride, int c, int hw, int tiles_per_split, float* __restrict__ part_num,
           float* __restrict__ part_den) {
  constexpr int ND = 2 * (P - 1);
  constexpr int NDP = P <= 3 ? 4 : 8;
  constexpr int K = 2 * P;
  extern __shared__ __align__(16) float smem[];
  const int tpp = tpp_of(c);
  const int ldf = 4 * kQPT * tpp + 4;                             // Ft row stride: >= c, = 4 * odd  => conflict-free
  float* Ft = smem;                                        // [kTW][ldf] (also the epilogue's fold buffer)
  float* table = Ft + ft_floats(c, P);                     // [c][NDP]
  float* red = table + (ND ? c * NDP : 0);                 // [kWarps][NDP][kTW]
  float* wgt = red + (ND ? kWarps * NDP * kTW : 0);        // [2][kTW][4]   softmax * mask
  __shared__ float konst[8];
  __shared__ unsigned live_mask[2];                        // bit x set <=> group g has a non-zero weight at pixel x

  const int split = blockIdx.x, img = blockIdx.y, nsplit = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = (hw + kTW - 1) / kTW;
#if PEMP_MPA_INTERLEAVE
  // CTA `split` of an image walks tiles split, split + nsplit, ...: the CTAs of one image (adjacent block ids, so
  // co-resident) touch adjacent 128-byte pieces of every channel row at about the same time (DRAM page locality)
  const int t_begin = 0, t_end = (ntiles - split + nsplit - 1) / nsplit;
#define PEMP_TILE(t) (split + (t) * nsplit)
#else
  const int t_begin = split * tiles_per_split, t_end = min(ntiles, t_begin + tiles_per_split);
#define PEMP_TILE(t) (t)
#endif
  const int quads = c >> 2;

  if (ND) {
    for (int i = tid; i < c * NDP; i += kThreads) table[i] = __ldg(table_g + i);
    if (tid < ND) konst[tid] = __ldg(konst_g + tid);
  }
  // channels beyond c in the padded Ft rows are never written by phase A: clear them once
  for (int i = tid; i < kTW * (ldf - c); i += kThreads) Ft[(i / (ldf - c)) * ldf + c + i % (ldf - c)] = 0.f;

  const float* img_base = fts + (img / S) * ep_stride + static_cast<long long>(img % S) * c * hw;
  const float* fgp = fg + img * mask_stride;
  const float* bgp = bg + img * mask_stride;

  // phase-A ownership: contiguous quads per warp
  const int qw = (quads + kWarps - 1) / kWarps;
  const int q_lo = warp * qw, q_hi = min(quads, q_lo + qw);
  // phase-B ownership (tpp is a power of two)
  const int tpp_shift = 31 - __clz(tpp);
  const int ng = kThreads >> tpp_shift, ppg = kTW / ng;
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
  float den[P];   // threads < 64: denominators of group tid/32, summed over this lane's pixels
#pragma unroll
  for (int j = 0; j < P; ++j) den[j] = 0.f;

  __syncthreads();

  // Software pipeline: the first kPF quads of a warp's share (all of it for c <= 512) are loaded one tile ahead
  // into registers, so the row loads of tile t+1 are in flight while tile t is in its softmax / phase-B part.
  constexpr int kPF = PEMP_MPA_PF;
  float v[kPF ? kPF : 1][4];
  auto prefetch = [&](int tile) {
    const int xx = min(PEMP_TILE(tile) * kTW + lane, hw - 1);        // dead pixels read a valid address; their weights are zero
    const float* prow = img_base + static_cast<long long>(q_lo * 4) * hw + xx;
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      if (q_lo + u < q_hi) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[u][e] = __ldg(prow);
          prow += hw;
        }
      }
    }
  };
  if (t_begin < t_end) prefetch(t_begin);

  for (int t = t_begin; t < t_end; ++t) {
    const int x = PEMP_TILE(t) * kTW + lane;
    const bool live = x < hw;
    const int xc = live ? x : hw - 1;

    // ---------------- phase A: dot with the difference table, transpose into Ft ----------------------
    float pd[NDP];
#pragma unroll
    for (int d = 0; d < NDP; ++d) pd[d] = 0.f;
    auto consume = [&](int qq, const float (&r)[4]) {
      if (ND) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4* t4 = reinterpret_cast<const float4*>(table + (qq * 4 + e) * NDP);
          float4 ta = t4[0];
          pd[0] = fmaf(r[e], ta.x, pd[0]);
          pd[1] = fmaf(r[e], ta.y, pd[1]);
          if (ND > 2) {
            pd[2] = fmaf(r[e], ta.z, pd[2]);
            pd[3] = fmaf(r[e], ta.w, pd[3]);
          }
          if (ND > 4) {
            float4 tb = t4[1];
            pd[4] = fmaf(r[e], tb.x, pd[4]);
            pd[5] = fmaf(r[e], tb.y, pd[5]);
          }
        }
      }
      *reinterpret_cast<float4*>(Ft + lane * ldf + qq * 4) = make_float4(r[0], r[1], r[2], r[3]);
    };
#pragma unroll
    for (int u = 0; u < kPF; ++u)
      if (q_lo + u < q_hi) consume(q_lo + u, v[u]);
    // channels beyond the prefetched share (c > 128 * kPF): plain load-then-use, U quads at a time
    constexpr int U = PEMP_MPA_U;
    if (q_lo + kPF < q_hi) {
      const float* prow = img_base + static_cast<long long>((q_lo + kPF) * 4) * hw + xc;
      for (int q = q_lo + kPF; q < q_hi; q += U) {
        float r[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (q + u < q_hi) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              r[u][e] = __ldg(prow);
              prow += hw;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (q + u < q_hi) consume(q + u, r[u]);
      }
    }
    if (ND) {
#pragma unroll
      for (int d = 0; d < ND; ++d) red[(warp * NDP + d) * kTW + lane] = pd[d];
    }
    __syncthreads();

    // ---------------- softmax weights: threads 0..31 foreground group, 32..63 background group ----------
    if (tid < 64) {
      const int g = tid >> 5;
      float m = live ? __ldg((g ? bgp : fgp) + x) : 0.f;
      float e[P];
      e[0] = 0.f;
      float mx = 0.f;
#pragma unroll
      for (int j = 1; j < P; ++j) {
        float s = 0.f;
        for (int wv = 0; wv < kWarps; ++wv) s += red[(wv * NDP + g * (P - 1) + j - 1) * kTW + lane];
        e[j] = s + konst[g * (P - 1) + j - 1];
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
      *reinterpret_cast<float4*>(wgt + (g * kTW + lane) * 4) = make_float4(w4[0], w4[1], w4[2], w4[3]);
      const unsigned bal = __ballot_sync(kFull, any);
      if (lane == 0) live_mask[g] = bal;
    }
    if (t + 1 < t_end) prefetch(t + 1);
    __syncthreads();

    // ---------------- phase B: out[c, k] += f[c, x] * A[k, x] over the pixels whose weights are non-zero
    // (warp-uniform pixel lists from the two bit masks; two pixels per step for load/FMA overlap)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      unsigned m = (live_mask[g] >> (grp * ppg)) & (ppg == 32 ? 0xffffffffu : ((1u << ppg) - 1u));
      while (m) {
        const int i0 = __ffs(m) - 1;
        m &= m - 1;
        const bool two = m != 0;
        const int i1 = two ? __ffs(m) - 1 : i0;
        m &= m - 1;
        const int p0 = grp * ppg + i0, p1 = grp * ppg + i1;
        const float4 w0 = *reinterpret_cast<const float4*>(wgt + (g * kTW + p0) * 4);
        float4 w1 = *reinterpret_cast<const float4*>(wgt + (g * kTW + p1) * 4);
        if (!two) w1 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float wa[4] = {w0.x, w0.y, w0.z, w0.w}, wb[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int a = 0; a < kQPT; ++a) {
          const float4 f0 = *reinterpret_cast<const float4*>(Ft + p0 * ldf + (qb + a * tpp) * 4);
          const float4 f1 = *reinterpret_cast<const float4*>(Ft + p1 * ldf + (qb + a * tpp) * 4);
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
    __syncthreads();
  }

  // ---------------- epilogue: fold the pixel groups (fixed order), write the partials ------------------
  float* fold = Ft;   // reuse: [ng][2*2*4*P][tpp]
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
  __syncthreads();
  if (grp == 0) {
    float* out = part_num + (static_cast<long long>(img) * nsplit + split) * c * K;
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
  if (tid < 64) {
    const int g = tid >> 5;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      float s = warp_sum(den[j]);
      if (lane == 0) part_den[(static_cast<long long>(img) * nsplit + split) * K + g * P + j] = s;
    }
  }
}

// one thread per (b, channel, k)
__global__ void mpa_finalize_kernel(const float* __restrict__ part_num, const float* __restrict__ part_den, int B, int S,
                                    int c, int P, int nsplit, float eps, float* __restrict__ fg_proto,
                                    float* __restrict__ bg_proto, float* __restrict__ adaptive_p) {
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
  }
  float v = accum / static_cast<float>(S);
  int g = k / P, j = k - g * P;
  (g == 0 ? fg_proto : bg_proto)[(static_cast<long long>(b) * c + ch) * P + j] = v;
  if (adaptive_p) adaptive_p[(static_cast<long long>(b) * c + ch) * K + k] = v;
}

int pick_splits(int imgs, int ntiles) {
  // aim for >= 4 waves of 3 CTAs/SM on 148 SMs, never more splits than tiles, at least 2 tiles per split
  int want = (PEMP_MPA_WAVES * 148 * 2 + imgs - 1) / imgs;
  int cap = ntiles / 2 > 0 ? ntiles / 2 : 1;
  int n = want < cap ? want : cap;
  return n < 1 ? 1 : n;
}

struct Plan {
  int nsplit, tiles_per_split;
  size_t off_table, off_konst, off_num, off_den, total;
};
Plan make_plan(int B, int S, int c, int hw, int P) {
  Plan p;
  const size_t imgs = static_cast<size_t>(B) * S;
  const int ntiles = (hw + kTW - 1) / kTW;
  p.nsplit = pick_splits(static_cast<int>(imgs), ntiles);
  p.tiles_per_split = (ntiles + p.nsplit - 1) / p.nsplit;
  p.nsplit = (ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.off_table = 0;
  p.off_konst = align_up(static_cast<size_t>(c) * ndp_of(P) * sizeof(float), 256);
  p.off_num = p.off_konst + 256;
  p.off_den = p.off_num + align_up(imgs * p.nsplit * c * 2 * P * sizeof(float), 256);
  p.total = p.off_den + align_up(imgs * p.nsplit * 2 * P * sizeof(float), 256);
  return p;
}

template <int P>
int launch(const float* fts, long long ep_stride, const float* ctr, const float* fg, const float* bg, long long mask_stride, int B, int S,
           int c, int hw, float eps, float* fg_proto, float* bg_proto, float* adaptive_p, char* ws, const Plan& pl,
           cudaStream_t st) {
  constexpr int ND = 2 * (P - 1), NDP = P <= 3 ? 4 : 8;
  float* table = reinterpret_cast<float*>(ws + pl.off_table);
  float* konst = reinterpret_cast<float*>(ws + pl.off_konst);
  float* num = reinterpret_cast<float*>(ws + pl.off_num);
  float* den = reinterpret_cast<float*>(ws + pl.off_den);
  if (ND) mpa_prepare_kernel<<<8, 256, 0, st>>>(ctr, c, P, table, konst);
  size_t smem = (static_cast<size_t>(ft_floats(c, P)) + (ND ? static_cast<size_t>(c) * NDP + kWarps * NDP * kTW : 0) +
                 2 * kTW * 4) * sizeof(float);
  if (smem > 227 * 1024) return PEMP_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(mpa_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid(pl.nsplit, static_cast<unsigned>(B) * S);
  mpa_kernel<P><<<grid, kThreads, smem, st>>>(fts, ep_stride ? ep_stride : static_cast<long long>(S) * c * hw, S, table, konst, fg, bg, mask_stride, c, hw, pl.tiles_per_split, num,
                                                  den);
  long long total = static_cast<long long>(B) * c * 2 * P;
  mpa_finalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(num, den, B, S, c, P, pl.nsplit, eps,
                                                                                 fg_proto, bg_proto, adaptive_p);
  return launch_status();
}

}  // namespace

extern "C" size_t pemp_meta_proto_attn_workspace_bytes(int B, int S, int c, int hw, int p) {
  if (B <= 0 || S <= 0 || c <= 0 || hw <= 0 || p < 1 || p > 4) return 0;
  return make_plan(B, S, c, hw, p).total;
}

extern "C" int pemp_meta_proto_attn(const float* fts, long long fts_episode_stride, const float* ctr, const float* fg, const float* bg,
                                    long long mask_stride, int B, int S, int c, int hw, int p, float eps, float* fg_proto,
                                    float* bg_proto, float* adaptive_p, void* workspace, size_t workspace_bytes,
                                    pemp_stream_t stream) {
  PEMP_REQUIRE(fts && ctr && fg && bg && fg_proto && bg_proto, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && c > 0 && hw > 0 && static_cast<long long>(B) * S <= 65535, PEMP_E_SHAPE);
  PEMP_REQUIRE(p >= 1 && p <= 4 && c % 4 == 0 && c <= kMaxChannels, PEMP_E_SHAPE);
  Plan pl = make_plan(B, S, c, hw, p);
  PEMP_REQUIRE(workspace && workspace_bytes >= pl.total, PEMP_E_WORKSPACE);
  PEMP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, PEMP_E_ALIGN);
  char* ws = static_cast<char*>(workspace);
  cudaStream_t st = as_stream(stream);
#define PEMP_MPA(PP) \
  return launch<PP>(fts, fts_episode_stride, ctr, fg, bg, mask_stride, B, S, c, hw, eps, fg_proto, bg_proto, adaptive_p, ws, pl, st)
  switch (p) {
    case 1: PEMP_MPA(1);
    case 2: PEMP_MPA(2);
    case 3: PEMP_MPA(3);
    default: PEMP_MPA(4);
  }
#undef PEMP_MPA
}
