// K1 / K8  masked average pooling at feature resolution.
//
// replaces  sum(f*m,-1)/(m.sum(-1)+eps); view(B,S,c).mean(1)
//   networks/pemp_stage1.py:223-227, pemp_stage2.py:196-200, canet.py:176-178, panet.py:181-186,
//   Weighted_GAP networks/pfenet.py:15-20; also the pooling half of K6 (baseline.py:105-110).
//
// Roofline: HBM.  Reads c*hw floats per image once, 2 FMA per float.  Layout fts [BS, c, hw].
//
// Mapping: grid = (pixel chunk, image).  A CTA owns one chunk of <= 32*kR pixels of one image; every lane
// keeps its kR mask weights (fg and bg) in registers for the whole kernel; each warp walks channel rows,
// two rows per step, so one step issues 2*kR independent coalesced loads per lane.  A row's two
// partial sums are reduced with shuffles and written to the workspace as part[image][chunk][c][2];
// `pool_finalize_kernel` adds the chunks in index order (deterministic), divides and averages the shots.
#include "common.cuh"

// TMA-fed persistent fast path for c in {256, 512} (pool_tma.cu); PEMP_E_ALIGN = not covered, nothing launched
int pemp_pool_tma_launch(const float* fts, long long ep_stride, const float* fg, const float* bg, long long mask_stride, int B,
                         int S, int c, int hw, float eps, const float* den_override, float* fg_proto, float* bg_proto, char* ws,
                         size_t ws_bytes, cudaStream_t st);
size_t pemp_pool_tma_workspace_bytes(int B, int S, int c, int hw);
#ifndef PEMP_POOL_TMA
#define PEMP_POOL_TMA 1
#endif
static int g_pool_path = 0;   // diagnostic switch, see pemp_debug_pool_path

namespace {

constexpr int kR = 16;        // max pixels per lane per chunk
constexpr int kWarps = 8;

__host__ __device__ inline int chunk_count(int hw) { return (hw + 32 * kR - 1) / (32 * kR); }
// balanced chunk length, multiple of 32
__host__ __device__ inline int chunk_len(int hw) {
  int n = chunk_count(hw);
  return ((hw + n - 1) / n + 31) / 32 * 32;
}

template <bool kTwo>
__global__ void __launch_bounds__(kWarps * 32)
pool_partial_kernel(const float* __restrict__ fts, long long ep_stride, int S, const float* __restrict__ fg, const float* __restrict__ bg,
                    long long mask_stride, int c, int hw, float* __restrict__ part, float* __restrict__ den) {
  const int chunk = blockIdx.x, img = blockIdx.y, nchunks = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int len = chunk_len(hw);
  const int x0 = chunk * len;
  const int R = len / 32;

  float mf[kR], mb[kR];
  const float* fgp = fg + img * mask_stride;
  const float* bgp = kTwo ? bg + img * mask_stride : nullptr;
#pragma unroll
  for (int i = 0; i < kR; ++i) {
    int x = x0 + lane + 32 * i;
    bool ok = i < R && x < hw;
    mf[i] = ok ? __ldg(fgp + x) : 0.f;
    mb[i] = (kTwo && ok) ? __ldg(bgp + x) : 0.f;
  }
  if (warp == 0) {   // denominators of this chunk
    float sf = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < kR; ++i) { sf += mf[i]; sb += mb[i]; }
    sf = warp_sum(sf);
    sb = warp_sum(sb);
    if (lane == 0) {
      den[(static_cast<long long>(img) * nchunks + chunk) * 2 + 0] = sf;
      den[(static_cast<long long>(img) * nchunks + chunk) * 2 + 1] = sb;
    }
  }

  const float* base = fts + (img / S) * ep_stride + static_cast<long long>(img % S) * c * hw + x0 + lane;
  float* out = part + (static_cast<long long>(img) * nchunks + chunk) * c * 2;
  for (int ch = warp * 2; ch < c; ch += kWarps * 2) {
    const bool two_rows = ch + 1 < c;
    const float* r0 = base + static_cast<long long>(ch) * hw;
    const float* r1 = r0 + hw;
    float v0[kR], v1[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      bool ok = i < R && x0 + lane + 32 * i < hw;
      v0[i] = ok ? __ldg(r0 + 32 * i) : 0.f;
      v1[i] = (ok && two_rows) ? __ldg(r1 + 32 * i) : 0.f;
    }
    float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      a0 = fmaf(v0[i], mf[i], a0);
      a1 = fmaf(v1[i], mf[i], a1);
      if (kTwo) {
        b0 = fmaf(v0[i], mb[i], b0);
        b1 = fmaf(v1[i], mb[i], b1);
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (kTwo) { b0 = warp_sum(b0); b1 = warp_sum(b1); }
    if (lane == 0) {
      out[ch * 2 + 0] = a0;
      out[ch * 2 + 1] = b0;
      if (two_rows) {
        out[ch * 2 + 2] = a1;
        out[ch * 2 + 3] = b1;
      }
    }
  }
}

// one thread per (b, channel): chunks summed in index order, ratio per shot, mean over shots
// (`fg_vecs.view(B,S,c).mean(1)`).  den_override (optional) [BS, 2] replaces the summed weights (K6 uses the
// exact full-resolution mask sums).
__global__ void pool_finalize_kernel(const float* __restrict__ part, const float* __restrict__ den,
                                     const float* __restrict__ den_override, int B, int S, int c, int nchunks, float eps,
                                     float* __restrict__ fg_proto, float* __restrict__ bg_proto) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * c) return;
  int b = i / c, ch = i - b * c;
  float accf = 0.f, accb = 0.f;
  for (int s = 0; s < S; ++s) {
    long long img = static_cast<long long>(b) * S + s;
    float nf = 0.f, nb = 0.f, df = 0.f, db = 0.f;
    for (int k = 0; k < nchunks; ++k) {
      const float* p = part + ((img * nchunks + k) * c + ch) * 2;
      nf += p[0];
      nb += p[1];
      df += den[(img * nchunks + k) * 2 + 0];
      db += den[(img * nchunks + k) * 2 + 1];
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

}  // namespace

extern "C" size_t pemp_map_pool_workspace_bytes(int B, int S, int c, int hw) {
  if (B <= 0 || S <= 0 || c <= 0 || hw <= 0) return 0;
  size_t imgs = static_cast<size_t>(B) * S, n = chunk_count(hw);
  size_t generic = align_up(imgs * n * c * 2 * sizeof(float), 256) + align_up(imgs * n * 2 * sizeof(float), 256);
  size_t tma = PEMP_POOL_TMA ? pemp_pool_tma_workspace_bytes(B, S, c, hw) : 0;
  return generic > tma ? generic : tma;
}

extern "C" int pemp_debug_pool_path(int mode) {
  const int old = g_pool_path;
  if (mode == 0 || mode == 1) g_pool_path = mode;
  return old;
}

// shared with fullres.cu / align.cu
int pemp_pool_launch(const float* fts, long long ep_stride, const float* fg, const float* bg, long long mask_stride, int B, int S, int c,
                     int hw, float eps, const float* den_override, float* fg_proto, float* bg_proto, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  PEMP_REQUIRE(fts && fg && fg_proto, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && c > 0 && hw > 0 && static_cast<long long>(B) * S <= 65535, PEMP_E_SHAPE);
  PEMP_REQUIRE(workspace && workspace_bytes >= pemp_map_pool_workspace_bytes(B, S, c, hw), PEMP_E_WORKSPACE);
  if (ep_stride == 0) ep_stride = static_cast<long long>(S) * c * hw;
  if (PEMP_POOL_TMA && g_pool_path != 1) {
    const int rc = pemp_pool_tma_launch(fts, ep_stride, fg, bg, mask_stride, B, S, c, hw, eps, den_override, fg_proto, bg_proto,
                                        static_cast<char*>(workspace), workspace_bytes, st);
    if (rc != PEMP_E_ALIGN) return rc;
  }
  const int n = chunk_count(hw);
  const size_t imgs = static_cast<size_t>(B) * S;
  float* part = static_cast<float*>(workspace);
  float* den = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up(imgs * n * c * 2 * sizeof(float), 256));
  dim3 grid(n, static_cast<unsigned>(imgs));
  if (bg)
    pool_partial_kernel<true><<<grid, kWarps * 32, 0, st>>>(fts, ep_stride, S, fg, bg, mask_stride, c, hw, part, den);
  else
    pool_partial_kernel<false><<<grid, kWarps * 32, 0, st>>>(fts, ep_stride, S, fg, nullptr, mask_stride, c, hw, part, den);
  int total = B * c;
  pool_finalize_kernel<<<(total + 255) / 256, 256, 0, st>>>(part, den, den_override, B, S, c, n, eps, fg_proto,
                                                            bg ? bg_proto : nullptr);
  return launch_status();
}

extern "C" int pemp_map_pool_lowres(const float* fts, long long fts_episode_stride, const float* fg, const float* bg, long long mask_stride, int B,
                                    int S, int c, int hw, float eps, float* fg_proto, float* bg_proto, void* workspace,
                                    size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(!bg || bg_proto, PEMP_E_NULL);
  return pemp_pool_launch(fts, fts_episode_stride, fg, bg, mask_stride, B, S, c, hw, eps, nullptr, fg_proto, bg_proto, workspace,
                          workspace_bytes, as_stream(stream));
}

extern "C" int pemp_weighted_gap(const float* supp_feat, const float* mask, int B, int c, int hw, float* out,
                                 void* workspace, size_t workspace_bytes, pemp_stream_t stream) {
  // Weighted_GAP (pfenet.py:15-20): avg_pool(f*m)*h*w / (avg_pool(m)*h*w + 0.0005) == sum(f*m)/(sum(m)+5e-4)
  return pemp_pool_launch(supp_feat, 0, mask, nullptr, hw, B, 1, c, hw, 0.0005f, nullptr, out, nullptr, workspace,
                          workspace_bytes, as_stream(stream));
}
