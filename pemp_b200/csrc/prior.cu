// K9  PFENet prior mask: per shot a dense [HWs x C] . [C x HWq] contraction with the cosine normalisation,
// the max over support pixels and the min-max normalisation over query pixels fused around it.
//
// replaces networks/pfenet.py:201-231 (per shot: two torch.norm, two bmm, a [B, HWs, HWq] similarity matrix
// of 52 MB written and re-read, max, min, max, normalise; then cat/mean).
//
// Work: 2*C*HWs*HWq FLOP per shot (53.1 GFLOP at C=2048, 60x60) - the only tensor-core-bound op of the head.
// Paths (argument `precision`):
//   2  3-term bf16 split on tcgen05 (prior_tc.cu), fp32-grade: THE PRODUCT PATH (default of the drop-in);
//   0  single bf16 product on tcgen05 (prior_tc.cu): 3x faster, cosines to 3e-3 (stated tolerance);
//   1  fp32 CUDA-core tiled GEMM (this file): a slow, independent implementation kept ONLY as the on-device anchor the
//      tests compare the tensor-core paths with (12 ms per 5-shot episode; nothing in the product selects it).
#include "common.cuh"

int pemp_prior_tc_launch(const float* q4, const float* s4, const float* smask, float* nq, float* ns, int B,
                         int S, int C, int hw_s, int hw_q, int precision, float* rowmax, float* prior, char* ws,
                         size_t ws_bytes, cudaStream_t st);
size_t pemp_prior_tc_workspace_bytes(int B, int S, int C, int hw_s, int hw_q, int precision);

namespace {

constexpr float kEps = 1e-7f;   // cosine_eps, pfenet.py:202

// x [planes][C][hw] (optionally times mask [planes][hw]) -> out [planes][hw] = sqrt(sum_c (x*m)^2)
// block = 32 pixels x 8 channel phases: coalesced 128-byte row segments, 8-way split of the channel loop
__global__ void col_norm_kernel(const float* __restrict__ x, const float* __restrict__ mask, int C, int hw,
                                float* __restrict__ out) {
  __shared__ float part[8][32];
  const int pl = blockIdx.y, tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 32 + tx;
  const bool ok = i < hw;
  const float m = (ok && mask) ? __ldg(mask + static_cast<long long>(pl) * hw + i) : 1.f;
  const float* p = x + static_cast<long long>(pl) * C * hw + (ok ? i : 0);
  float s = 0.f;
#pragma unroll 8
  for (int c = ty; c < C; c += 8) {
    float v = ok ? __ldg(p + static_cast<long long>(c) * hw) * m : 0.f;
    s = fmaf(v, v, s);
  }
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && ok) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += part[r][tx];
    out[static_cast<long long>(pl) * hw + i] = sqrtf(t);
  }
}

// ---- fp32 path ----------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 8, kGemmThreads = 256;

// grid (ceil(hw_q / BM), S*B).  M = query pixels i, N = support pixels j, K = channels.
__global__ void __launch_bounds__(kGemmThreads)
prior_fp32_kernel(const float* __restrict__ q4, const float* __restrict__ s4, const float* __restrict__ smask,
                  const float* __restrict__ nq, const float* __restrict__ ns, int B, int C, int hw_s, int hw_q,
                  float* __restrict__ rowmax) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ float nqs[BM], nss[BN], ms[BN];
  __shared__ float red[BM][17];

  const int sb = blockIdx.y;             // s * B + b
  const int b = sb % B;
  const int i0 = blockIdx.x * BM;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* qb = q4 + static_cast<long long>(b) * C * hw_q;
  const float* sbp = s4 + static_cast<long long>(sb) * C * hw_s;
  const float* mb = smask + static_cast<long long>(sb) * hw_s;

  if (tid < BM) nqs[tid] = i0 + tid < hw_q ? __ldg(nq + static_cast<long long>(b) * hw_q + i0 + tid) : 0.f;
  float best[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) best[r] = -INFINITY;

  for (int j0 = 0; j0 < hw_s; j0 += BN) {
    __syncthreads();
    if (tid < BN) {
      bool ok = j0 + tid < hw_s;
      nss[tid] = ok ? __ldg(ns + static_cast<long long>(sb) * hw_s + j0 + tid) : 0.f;
      ms[tid] = ok ? __ldg(mb + j0 + tid) : 0.f;
    }
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int s = 0; s < 8; ++s) acc[r][s] = 0.f;
    __syncthreads();
    for (int k0 = 0; k0 < C; k0 += BK) {
      // cooperative tile loads: 8 x 128 floats each, 4 per thread, coalesced along the pixel index
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int idx = tid + e * kGemmThreads;
        int k = idx >> 7, col = idx & 127;
        bool kk = k0 + k < C;
        As[k][col] = (kk && i0 + col < hw_q) ? __ldg(qb + static_cast<long long>(k0 + k) * hw_q + i0 + col) : 0.f;
        Bs[k][col] = (kk && j0 + col < hw_s) ? __ldg(sbp + static_cast<long long>(k0 + k) * hw_s + j0 + col) * ms[col] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 4 + 64]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4 + 64]);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int s = 0; s < 8; ++s) acc[r][s] = fmaf(a[r], bb[s], acc[r][s]);
      }
      __syncthreads();
    }
    // epilogue of this support tile: cosine normalisation and running max over j
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int il = ty * 4 + (r & 3) + (r >> 2) * 64;
      const float nqi = nqs[il];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int jl = tx * 4 + (s & 3) + (s >> 2) * 64;
        if (j0 + jl < hw_s) best[r] = fmaxf(best[r], acc[r][s] / (nss[jl] * nqi + kEps));
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) red[ty * 4 + (r & 3) + (r >> 2) * 64][tx] = best[r];
  __syncthreads();
  if (tid < BM && i0 + tid < hw_q) {
    float m = red[tid][0];
#pragma unroll
    for (int t = 1; t < 16; ++t) m = fmaxf(m, red[tid][t]);
    rowmax[static_cast<long long>(sb) * hw_q + i0 + tid] = m;
  }
}

// ---- tail: min-max normalise each shot over the query pixels, then mean over shots ----------------------
// grid = B, one CTA per batch element.  (pfenet.py:223-229)
__global__ void prior_tail_kernel(const float* __restrict__ rowmax, int B, int S, int hw_q, float* __restrict__ prior) {
  __shared__ float lo_s[32], hi_s[32];
  __shared__ float lo_b, hi_b;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int s = 0; s < S; ++s) {
    const float* r = rowmax + (static_cast<long long>(s) * B + b) * hw_q;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = tid; i < hw_q; i += blockDim.x) {
      float v = r[i];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(kFull, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(kFull, hi, o));
    }
    if (lane == 0) { lo_s[warp] = lo; hi_s[warp] = hi; }
    __syncthreads();
    if (tid == 0) {
      float l = lo_s[0], h = hi_s[0];
      for (int wv = 1; wv < static_cast<int>(blockDim.x >> 5); ++wv) { l = fminf(l, lo_s[wv]); h = fmaxf(h, hi_s[wv]); }
      lo_b = l;
      hi_b = h;
    }
    __syncthreads();
    const float l = lo_b, d = hi_b - lo_b + kEps;
    for (int i = tid; i < hw_q; i += blockDim.x) {
      float v = (r[i] - l) / d;
      float* o = prior + static_cast<long long>(b) * hw_q + i;
      *o = s == 0 ? v : *o + v;
    }
    __syncthreads();
  }
  const float fs = static_cast<float>(S);
  for (int i = tid; i < hw_q; i += blockDim.x) prior[static_cast<long long>(b) * hw_q + i] /= fs;
}

struct Plan {
  size_t off_nq, off_ns, off_rowmax, off_tc, total;
};
Plan make_plan(int B, int S, int C, int hw_s, int hw_q, int precision) {
  Plan p;
  p.off_nq = 0;
  p.off_ns = align_up(static_cast<size_t>(B) * hw_q * sizeof(float), 256);
  p.off_rowmax = p.off_ns + align_up(static_cast<size_t>(S) * B * hw_s * sizeof(float), 256);
  p.off_tc = p.off_rowmax + align_up(static_cast<size_t>(S) * B * hw_q * sizeof(float), 256);
  p.total = p.off_tc + (precision == 1 ? 0 : pemp_prior_tc_workspace_bytes(B, S, C, hw_s, hw_q, precision));
  return p;
}

}  // namespace

extern "C" size_t pemp_prior_mask_workspace_bytes(int B, int S, int C, int hw_s, int hw_q, int precision) {
  if (B <= 0 || S <= 0 || C <= 0 || hw_s <= 0 || hw_q <= 0 || precision < 0 || precision > 2) return 0;
  return make_plan(B, S, C, hw_s, hw_q, precision).total;
}

extern "C" int pemp_prior_mask(const float* q4, const float* s4, const float* smask, int B, int S, int C, int hw_s,
                               int hw_q, int precision, float* prior, float* rowmax_out, void* workspace,
                               size_t workspace_bytes, pemp_stream_t stream) {
  PEMP_REQUIRE(q4 && s4 && smask && prior, PEMP_E_NULL);
  PEMP_REQUIRE(B > 0 && S > 0 && C > 0 && hw_s > 0 && hw_q > 0 && precision >= 0 && precision <= 2, PEMP_E_SHAPE);
  PEMP_REQUIRE(static_cast<long long>(S) * B <= 65535, PEMP_E_SHAPE);
  Plan pl = make_plan(B, S, C, hw_s, hw_q, precision);
  PEMP_REQUIRE(workspace && workspace_bytes >= pl.total, PEMP_E_WORKSPACE);
  char* ws = static_cast<char*>(workspace);
  float* nq = reinterpret_cast<float*>(ws + pl.off_nq);
  float* ns = reinterpret_cast<float*>(ws + pl.off_ns);
  float* rowmax = rowmax_out ? rowmax_out : reinterpret_cast<float*>(ws + pl.off_rowmax);
  cudaStream_t st = as_stream(stream);

  if (precision == 1) {
    col_norm_kernel<<<dim3((hw_q + 31) / 32, B), dim3(32, 8), 0, st>>>(q4, nullptr, C, hw_q, nq);
    col_norm_kernel<<<dim3((hw_s + 31) / 32, S * B), dim3(32, 8), 0, st>>>(s4, smask, C, hw_s, ns);
    prior_fp32_kernel<<<dim3((hw_q + BM - 1) / BM, S * B), kGemmThreads, 0, st>>>(q4, s4, smask, nq, ns, B, C, hw_s, hw_q,
                                                                                rowmax);
    prior_tail_kernel<<<B, 1024, 0, st>>>(rowmax, B, S, hw_q, prior);
    return launch_status();
  }
  // tensor-core paths: memset nodes + pre-pass + tcgen05 GEMM (with the min / max of the row maxima) + parallel tail
  return pemp_prior_tc_launch(q4, s4, smask, nq, ns, B, S, C, hw_s, hw_q, precision, rowmax, prior, ws + pl.off_tc,
                              workspace_bytes - pl.off_tc, st);
}
