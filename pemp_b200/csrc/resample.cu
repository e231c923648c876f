// Resampling kernels: K0 nearest mask down-sampling, K4 bilinear up-sampling + argmax, K5 nearest label
// up-sampling, the stand-alone bilinear resize and the adjoint (transposed) bilinear operator of K6.
// All are tiny next to the feature streams (K1-K3); they are written for coalesced stores.
#include "common.cuh"
#include "iou_count.cuh"

// ------------------------------------------------------------------------------------------------ K0
// in [planes, H, W] -> out [planes, h, w]; one thread per output element.  Reads are a strided gather
// (every ~8th float of every ~8th row): 51 of 401 rows are touched, ~13 % of the mask bytes.
// grid (planes, ceil(h*w / 256)): 32-bit index arithmetic only (the first version decomposed a 64-bit flat index with three
// 64-bit divisions per element - 26 us for 1.66 M outputs, longer than the matching kernel that follows it).
__global__ void __launch_bounds__(256)
mask_nearest_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int h, int w, float sy, float sx) {
  const int i = blockIdx.y * 256 + threadIdx.x;
  if (i >= h * w) return;
  const int y = i / w, x = i - y * w;
  const long long pl = blockIdx.x;
  out[pl * h * w + i] = __ldg(in + (pl * H + nearest_src(y, sy, H)) * W + nearest_src(x, sx, W));
}

extern "C" int pemp_mask_nearest(const float* in, int planes, int H, int W, int h, int w, float* out,
                                 pemp_stream_t stream) {
  PEMP_REQUIRE(in && out, PEMP_E_NULL);
  PEMP_REQUIRE(planes > 0 && H > 0 && W > 0 && h > 0 && w > 0 && static_cast<long long>(h) * w <= 65535LL * 256, PEMP_E_SHAPE);
  // ATen computes the scale as float(in) / out (UpSample.h compute_scales_value with no user scale)
  float sy = static_cast<float>(H) / static_cast<float>(h), sx = static_cast<float>(W) / static_cast<float>(w);
  mask_nearest_kernel<<<dim3(planes, (h * w + 255) / 256), 256, 0, as_stream(stream)>>>(in, out, H, W, h, w, sy, sx);
  return launch_status();
}

// K0 on the label map the data set stores (one uint8 plane, 1 = object, 0 = background, 255 = boundary / ignore) instead of
// the two float planes `stack(fg, bg)` the loader expands it to (data_kits/pascal_voc.py:209-210, 226-231): identical low-res
// masks, 1 byte per pixel on the host link instead of 8.  labels [planes, H, W] -> out [planes, 2, h, w] (fg, bg).
__global__ void __launch_bounds__(256)
mask_nearest_labels_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int H, int W, int h, int w, float sy, float sx) {
  const int i = blockIdx.y * 256 + threadIdx.x;
  if (i >= h * w) return;
  const int y = i / w, x = i - y * w;
  const long long pl = blockIdx.x;
  const uint8_t v = __ldg(in + (pl * H + nearest_src(y, sy, H)) * W + nearest_src(x, sx, W));
  float* o = out + pl * 2 * h * w + i;
  o[0] = v == 1 ? 1.f : 0.f;
  o[h * w] = v == 0 ? 1.f : 0.f;
}

extern "C" int pemp_mask_nearest_labels(const uint8_t* labels, int planes, int H, int W, int h, int w, float* out,
                                        pemp_stream_t stream) {
  PEMP_REQUIRE(labels && out, PEMP_E_NULL);
  PEMP_REQUIRE(planes > 0 && H > 0 && W > 0 && h > 0 && w > 0 && static_cast<long long>(h) * w <= 65535LL * 256, PEMP_E_SHAPE);
  float sy = static_cast<float>(H) / static_cast<float>(h), sx = static_cast<float>(W) / static_cast<float>(w);
  mask_nearest_labels_kernel<<<dim3(planes, (h * w + 255) / 256), 256, 0, as_stream(stream)>>>(labels, out, H, W, h, w, sy, sx);
  return launch_status();
}

// ------------------------------------------------------------------------------------------------ K5
__global__ void nearest_i64_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out, int planes, int h, int w,
                                   int H, int W, float sy, float sx) {
  long long total = static_cast<long long>(planes) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int X = static_cast<int>(i % W);
    long long t = i / W;
    int Y = static_cast<int>(t % H);
    long long pl = t / H;
    out[i] = __ldg(in + (pl * h + nearest_src(Y, sy, h)) * w + nearest_src(X, sx, w));
  }
}

extern "C" int pemp_nearest_resize_i64(const int64_t* in, int planes, int h, int w, int H, int W, int64_t* out,
                                       pemp_stream_t stream) {
  PEMP_REQUIRE(in && out, PEMP_E_NULL);
  PEMP_REQUIRE(planes > 0 && H > 0 && W > 0 && h > 0 && w > 0, PEMP_E_SHAPE);
  long long total = static_cast<long long>(planes) * H * W;
  int block = 256;
  int grid = static_cast<int>(llmin((total + block - 1) / block, 148LL * 16));
  float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
  nearest_i64_kernel<<<grid, block, 0, as_stream(stream)>>>(in, out, planes, h, w, H, W, sy, sx);
  return launch_status();
}

// ------------------------------------------------------------------------------------------------ K4
// pred [N, 2, h, w] -> logits [N, 2, H, W] (optional), mask8 / mask64 [N, H, W] (optional).
// One thread produces 4 consecutive outputs of the flattened [N*H*W] index space so the uint8 mask is
// written with aligned 32-bit stores whatever W is (401 is odd).  The low-res map (2*h*w floats, 20.8 KB
// at 51x51) stays in L1/L2; the kernel is bound by the mask / logits store.
template <bool kLogits, bool kMask8, bool kMask64>
__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ pred, float* __restrict__ logits, uint8_t* __restrict__ mask8,
                       int64_t* __restrict__ mask64, long long total, int h, int w, int H, int W, float sy, float sx) {
  const int HW = H * W;
  const int hw = h * w;
  const long long nquads = (total + 3) >> 2;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < nquads;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    // one 64-bit and one 32-bit division per four outputs; (Y, X) then advance incrementally
    const long long i0 = q << 2;
    int n = static_cast<int>(i0 / HW);
    const int r0 = static_cast<int>(i0 - static_cast<long long>(n) * HW);
    int Y = r0 / W, X = r0 - Y * W;
    const float* p0 = pred + static_cast<long long>(n) * 2 * hw;
    Lerp ly = lerp_coeff(Y, sy, h);
    const float* rt = p0 + ly.i0 * w;          // top / bottom source rows of channel 0 (channel 1 is +hw)
    const float* rb = p0 + ly.i1 * w;
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (i0 + e < total) {
        const Lerp lx = lerp_coeff(X, sx, w);
        float v[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const float a = __ldg(rt + ch * hw + lx.i0), b = __ldg(rt + ch * hw + lx.i1);
          const float c = __ldg(rb + ch * hw + lx.i0), d = __ldg(rb + ch * hw + lx.i1);
          v[ch] = lerp2(ly.l0, lerp2(lx.l0, a, lx.l1, b), ly.l1, lerp2(lx.l0, c, lx.l1, d));
        }
        const uint32_t m = v[1] > v[0] ? 1u : 0u;   // first index wins ties => background
        packed |= m << (8 * e);
        if (kLogits) {
          const long long o = static_cast<long long>(n) * 2 * HW + static_cast<long long>(Y) * W + X;
          logits[o] = v[0];
          logits[o + HW] = v[1];
        }
        if (kMask64) mask64[i0 + e] = static_cast<int64_t>(m);
        if (++X == W) {                             // next output row (never crosses an image inside a quad
          X = 0;                                    // unless H*W % 4 != 0: then (Y == H) wraps to the next image)
          if (++Y == H) {
            Y = 0;
            ++n;
            p0 += 2 * hw;
          }
          ly = lerp_coeff(Y, sy, h);
          rt = p0 + ly.i0 * w;
          rb = p0 + ly.i1 * w;
        }
      }
    }
    if (kMask8) {
      if (i0 + 3 < total) {
        reinterpret_cast<uint32_t*>(mask8)[q] = packed;
      } else {
        for (int e = 0; i0 + e < total; ++e) mask8[i0 + e] = static_cast<uint8_t>(packed >> (8 * e));
      }
    }
  }
}

// K4 fast path (uint8 mask only - the evaluator's case).  ATen evaluates the bilinear sample as
//   v = fma(ly.l0, Hrow(i0, X), ly.l1 * Hrow(i1, X)),   Hrow(y, X) = fma(lx.l0, p[y][x0], lx.l1 * p[y][x1]),
// so the horizontal pass of a source row can be computed ONCE and shared by every output row that uses it (8 output
// rows per source row at 51 -> 401) without changing a single bit.  A CTA owns a band of kBandRows output rows of one
// image: it first fills shared memory with Hrow for the few source rows the band touches (both channels), then each
// thread produces aligned quads of the flattened output range of the band: 4 shared loads, 2 lerps and a compare per
// pixel instead of 8 cached global loads and 6 lerps.  Quads that straddle the band boundary are written per byte.
#ifndef PEMP_K4_BAND
#define PEMP_K4_BAND 16
#endif
constexpr int kBandRows = PEMP_K4_BAND, kBandMaxSrc = 8;
__host__ __device__ inline int band_pitch(int W) { return W + (W >> 4) + 1; }   // float2 elements per staged source row
// kHist: the FewShotMetric counts of K10 (core/metrics.py:9-23) are taken from the mask bytes while they are still in
// registers - one launch and one read of the mask fewer (pemp_upsample_argmax_hist).
template <bool kHist>
__global__ void __launch_bounds__(256)
upsample_argmax_band_kernel(const float* __restrict__ pred, uint8_t* __restrict__ mask8, int h, int w, int H, int W,
                            float sy, float sx, int bands, const uint8_t* __restrict__ ref, const int64_t* __restrict__ cls,
                            int num_classes, unsigned long long* __restrict__ stat) {
  // hrow [nsrc][Wp]{ch 0, ch 1}: the two classes of a pixel side by side, so a sample is two 8-byte loads (top and bottom
  // source row) instead of four 4-byte ones; pixel X sits at X + (X >> 4) (Wp = band_pitch(W)): the lanes of a warp read
  // pixels 4 apart (a lane owns a quad), which is a 4-way bank conflict in a dense row and conflict-free with one pad element
  // per 16.  vrow [kBandRows]: the vertical coefficients of the band's output rows, computed once.  (Round 2: ncu counted
  // 18.6 thread instructions per pixel, most of them index arithmetic - a 64-bit division per quad, the row coefficients
  // per quad, bounds and wrap tests per pixel - and 3.7 M of 5.1 M shared wavefronts were conflicts of the quad reads.)
  extern __shared__ float hrow[];
  __shared__ float4 vrow[kBandRows];                 // { (i0 - src0) * Wp, (i1 - src0) * Wp (as int bits), l0, l1 }
  const int Wp = band_pitch(W);
  const int n = blockIdx.x / bands, band = blockIdx.x - n * bands;
  const int Y0 = band * kBandRows, Y1 = min(H, Y0 + kBandRows);
  const int src0 = lerp_coeff(Y0, sy, h).i0, nsrc = lerp_coeff(Y1 - 1, sy, h).i1 - src0 + 1;
  const int hw = h * w;
  const float* p0 = pred + static_cast<long long>(n) * 2 * hw + src0 * w;
  for (int X = threadIdx.x; X < W; X += blockDim.x) {
    const Lerp lx = lerp_coeff(X, sx, w);
    for (int r = 0; r < nsrc; ++r) {
      const float* row = p0 + r * w;
      const float a = lerp2(lx.l0, __ldg(row + lx.i0), lx.l1, __ldg(row + lx.i1));
      const float c = lerp2(lx.l0, __ldg(row + hw + lx.i0), lx.l1, __ldg(row + hw + lx.i1));
      reinterpret_cast<float2*>(hrow)[r * Wp + X + (X >> 4)] = make_float2(a, c);
    }
  }
  if (threadIdx.x < kBandRows) {
    const Lerp ly = lerp_coeff(min(Y0 + static_cast<int>(threadIdx.x), H - 1), sy, h);
    vrow[threadIdx.x] = make_float4(__int_as_float((ly.i0 - src0) * Wp), __int_as_float((ly.i1 - src0) * Wp), ly.l0, ly.l1);
  }
  __syncthreads();
  const long long HW = static_cast<long long>(H) * W;
  const long long base0 = n * HW + static_cast<long long>(Y0) * W, base1 = n * HW + static_cast<long long>(Y1) * W;
  const float2* hr = reinterpret_cast<const float2*>(hrow);
  PempCounts cnt = {0, 0, 0, 0, 0, 0};
  // one sample: v_c = fma(l0, top_c, l1 * bottom_c) for the two classes, class 1 wins only if strictly larger
  auto sample = [&](const float4 vy, int X) -> uint32_t {
    const int xp = X + (X >> 4);
    const float2 t = hr[__float_as_int(vy.x) + xp], bt = hr[__float_as_int(vy.y) + xp];
    const float v0 = lerp2(vy.z, t.x, vy.w, bt.x);
    const float v1 = lerp2(vy.z, t.y, vy.w, bt.y);
    return v1 > v0 ? 1u : 0u;
  };
  // quads of the flattened output (aligned 32-bit stores although W is odd); a thread's quads are 4 * blockDim apart, so
  // its (row, column) advances by constants instead of being re-derived with a division
  const long long q0 = (base0 >> 2) + threadIdx.x;
  const int stepY = static_cast<int>(4 * blockDim.x) / W, stepX = static_cast<int>(4 * blockDim.x) - stepY * W;
  int rel = static_cast<int>((q0 << 2) - base0);     // offset of the quad's first element in the band (may be < 0 for the first quad)
  int Y = rel >= 0 ? rel / W : 0, X = rel >= 0 ? rel - Y * W : rel;
  const int band_len = static_cast<int>(base1 - base0);
  for (long long q = q0; rel < band_len; q += blockDim.x, rel += 4 * blockDim.x) {
    uint32_t packed = 0;
    if (rel >= 0 && rel + 3 < band_len && X + 3 < W) {         // whole quad inside the band and inside one row
      const float4 vy = vrow[Y];
#pragma unroll
      for (int e = 0; e < 4; ++e) packed |= sample(vy, X + e) << (8 * e);
      reinterpret_cast<uint32_t*>(mask8)[q] = packed;
      if (kHist) pemp_count_word(packed, __ldg(reinterpret_cast<const uint32_t*>(ref) + q), cnt);
    } else {                                                   // band boundary or a quad that wraps to the next row
      int yy = Y, xx = rel < 0 ? 0 : X;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = rel + e;
        if (r >= 0 && r < band_len) packed |= sample(vrow[yy], xx) << (8 * e);
        if (r >= 0 && ++xx == W) {
          xx = 0;
          ++yy;
        }
      }
      if (rel >= 0 && rel + 3 < band_len) {
        reinterpret_cast<uint32_t*>(mask8)[q] = packed;
        if (kHist) pemp_count_word(packed, __ldg(reinterpret_cast<const uint32_t*>(ref) + q), cnt);
      } else {
        const long long i0 = q << 2;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (rel + e >= 0 && rel + e < band_len) {
            mask8[i0 + e] = static_cast<uint8_t>(packed >> (8 * e));
            if (kHist) pemp_count_byte(static_cast<uint8_t>(packed >> (8 * e)), __ldg(ref + i0 + e), cnt);
          }
      }
    }
    X += stepX;
    Y += stepY;
    if (X >= W) {
      X -= W;
      ++Y;
    } else if (X < 0) {                                        // (only after a first quad that starts before the band)
      X += W;
      --Y;
    }
  }
  if (kHist) {   // block reduction, then six integer atomics (order independent => exact)
    __shared__ unsigned red[6][8];
    unsigned v[6] = {cnt.tp0, cnt.fp0, cnt.fn0, cnt.tp1, cnt.fp1, cnt.fn1};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      unsigned sres = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sres += __shfl_xor_sync(kFull, sres, o);
      if (lane == 0) red[k][warp] = sres;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      unsigned long long tot = 0;
      for (int wv = 0; wv < 8; ++wv) tot += red[threadIdx.x][wv];
      tot >>= 3;
      if (tot) {
        const int k = threadIdx.x;
        const long long row = k < 3 ? 0 : cls[n];
        if (row >= 0 && row <= num_classes) atomicAdd(stat + row * 3 + (k % 3), tot);
      }
    }
  }
}

extern "C" int pemp_upsample_argmax(const float* pred, int N, int h, int w, int H, int W, float* logits, uint8_t* mask8,
                                    int64_t* mask64, pemp_stream_t stream) {
  PEMP_REQUIRE(pred, PEMP_E_NULL);
  PEMP_REQUIRE(logits || mask8 || mask64, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && H > 0 && W > 0 && h > 0 && w > 0, PEMP_E_SHAPE);
  PEMP_REQUIRE((reinterpret_cast<uintptr_t>(mask8) & 3) == 0, PEMP_E_ALIGN);
  long long total = static_cast<long long>(N) * H * W;
  int block = 256;
  int grid = static_cast<int>(llmin((total / 4 + block) / block, 148LL * 32));
  float sy = lerp_scale(h, H), sx = lerp_scale(w, W);
  cudaStream_t st = as_stream(stream);
  if (mask8 && !logits && !mask64) {
    // source rows one band can touch: floor((kBandRows - 1) * sy) + 3 (top row, its partner, rounding)
    const int nsrc_max = static_cast<int>((kBandRows - 1) * sy) + 3;
    const size_t smem = static_cast<size_t>(nsrc_max < h ? nsrc_max : h) * 2 * band_pitch(W) * sizeof(float);
    if (nsrc_max <= kBandMaxSrc && smem <= 48 * 1024) {
      const int bands = (H + kBandRows - 1) / kBandRows;
      upsample_argmax_band_kernel<false><<<static_cast<unsigned>(N) * bands, 256, smem, st>>>(pred, mask8, h, w, H, W, sy, sx, bands,
                                                                                               nullptr, nullptr, 0, nullptr);
      return launch_status();
    }
  }
#define PEMP_LAUNCH_UA(L, M8, M64)                                                                              \
  upsample_argmax_kernel<L, M8, M64><<<grid, block, 0, st>>>(pred, logits, mask8, mask64, total, h, w, H, W, sy, sx)
  int sel = (logits ? 4 : 0) | (mask8 ? 2 : 0) | (mask64 ? 1 : 0);
  switch (sel) {
    case 1: PEMP_LAUNCH_UA(false, false, true); break;
    case 2: PEMP_LAUNCH_UA(false, true, false); break;
    case 3: PEMP_LAUNCH_UA(false, true, true); break;
    case 4: PEMP_LAUNCH_UA(true, false, false); break;
    case 5: PEMP_LAUNCH_UA(true, false, true); break;
    case 6: PEMP_LAUNCH_UA(true, true, false); break;
    default: PEMP_LAUNCH_UA(true, true, true); break;
  }
#undef PEMP_LAUNCH_UA
  return launch_status();
}

// K4 + K10 in one launch for the evaluator: uint8 masks and the FewShotMetric counts of (mask, ref, cls).
// Falls back to the two separate kernels when the banded kernel does not cover the geometry.
extern "C" int pemp_upsample_argmax_hist(const float* pred, int N, int h, int w, int H, int W, uint8_t* mask8,
                                         const uint8_t* ref, const int64_t* cls, int num_classes, int64_t* stat,
                                         pemp_stream_t stream) {
  PEMP_REQUIRE(pred && mask8 && ref && cls && stat, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && h > 0 && w > 0 && num_classes > 0, PEMP_E_SHAPE);
  PEMP_REQUIRE((reinterpret_cast<uintptr_t>(mask8) & 3) == 0 && (reinterpret_cast<uintptr_t>(ref) & 3) == 0, PEMP_E_ALIGN);
  const float sy = lerp_scale(h, H), sx = lerp_scale(w, W);
  const int nsrc_max = static_cast<int>((kBandRows - 1) * sy) + 3;
  const size_t smem = static_cast<size_t>(nsrc_max < h ? nsrc_max : h) * 2 * band_pitch(W) * sizeof(float);
  if (nsrc_max <= kBandMaxSrc && smem <= 48 * 1024) {
    const int bands = (H + kBandRows - 1) / kBandRows;
    upsample_argmax_band_kernel<true><<<static_cast<unsigned>(N) * bands, 256, smem, as_stream(stream)>>>(
        pred, mask8, h, w, H, W, sy, sx, bands, ref, cls, num_classes, reinterpret_cast<unsigned long long*>(stat));
    return launch_status();
  }
  int rc = pemp_upsample_argmax(pred, N, h, w, H, W, nullptr, mask8, nullptr, stream);
  if (rc != PEMP_OK) return rc;
  return pemp_iou_hist(mask8, ref, cls, N, static_cast<long long>(H) * W, num_classes, stat, stream);
}

// single-plane bilinear resize with the same arithmetic (PFENet's mask resize, pfenet.py:191,205)
__global__ void bilinear_resize_kernel(const float* __restrict__ in, float* __restrict__ out, long long total, int h,
                                       int w, int H, int W, float sy, float sx) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int X = static_cast<int>(i % W);
    long long t = i / W;
    int Y = static_cast<int>(t % H);
    const float* p = in + (t / H) * h * w;
    Lerp ly = lerp_coeff(Y, sy, h), lx = lerp_coeff(X, sx, w);
    float a = __ldg(p + ly.i0 * w + lx.i0), b = __ldg(p + ly.i0 * w + lx.i1);
    float c = __ldg(p + ly.i1 * w + lx.i0), d = __ldg(p + ly.i1 * w + lx.i1);
    out[i] = lerp2(ly.l0, lerp2(lx.l0, a, lx.l1, b), ly.l1, lerp2(lx.l0, c, lx.l1, d));
  }
}

extern "C" int pemp_bilinear_resize(const float* in, int planes, int h, int w, int H, int W, float* out,
                                    pemp_stream_t stream) {
  PEMP_REQUIRE(in && out, PEMP_E_NULL);
  PEMP_REQUIRE(planes > 0 && H > 0 && W > 0 && h > 0 && w > 0, PEMP_E_SHAPE);
  long long total = static_cast<long long>(planes) * H * W;
  int block = 256;
  int grid = static_cast<int>(llmin((total + block - 1) / block, 148LL * 32));
  bilinear_resize_kernel<<<grid, block, 0, as_stream(stream)>>>(in, out, total, h, w, H, W, lerp_scale(h, H),
                                                                lerp_scale(w, W));
  return launch_status();
}

// ------------------------------------------------------------------------------------------- K6 adjoint
// wt[y, x] = sum_{Y, X} m[Y, X] * a_y(Y) * b_x(X), where a_y(Y) is the weight the forward bilinear
// operator (h -> H, align_corners) gives source row y for output row Y.  With it
//   sum_{YX} m * (U f) == sum_{yx} f * wt      (baseline.py:100-110 without the 329 MB/shot up-sampled copy).
// One CTA per (plane, low-res row y): stage 1 reduces the <= 2/scale contributing mask rows into a
// weighted row r[X] in shared memory (coalesced reads of the mask), stage 2 applies the column weights.
// msum[plane] (optional) accumulates the plain mask sum (the reference's denominator, exact for 0/1 masks).
__global__ void bilinear_adjoint_kernel(const float* __restrict__ mask, float* __restrict__ wt, float* __restrict__ msum,
                                        int H, int W, int h, int w, float sy, float sx) {
  extern __shared__ float row[];   // [W]
  const int y = blockIdx.x, pl = blockIdx.y;
  const float* m = mask + static_cast<long long>(pl) * H * W;
  // rows Y whose i0 or i1 equals y lie in [ (y-1)/sy, (y+1)/sy ]; widen by one for rounding
  int Ylo = 0, Yhi = H - 1;
  if (sy > 0.f) {
    Ylo = max(0, static_cast<int>(floorf((y - 1) / sy)) - 1);
    Yhi = min(H - 1, static_cast<int>(ceilf((y + 1) / sy)) + 1);
  }
  for (int X = threadIdx.x; X < W; X += blockDim.x) {
    float acc = 0.f;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
      Lerp l = lerp_coeff(Y, sy, h);
      float wgt = (l.i0 == y ? l.l0 : 0.f) + (l.i1 == y ? l.l1 : 0.f);
      if (wgt != 0.f) acc = fmaf(wgt, __ldg(m + static_cast<long long>(Y) * W + X), acc);
    }
    row[X] = acc;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    int Xlo = 0, Xhi = W - 1;
    if (sx > 0.f) {
      Xlo = max(0, static_cast<int>(floorf((x - 1) / sx)) - 1);
      Xhi = min(W - 1, static_cast<int>(ceilf((x + 1) / sx)) + 1);
    }
    float acc = 0.f;
    for (int X = Xlo; X <= Xhi; ++X) {
      Lerp l = lerp_coeff(X, sx, w);
      float wgt = (l.i0 == x ? l.l0 : 0.f) + (l.i1 == x ? l.l1 : 0.f);
      if (wgt != 0.f) acc = fmaf(wgt, row[X], acc);
    }
    wt[(static_cast<long long>(pl) * h + y) * w + x] = acc;
  }
}

// plain per-plane sum (deterministic: one CTA per plane, fixed reduction tree)
__global__ void plane_sum_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  __shared__ float part[32];
  const float* p = in + blockIdx.x * n;
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += __ldg(p + i);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
  }
}

// ---- K6 adjoint, single-read version (used by pemp_map_pool_fullres, which owns a workspace) -------------------
// Every mask row Y belongs to exactly one low-res row i0(Y); it contributes l0(Y) to row i0 and l1(Y) to row i1 (= i0
// or i0 + 1).  CTA (y, plane) reads ITS mask rows once (the kernel above reads every row twice and the mask sum a third
// time): thread <-> columns X, X + 256 with all row loads independent, two weighted rows A (-> y) and B (-> y + 1) in
// shared memory, then the column pass with per-column coefficients computed once; it writes a[y][x], b[y][x] and the
// plain sum of its rows.  `adjoint_combine_kernel` forms wt[y] = a[y] + b[y-1] and adds the h row sums in index order
// (deterministic).  Measured: 222 us for 640 planes of 401 x 401 (1.9 TB/s; 51 x 640 small CTAs, bound by the per-CTA
// dependency chain - a version with 4 low-res rows per CTA was slower); the generic kernel above needs 390 us.
constexpr int kAdjMaxRows = 32;    // mask rows one low-res row can own in the fast kernel (else the generic kernel)
__global__ void __launch_bounds__(256)
adjoint_rows_kernel(const float* __restrict__ mask, float* __restrict__ ab, float* __restrict__ rsum, int H, int W,
                    int h, int w, float sy, float sx) {
  extern __shared__ float sh[];      // A[W], B[W], l0[W], l1[W], i0[W] (as int)
  float* A = sh;
  float* Bv = sh + W;
  float* cl0 = sh + 2 * W;
  float* cl1 = sh + 3 * W;
  int* ci0 = reinterpret_cast<int*>(sh + 4 * W);
  __shared__ float wa[kAdjMaxRows], wb[kAdjMaxRows], red[8];
  __shared__ int range[2];
  const int y = blockIdx.x, pl = blockIdx.y;
  const float* m = mask + static_cast<long long>(pl) * H * W;
  if (threadIdx.x == 0) {
    // mask rows with i0(Y) == y form a contiguous range inside [y/sy - 1, (y+1)/sy + 1]
    int Ylo = 0, Yhi = H - 1;
    if (sy > 0.f) {
      Ylo = max(0, static_cast<int>(floorf(y / sy)) - 1);
      Yhi = min(H - 1, static_cast<int>(ceilf((y + 1) / sy)) + 1);
    }
    while (Ylo <= Yhi && lerp_coeff(Ylo, sy, h).i0 != y) ++Ylo;
    while (Yhi >= Ylo && lerp_coeff(Yhi, sy, h).i0 != y) --Yhi;
    range[0] = Ylo;
    range[1] = Yhi - Ylo + 1;
  }
  __syncthreads();
  const int Ylo = range[0], nr = range[1];
  if (threadIdx.x < nr) {            // row weights: l0 -> row y, l1 -> row i1 (y or y + 1)
    const Lerp l = lerp_coeff(Ylo + threadIdx.x, sy, h);
    wa[threadIdx.x] = l.l0 + (l.i1 == y ? l.l1 : 0.f);
    wb[threadIdx.x] = l.i1 == y ? 0.f : l.l1;
  }
  __syncthreads();
  float s = 0.f;
  for (int X = threadIdx.x; X < W; X += blockDim.x) {
    const Lerp lx = lerp_coeff(X, sx, w);
    cl0[X] = lx.l0;
    cl1[X] = lx.i1 != lx.i0 ? lx.l1 : 0.f;                 // i1 == i0 only at the last column, where l1 == 0 anyway
    ci0[X] = lx.i0;
    float a = 0.f, b = 0.f;
    const float* mc = m + static_cast<long long>(Ylo) * W + X;
#pragma unroll 8
    for (int r = 0; r < nr; ++r) {
      const float v = __ldg(mc + static_cast<long long>(r) * W);
      s += v;
      a = fmaf(wa[r], v, a);
      b = fmaf(wb[r], v, b);
    }
    A[X] = a;
    Bv[X] = b;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    rsum[pl * h + y] = t;
  }
  // column pass: column X feeds x = i0(X) with l0 and x = i0(X) + 1 with l1
  for (int i = threadIdx.x; i < 2 * w; i += blockDim.x) {
    const int which = i / w, x = i - which * w;
    int Xlo = 0, Xhi = W - 1;
    if (sx > 0.f) {
      Xlo = max(0, static_cast<int>(floorf((x - 1) / sx)) - 1);
      Xhi = min(W - 1, static_cast<int>(ceilf((x + 1) / sx)) + 1);
    }
    const float* r = sh + which * W;
    float acc = 0.f;
    for (int X = Xlo; X <= Xhi; ++X) {
      const int i0 = ci0[X];
      const float wgt = i0 == x ? cl0[X] : (i0 + 1 == x ? cl1[X] : 0.f);
      acc = fmaf(wgt, r[X], acc);
    }
    ab[((static_cast<long long>(pl) * h + y) * 2 + which) * w + x] = acc;
  }
}

__global__ void adjoint_combine_kernel(const float* __restrict__ ab, const float* __restrict__ rsum, float* __restrict__ wt,
                                       float* __restrict__ msum, int h, int w) {
  const int pl = blockIdx.x;
  const float* p = ab + static_cast<long long>(pl) * h * 2 * w;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    float v = p[(y * 2 + 0) * w + x];
    if (y > 0) v += p[((y - 1) * 2 + 1) * w + x];
    wt[static_cast<long long>(pl) * h * w + i] = v;
  }
  if (msum && threadIdx.x == 0) {
    float t = 0.f;
    for (int y = 0; y < h; ++y) t += rsum[pl * h + y];
    msum[pl] = t;
  }
}

// ---- K6 adjoint, table-driven version (round 2) ---------------------------------------------------------------
// adjoint_rows_kernel above spends most of a CTA's life before its first mask load: thread 0 walks lerp_coeff to find the row
// range, a barrier, the row weights, a barrier, 401 lerp_coeff evaluations for the column coefficients.  All of that depends
// only on (H, W, h, w): `adjoint_tables_kernel` (one CTA, once per call) writes
//   rowinfo[y] = {first mask row with i0 == y, number of such rows},  wab[y][r] = {weight into row y, weight into row y + 1},
//   xlo[x], xn[x], xw[x][t]: the contiguous run of mask columns that feed low-res column x and their weights,
// and `adjoint_rows2_kernel` starts with its loads: CTA (y, image) reads the mask rows of NP planes (fg and bg of one support
// image: 2 x 2 columns x ~8 rows = 32 independent loads per thread), one barrier, the column pass as a dot product with the
// tap table (no compares), and the row sums.  Same a / b / rsum outputs as before, same combine kernel.
constexpr int kAdjMaxTaps = 64;
struct AdjTables {
  int2* rowinfo;      // [h]
  float2* wab;        // [h][kAdjMaxRows]
  int2* xinfo;        // [w]  {xlo, xn}
  float* xw;          // [kAdjMaxTaps][w]  (tap-major: the lanes of a warp own consecutive x and read consecutive floats)
};
__global__ void adjoint_tables_kernel(AdjTables t, int H, int W, int h, int w, float sy, float sx) {
  for (int y = threadIdx.x; y < h; y += blockDim.x) {
    int Ylo = 0, Yhi = H - 1;
    if (sy > 0.f) {
      Ylo = max(0, static_cast<int>(floorf(y / sy)) - 1);
      Yhi = min(H - 1, static_cast<int>(ceilf((y + 1) / sy)) + 1);
    }
    while (Ylo <= Yhi && lerp_coeff(Ylo, sy, h).i0 != y) ++Ylo;
    while (Yhi >= Ylo && lerp_coeff(Yhi, sy, h).i0 != y) --Yhi;
    const int nr = max(0, Yhi - Ylo + 1);
    t.rowinfo[y] = make_int2(Ylo, nr);
    for (int r = 0; r < nr && r < kAdjMaxRows; ++r) {
      const Lerp l = lerp_coeff(Ylo + r, sy, h);
      t.wab[y * kAdjMaxRows + r] = make_float2(l.l0 + (l.i1 == y ? l.l1 : 0.f), l.i1 == y ? 0.f : l.l1);
    }
  }
  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    int Xlo = 0, Xhi = W - 1;
    if (sx > 0.f) {
      Xlo = max(0, static_cast<int>(floorf((x - 1) / sx)) - 1);
      Xhi = min(W - 1, static_cast<int>(ceilf((x + 1) / sx)) + 1);
    }
    int first = -1, n = 0;
    for (int X = Xlo; X <= Xhi; ++X) {
      const Lerp lx = lerp_coeff(X, sx, w);
      const float wgt = lx.i0 == x ? lx.l0 : ((lx.i0 + 1 == x && lx.i1 != lx.i0) ? lx.l1 : 0.f);
      const bool feeds = lx.i0 == x || (lx.i0 + 1 == x && lx.i1 != lx.i0);
      if (feeds) {
        if (first < 0) first = X;
        if (X - first < kAdjMaxTaps) t.xw[(X - first) * w + x] = wgt;
        n = X - first + 1;
      }
    }
    t.xinfo[x] = make_int2(first < 0 ? 0 : first, n);
  }
}

// shared-memory index of column X: one pad float per 32 columns.  The column pass reads column xlo(x) + k with lane <-> x, i.e.
// at a lane stride of ~W/w (8) floats - an 8-way bank conflict in a plain row (ncu: 10.1 M of 14.6 M wavefronts); with the pad
// lanes 4m + j land on banks m + 8j: conflict-free, and the row pass (lane <-> consecutive X of one 32-aligned group) stays so.
__device__ __forceinline__ int adj_pad(int X) { return X + (X >> 5); }
__host__ __device__ constexpr int adj_row_floats(int W) { return W + (W >> 5) + 1; }

#ifndef PEMP_ADJ_THREADS
#define PEMP_ADJ_THREADS 128
#endif
#ifndef PEMP_ADJ_MINB
#define PEMP_ADJ_MINB 8
#endif
constexpr int kAdjT = PEMP_ADJ_THREADS;
constexpr int kAdjUnroll = 9;      // mask rows per low-res row at the usual 8x geometry: 8, 9 at some rows
// KW > 0: the row pitch is a compile-time constant, so the (row, plane) offsets of a thread's loads are instruction immediates
// (KW = 0 takes W from the argument; ncu of that version: 51 % of all instructions were 64-bit address arithmetic).
// T = float: NP planes of floats per image.  T = uint8_t: ONE label plane per image (1 object / 0 background / 255 boundary, the
// map data_kits/pascal_voc.py:209-210 expands into the float `sup_mask`); a byte yields both planes, fg = (b == 1), bg = (b == 0):
// half the loads and an eighth of the mask bytes (`pemp_map_pool_fullres_labels`).
template <int NP, typename T>
__device__ __forceinline__ void adj_load(const T* p, int PS, float (&v)[NP]) {
  if constexpr (sizeof(T) == 1) {
    static_assert(sizeof(T) != 1 || NP == 2, "a label byte stands for the fg and the bg plane");
    const unsigned b = __ldg(p);
    v[0] = b == 1u ? 1.f : 0.f;
    v[NP - 1] = b == 0u ? 1.f : 0.f;
  } else {
#pragma unroll
    for (int q = 0; q < NP; ++q) v[q] = __ldg(p + q * PS);
  }
}
template <int NP, int KW, typename T>
__global__ void __launch_bounds__(kAdjT, PEMP_ADJ_MINB)
adjoint_rows2_kernel(const T* __restrict__ mask, AdjTables t, float* __restrict__ ab, float* __restrict__ rsum, int H,
                     int W_arg, int h, int w) {
  const int W = KW > 0 ? KW : W_arg;
  extern __shared__ float sh[];      // [NP][2][Wp]: A (-> row y) and B (-> row y + 1) of every plane, padded rows
  __shared__ float red[NP][kAdjT / 32];
  const int Wp = adj_row_floats(W);
  const int y = blockIdx.x, img = blockIdx.y;
  const int2 ri = t.rowinfo[y];
  const int Ylo = ri.x, nr = ri.y;
  const float2* __restrict__ wab = t.wab + y * kAdjMaxRows;
  // 32-bit offsets from one CTA-uniform base: the first version spent 51 % of its instructions (ncu: IMAD / LEA / IADD3, 6.6 per
  // load) on 64-bit address arithmetic of the form ((p * H + r) * W + X)
  constexpr int kPlanesIn = sizeof(T) == 1 ? 1 : NP;       // stored planes per image
  const T* m = mask + (static_cast<long long>(img) * kPlanesIn * H + Ylo) * W;
  const int PS = H * W;
  float s[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) s[p] = 0.f;
  for (int X = threadIdx.x; X < W; X += kAdjT) {
    float a[NP], b[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) a[p] = b[p] = 0.f;
    const T* q = m + X;
    if (KW > 0 && nr <= kAdjUnroll) {
#pragma unroll
      for (int r = 0; r < kAdjUnroll; ++r) {
        if (r < nr) {
          const float2 wr = __ldg(wab + r);
          float v[NP];
          adj_load<NP, T>(q + r * KW, PS, v);
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            s[p] += v[p];
            a[p] = fmaf(wr.x, v[p], a[p]);
            b[p] = fmaf(wr.y, v[p], b[p]);
          }
        }
      }
    } else {
      int off = 0;
#pragma unroll 8
      for (int r = 0; r < nr; ++r, off += W) {
        const float2 wr = __ldg(wab + r);
        float v[NP];
        adj_load<NP, T>(q + off, PS, v);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          s[p] += v[p];
          a[p] = fmaf(wr.x, v[p], a[p]);
          b[p] = fmaf(wr.y, v[p], b[p]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      sh[(p * 2 + 0) * Wp + adj_pad(X)] = a[p];
      sh[(p * 2 + 1) * Wp + adj_pad(X)] = b[p];
    }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    s[p] = warp_sum(s[p]);
    if ((threadIdx.x & 31) == 0) red[p][threadIdx.x >> 5] = s[p];
  }
  __syncthreads();
  if (threadIdx.x < NP) {
    float tsum = 0.f;
    for (int i = 0; i < kAdjT / 32; ++i) tsum += red[threadIdx.x][i];
    rsum[(img * NP + threadIdx.x) * h + y] = tsum;
  }
  // column pass: output (plane p, which in {a, b}, x) = sum_t xw[x][t] * row[xlo[x] + t]
  for (int i = threadIdx.x; i < NP * 2 * w; i += kAdjT) {
    const int pw = i / w, x = i - pw * w;             // pw = p * 2 + which
    const int2 xi = __ldg(t.xinfo + x);
    const float* r = sh + pw * Wp;
    const float* wx = t.xw + x;
    float acc = 0.f;
    int X = xi.x;
    for (int k = 0; k < xi.y; ++k, ++X, wx += w) acc = fmaf(__ldg(wx), r[X + (X >> 5)], acc);
    const int p = pw >> 1, which = pw & 1;
    ab[((static_cast<long long>(img * NP + p) * h + y) * 2 + which) * w + x] = acc;
  }
}

// instantiations for the mask widths of the reference's data sets (401 PASCAL, 417 COCO / PANet, 473 PFENet, 321); others: KW = 0
template <int NP, typename T>
static void adjoint_rows2_launch(dim3 grid, size_t smem, cudaStream_t st, const T* mask, AdjTables t, float* ab, float* rsum,
                                 int H, int W, int h, int w) {
  switch (W) {
    case 401: adjoint_rows2_kernel<NP, 401, T><<<grid, kAdjT, smem, st>>>(mask, t, ab, rsum, H, W, h, w); break;
    case 417: adjoint_rows2_kernel<NP, 417, T><<<grid, kAdjT, smem, st>>>(mask, t, ab, rsum, H, W, h, w); break;
    case 473: adjoint_rows2_kernel<NP, 473, T><<<grid, kAdjT, smem, st>>>(mask, t, ab, rsum, H, W, h, w); break;
    case 321: adjoint_rows2_kernel<NP, 321, T><<<grid, kAdjT, smem, st>>>(mask, t, ab, rsum, H, W, h, w); break;
    default: adjoint_rows2_kernel<NP, 0, T><<<grid, kAdjT, smem, st>>>(mask, t, ab, rsum, H, W, h, w); break;
  }
}

static size_t adj_tables_bytes(int h, int w) {
  return align_up(static_cast<size_t>(h) * sizeof(int2), 256) + align_up(static_cast<size_t>(h) * kAdjMaxRows * sizeof(float2), 256) +
         align_up(static_cast<size_t>(w) * sizeof(int2), 256) + align_up(static_cast<size_t>(w) * kAdjMaxTaps * sizeof(float), 256);
}
size_t pemp_adjoint_scratch_bytes(int planes, int h, int w) {
  return align_up(static_cast<size_t>(planes) * h * 2 * w * sizeof(float), 256) + align_up(static_cast<size_t>(planes) * h * sizeof(float), 256) +
         adj_tables_bytes(h, w);
}
static bool adj_fast_geometry(int H, int W, int h, int w) {
  // the fast kernels keep a few rows of W floats in shared memory and bound the rows / columns one low-res row / column owns
  return !(5 * W * sizeof(float) > 48 * 1024 || (H > h && (H + h - 1) / h + 2 > kAdjMaxRows) ||
           (W > w && 2 * ((W + w - 1) / w) + 3 > kAdjMaxTaps));
}
struct AdjScratch {
  float* ab;
  float* rsum;
  AdjTables t;
};
static AdjScratch adj_scratch(char* scratch, int planes, int h, int w) {
  AdjScratch a;
  a.ab = reinterpret_cast<float*>(scratch);
  char* p = scratch + align_up(static_cast<size_t>(planes) * h * 2 * w * sizeof(float), 256);
  a.rsum = reinterpret_cast<float*>(p);
  p += align_up(static_cast<size_t>(planes) * h * sizeof(float), 256);
  a.t.rowinfo = reinterpret_cast<int2*>(p);
  p += align_up(static_cast<size_t>(h) * sizeof(int2), 256);
  a.t.wab = reinterpret_cast<float2*>(p);
  p += align_up(static_cast<size_t>(h) * kAdjMaxRows * sizeof(float2), 256);
  a.t.xinfo = reinterpret_cast<int2*>(p);
  p += align_up(static_cast<size_t>(w) * sizeof(int2), 256);
  a.t.xw = reinterpret_cast<float*>(p);
  return a;
}

int pemp_adjoint_launch(const float* mask, int planes, int H, int W, int h, int w, float* wt, float* msum, char* scratch,
                        cudaStream_t st) {
  if (!adj_fast_geometry(H, W, h, w))
    return pemp_bilinear_adjoint(mask, planes, H, W, h, w, wt, msum, reinterpret_cast<pemp_stream_t>(st));
  const AdjScratch a = adj_scratch(scratch, planes, h, w);
  const float sy = lerp_scale(h, H), sx = lerp_scale(w, W);
#ifdef PEMP_ADJ_V1
  adjoint_rows_kernel<<<dim3(h, planes), 256, 5 * W * sizeof(float), st>>>(mask, a.ab, a.rsum, H, W, h, w, sy, sx);
#else
  adjoint_tables_kernel<<<1, 256, 0, st>>>(a.t, H, W, h, w, sy, sx);
#ifdef PEMP_ADJ_NP1
  if (false)
#else
  if (planes % 2 == 0)
#endif
    adjoint_rows2_launch<2, float>(dim3(h, planes / 2), 4 * adj_row_floats(W) * sizeof(float), st, mask, a.t, a.ab, a.rsum, H, W, h, w);
  else
    adjoint_rows2_launch<1, float>(dim3(h, planes), 2 * adj_row_floats(W) * sizeof(float), st, mask, a.t, a.ab, a.rsum, H, W, h, w);
#endif
  adjoint_combine_kernel<<<planes, 256, 0, st>>>(a.ab, a.rsum, wt, msum, h, w);
  return launch_status();
}

// labels [images, H, W] uint8 -> the float planes stack((label == 1), (label == 0)) [images, 2, H, W] (only for geometries
// the fast adjoint does not cover; `expanded` is workspace)
__global__ void labels_expand_kernel(const uint8_t* __restrict__ lab, float* __restrict__ out, long long images, long long HW) {
  const long long total = images * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / HW, r = i - n * HW;
    const uint8_t v = __ldg(lab + i);
    out[(n * 2 + 0) * HW + r] = v == 1 ? 1.f : 0.f;
    out[(n * 2 + 1) * HW + r] = v == 0 ? 1.f : 0.f;
  }
}
size_t pemp_adjoint_labels_extra_bytes(int images, int H, int W, int h, int w) {
  return adj_fast_geometry(H, W, h, w) ? 0 : align_up(static_cast<size_t>(images) * 2 * H * W * sizeof(float), 256);
}
// Same as pemp_adjoint_launch for the label map of `images` support images (2 * images weight planes come out).
int pemp_adjoint_launch_labels(const uint8_t* labels, int images, int H, int W, int h, int w, float* wt, float* msum, char* scratch,
                               float* expanded, cudaStream_t st) {
  const int planes = 2 * images;
  if (!adj_fast_geometry(H, W, h, w)) {
    if (!expanded) return PEMP_E_WORKSPACE;
    const long long total = static_cast<long long>(images) * H * W;
    labels_expand_kernel<<<static_cast<unsigned>(llmin((total + 255) / 256, 148LL * 16)), 256, 0, st>>>(labels, expanded, images,
                                                                                                      static_cast<long long>(H) * W);
    return pemp_bilinear_adjoint(expanded, planes, H, W, h, w, wt, msum, reinterpret_cast<pemp_stream_t>(st));
  }
  const AdjScratch a = adj_scratch(scratch, planes, h, w);
  adjoint_tables_kernel<<<1, 256, 0, st>>>(a.t, H, W, h, w, lerp_scale(h, H), lerp_scale(w, W));
  adjoint_rows2_launch<2, uint8_t>(dim3(h, images), 4 * adj_row_floats(W) * sizeof(float), st, labels, a.t, a.ab, a.rsum, H, W, h, w);
  adjoint_combine_kernel<<<planes, 256, 0, st>>>(a.ab, a.rsum, wt, msum, h, w);
  return launch_status();
}

extern "C" int pemp_bilinear_adjoint(const float* mask, int planes, int H, int W, int h, int w, float* wt, float* msum,
                                     pemp_stream_t stream) {
  PEMP_REQUIRE(mask && wt, PEMP_E_NULL);
  PEMP_REQUIRE(planes > 0 && H > 0 && W > 0 && h > 0 && w > 0 && W <= 12000, PEMP_E_SHAPE);
  cudaStream_t st = as_stream(stream);
  dim3 grid(h, planes);
  bilinear_adjoint_kernel<<<grid, 256, W * sizeof(float), st>>>(mask, wt, msum, H, W, h, w, lerp_scale(h, H),
                                                                lerp_scale(w, W));
  if (msum) plane_sum_kernel<<<planes, 1024, 0, st>>>(mask, msum, static_cast<long long>(H) * W);
  return launch_status();
}
