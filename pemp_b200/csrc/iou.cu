// K10  FewShotMetric.update(): per-episode confusion counts accumulated into stat[(C+1), 3].
//
// replaces core/metrics.py:9-23 (NumPy, 12 boolean passes per episode on the host).
//   for j in {0 (background), 1 (foreground)}, over pixels with ref != 255:
//     tp_j = #(pred == j and ref == j);  fp_j = #(pred == j and ref != j);  fn_j = #(pred != j and ref == j)
//   stat[0] += (tp_0, fp_0, fn_0);  stat[cls] += (tp_1, fp_1, fn_1)
//
// Roofline: HBM, 2 bytes per pixel.  One pass with 16-byte loads and byte-SIMD compares; counts are
// integers, so the int64 atomics into `stat` are order independent and the result is exact.
#include "common.cuh"
#include "iou_count.cuh"

namespace {

constexpr int kThreads = 256;

using Counts = PempCounts;
#define count_word pemp_count_word
#define count_byte pemp_count_byte

// grid = (chunks, N).  Each CTA handles a contiguous byte range of one episode.
__global__ void __launch_bounds__(kThreads)
iou_hist_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ ref, const int64_t* __restrict__ cls,
                long long npix, int num_classes, unsigned long long* __restrict__ stat) {
  const int n = blockIdx.y;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per, hi = min(npix, lo + per);
  const uint8_t* p = pred + n * npix;
  const uint8_t* r = ref + n * npix;
  Counts k = {0, 0, 0, 0, 0, 0};
  if (lo < hi) {
    // vector body only when both streams share their 16-byte phase (always true for two 16B-aligned tensors)
    const uintptr_t pa = reinterpret_cast<uintptr_t>(p + lo), ra = reinterpret_cast<uintptr_t>(r + lo);
    long long head = hi - lo;
    long long nvec = 0;
    if ((pa & 15) == (ra & 15)) {
      head = llmin(hi - lo, (16 - (pa & 15)) & 15);
      nvec = (hi - lo - head) / 16;
    }
    for (long long i = threadIdx.x; i < head; i += kThreads) count_byte(p[lo + i], r[lo + i], k);
    const uint4* pv = reinterpret_cast<const uint4*>(p + lo + head);
    const uint4* rv = reinterpret_cast<const uint4*>(r + lo + head);
    for (long long i = threadIdx.x; i < nvec; i += kThreads) {
      uint4 a = __ldg(pv + i), b = __ldg(rv + i);
      count_word(a.x, b.x, k);
      count_word(a.y, b.y, k);
      count_word(a.z, b.z, k);
      count_word(a.w, b.w, k);
    }
    for (long long i = lo + head + nvec * 16 + threadIdx.x; i < hi; i += kThreads) count_byte(p[i], r[i], k);
  }
  // block reduction
  __shared__ unsigned red[6][kThreads / 32];
  unsigned v[6] = {k.tp0, k.fp0, k.fn0, k.tp1, k.fp1, k.fn1};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    unsigned s = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (lane == 0) red[q][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    unsigned long long s = 0;
    for (int wv = 0; wv < kThreads / 32; ++wv) s += red[threadIdx.x][wv];
    s >>= 3;   // popc counted 8 bits per matching byte
    if (s) {
      const int q = threadIdx.x;
      long long row = q < 3 ? 0 : cls[n];
      if (row >= 0 && row <= num_classes) atomicAdd(stat + row * 3 + (q % 3), s);
    }
  }
}

}  // namespace

extern "C" int pemp_iou_hist(const uint8_t* pred, const uint8_t* ref, const int64_t* cls, int N, long long npix,
                             int num_classes, int64_t* stat, pemp_stream_t stream) {
  PEMP_REQUIRE(pred && ref && cls && stat, PEMP_E_NULL);
  PEMP_REQUIRE(N > 0 && N <= 65535 && npix > 0 && num_classes > 0, PEMP_E_SHAPE);
  // per-thread 32-bit counters hold 8 * bytes seen: keep a CTA's share below 2^28 bytes
  int chunks = static_cast<int>(llmin(64, llmax(1, npix / 8192)));
  while ((npix + chunks - 1) / chunks > (1LL << 28)) chunks *= 2;
  dim3 grid(chunks, N);
  iou_hist_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(pred, ref, cls, npix, num_classes,
                                                            reinterpret_cast<unsigned long long*>(stat));
  return launch_status();
}
