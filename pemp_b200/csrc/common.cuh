// Shared helpers for the sm_100a kernels of libpemp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pemp_b200.h"

#define PEMP_REQUIRE(cond, code) \
  do {                           \
    if (!(cond)) return (code);  \
  } while (0)

static inline cudaStream_t as_stream(pemp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Every entry point ends with this: report the launch status without synchronising.
static inline int launch_status() { return static_cast<int>(cudaPeekAtLastError()); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ATen's nearest rule (UpSample.h nearest_neighbor_compute_source_index): floor(dst * scale), clamped.
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  int s = static_cast<int>(floorf(static_cast<float>(dst) * scale));
  return s < in_size - 1 ? s : in_size - 1;
}

// ATen's align_corners=True bilinear coefficients (UpSample.h area_pixel_compute_source_index):
//   real = scale * dst;  i0 = (int)real;  i1 = i0 + (i0 < in-1);  l1 = real - i0;  l0 = 1 - l1.
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_coeff(int dst, float scale, int in_size) {
  Lerp r;
  float real = __fmul_rn(scale, static_cast<float>(dst));
  r.i0 = static_cast<int>(real);
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = __fsub_rn(real, static_cast<float>(r.i0));
  r.l0 = __fsub_rn(1.0f, r.l1);
  return r;
}
static inline float lerp_scale(int in_size, int out_size) {
  return out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.0f;
}
// out = fma(l0, a, l1*b): the operand order ATen's kernels compile to (probed against torch 2.11 CPU,
// bit exact; the probe is recorded in DESIGN.md).
__device__ __forceinline__ float lerp2(float l0, float a, float l1, float b) {
  return __fmaf_rn(l0, a, __fmul_rn(l1, b));
}
