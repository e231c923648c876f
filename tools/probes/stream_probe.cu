// Micro-benchmark: how does HBM throughput depend on the width of the contiguous piece each CTA reads from every
// channel row of a [N, 512, 2601] fp32 tensor (row pitch 10404 B)?  One CTA = one tile of W pixels x 512 rows.
//   nvcc -arch=sm_100a -O3 -o /tmp/stream_probe tools/probes/stream_probe.cu && /tmp/stream_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int PPL>   // pixels per lane: tile width = 32 * PPL
__global__ void __launch_bounds__(256) probe(const float* __restrict__ x, int c, int hw, float* __restrict__ out) {
  const int tiles = (hw + 32 * PPL - 1) / (32 * PPL);
  const int img = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = min(tile * 32 * PPL, hw - 32 * PPL);
  const float* p = x + ((long long)img * c + warp * (c / 8)) * hw + x0 + lane;
  float acc = 0.f;
  constexpr int ROWS = 32 / PPL < 1 ? 1 : 32 / PPL;      // keep 32 loads in flight per lane
  for (int r = 0; r < c / 8; r += ROWS) {
    float v[ROWS][PPL];
#pragma unroll
    for (int i = 0; i < ROWS; ++i)
#pragma unroll
      for (int s = 0; s < PPL; ++s) v[i][s] = __ldg(p + (long long)(r + i) * hw + 32 * s);
#pragma unroll
    for (int i = 0; i < ROWS; ++i)
#pragma unroll
      for (int s = 0; s < PPL; ++s) acc += v[i][s];
  }
  if (acc == 12345.678f) out[0] = acc;
}

template <int PPL>
void run(const float* x, int n, int c, int hw, float* out) {
  const int tiles = (hw + 32 * PPL - 1) / (32 * PPL);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) probe<PPL><<<n * tiles, 256>>>(x, c, hw, out);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) probe<PPL><<<n * tiles, 256>>>(x, c, hw, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  double bytes = (double)n * c * hw * 4;
  printf("tile width %4d px (%5d B per row piece): %.3f ms  %.0f GB/s (%d CTAs)\n", 32 * PPL, 128 * PPL, ms, bytes / ms / 1e6, n * tiles);
}

int main() {
  const int n = 320, c = 512, hw = 2601;
  float *x, *out;
  cudaMalloc(&x, (size_t)n * c * hw * 4);
  cudaMalloc(&out, 4);
  cudaMemset(x, 0, (size_t)n * c * hw * 4);
  run<1>(x, n, c, hw, out);
  run<2>(x, n, c, hw, out);
  run<4>(x, n, c, hw, out);
  run<8>(x, n, c, hw, out);
  run<16>(x, n, c, hw, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
