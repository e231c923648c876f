// Probe: issue rate of the warp-level `mma.sync.m16n8k8` tf32 instruction on sm_100a (legacy tensor path, SASS HMMA).
// The backward kernels of the head (csrc/train.cu) are three thin GEMMs per 32-pixel tile (K = 2P + 4 = 10 table columns);
// their CUDA-core form is bound by broadcast table-row reads from shared memory, so the question is only whether the legacy
// path is fast enough for ~2000 m16n8k8 per 128 KB of HBM traffic per SM (~5000 clocks at the HBM rate).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/bin/mma_sync_probe tools/probes/mma_sync_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(1024) mma_loop(int iters, float* out, long long* clk) {
  float c[CHAINS][4];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3u, a2 = threadIdx.x * 5u, a3 = threadIdx.x * 7u, b0 = 0x3f800000u, b1 = 0x3f000000u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&clk, 148 * sizeof(long long));
  const int iters = 2000;
  for (int threads : {128, 256, 512, 1024}) {
    mma_loop<8><<<148, threads>>>(iters, out, clk);
    cudaDeviceSynchronize();
    mma_loop<8><<<148, threads>>>(iters, out, clk);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double mmas_per_sm = static_cast<double>(iters) * 8 * (threads / 32);
    printf("{\"probe\": \"mma.sync m16n8k8 tf32\", \"threads_per_sm\": %d, \"chains_per_warp\": 8, \"clk_per_mma_per_sm\": %.3f, "
           "\"tf32_fma_per_clk_per_sm\": %.1f, \"err\": %d}\n",
           threads, mx / mmas_per_sm, 1024.0 * mmas_per_sm / mx, static_cast<int>(e));
  }
  return 0;
}
