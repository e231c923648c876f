// Probe: can a TMA box start at an inner coordinate that is NOT a multiple of 16 bytes?
// Tensor = [groups][4*hw] fp32 (pitch 16*hw bytes, hw odd), box = 32 floats x 128 groups, SWIZZLE_128B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_unaligned_probe.cu -lcuda
//   ./tma_probe <inner_coord> [prefetch 0|1]
// Prints the number of mismatching elements (after undoing the swizzle); a trap shows up as a CUDA error.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int pf, float* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  float* box = reinterpret_cast<float*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (pf) asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128 * 32 * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(box)),
        "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(0)
        : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar))
      : "memory");
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
    const int r = i >> 5, px = i & 31;
    out[i] = box[r * 32 + (((px >> 2) ^ (r & 7)) << 2) + (px & 3)];
  }
}

int main(int argc, char** argv) {
  const int c0 = argc > 1 ? atoi(argv[1]) : 0, pf = argc > 2 ? atoi(argv[2]) : 0;
  const int hw = 2601, groups = 256;
  std::vector<float> h(static_cast<size_t>(groups) * 4 * hw);
  for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<float>(i % 1000003);
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 128 * 32 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t dims[3] = {4ull * hw, 128, 2};
  cuuint64_t strides[2] = {16ull * hw, 16ull * hw * 128};
  cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<Fn>(p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode=%d ", static_cast<int>(r));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 32 * 4 + 1024);
  probe<<<1, 256, 128 * 32 * 4 + 1024>>>(map, c0, 0, pf, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("coord=%d prefetch=%d sync=%s ", c0, pf, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> o(128 * 32);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int g = 0; g < 128; ++g)
      for (int x = 0; x < 32; ++x)
        if (o[g * 32 + x] != h[static_cast<size_t>(g) * 4 * hw + c0 + x]) ++bad;
    printf("mismatches=%d", bad);
  }
  printf("\n");
  return 0;
}
