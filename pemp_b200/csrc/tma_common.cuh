// Shared pieces of the TMA-fed persistent kernels (mpa_tma.cu, cosine_tma.cu, pool_tma.cu): mbarrier / bulk-tensor PTX
// wrappers, the flat tile partition, and the "four rows per group" tensor map that makes [.., c, hw] fp32 maps with an
// odd hw loadable by TMA.
//
// Layout contract of that map (see mpa_tma.cu for the derivation): the operand is [episodes][images * c/4 groups]
// [4*hw floats]; a box of 32 floats x (c/4) groups whose inner coordinate is  (e*hw + x_nom) & ~3  holds the channels
// 4g + e, column i = pixel x_nom + i - o_e with o_e = (e*hw + x_nom) & 3; 128-byte swizzle (16-byte chunk j of row r is at
// chunk j ^ (r & 7)); box origins must be 16-byte aligned (tools/probes/tma_unaligned_probe.cu).
//
// Ring contract: a ring is a whole number of tiles (4 boxes per tile), so slot s always carries the same channel class
// and the warp that waits for use u+1 of a slot is the one that consumed use u - mbarrier parity waits are only sound
// when the waiter is at most one phase behind (a slot shared by two classes failed in the field, see cosine_tma.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

// L2 promotion of the box rows (128-byte pieces at a 16*hw-byte pitch).  Measured on K2 / K3 / K1 at the bench shapes:
// 256 B 0.320 / 0.066 / 0.285 ms, 128 B and none 0.345 / 0.072 / 0.329 ms, 64 B 0.424 / 0.091 / 0.410 ms.
#ifndef PEMP_TMA_L2PROMO
#define PEMP_TMA_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif

namespace pemp_tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// %2 is the suspend-time hint: the thread sleeps in hardware until the phase completes instead of spinning
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PEMP_TMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra PEMP_TMA_DONE;\n"
      "bra PEMP_TMA_WAIT;\n"
      "PEMP_TMA_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// CTA that owns flat tile t when CTA b owns [T*b/G, T*(b+1)/G)
__host__ __device__ inline int owner_of(long long t, long long T, int G) { return static_cast<int>(((t + 1) * G - 1) / T); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;     // immutable after first resolution; benign race (same value)
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [episodes][images_per_episode * c/4 groups][4*hw floats] over `base` (16-byte aligned), episode stride in floats
// (a multiple of 4); box = 32 floats x c/4 groups x 1.  false: not encodable (the caller falls back to a generic kernel).
// box_rows: groups per box (default c/4 = all channels of one class; the backward kernel loads half boxes).
static inline bool make_rows4_map(CUtensorMap* map, const float* base, int episodes, int images_per_episode, int c, int hw,
                                  long long episode_stride, int box_rows = 0, CUtensorMapL2promotion promo = PEMP_TMA_L2PROMO) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (episode_stride & 3) != 0 || (c & 3) != 0) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(4) * hw, static_cast<cuuint64_t>(images_per_episode) * (c / 4),
                        static_cast<cuuint64_t>(episodes)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(16) * hw, static_cast<cuuint64_t>(episode_stride) * 4};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_rows > 0 ? box_rows : c / 4), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace pemp_tma
