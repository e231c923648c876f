// Byte-SIMD confusion counts shared by K10 (iou.cu) and the fused up-sample + argmax + counts kernel (resample.cu).
//   for j in {0 (background), 1 (foreground)}, over pixels with ref != 255      (core/metrics.py:15-19)
//     tp_j = #(pred == j and ref == j);  fp_j = #(pred == j and ref != j);  fn_j = #(pred != j and ref == j)
// Counters hold 8 x the number of matching bytes (popc of 0xff per byte); shift right by 3 at the end.
#pragma once
#include <stdint.h>

struct PempCounts {
  unsigned tp0, fp0, fn0, tp1, fp1, fn1;
};

__device__ __forceinline__ void pemp_count_word(uint32_t p, uint32_t r, PempCounts& k) {
  const uint32_t p0 = __vcmpeq4(p, 0x00000000u), p1 = __vcmpeq4(p, 0x01010101u);
  const uint32_t r0 = __vcmpeq4(r, 0x00000000u), r1 = __vcmpeq4(r, 0x01010101u);
  const uint32_t valid = ~__vcmpeq4(r, 0xffffffffu);
  k.tp0 += __popc(p0 & r0);
  k.fp0 += __popc(p0 & ~r0 & valid);
  k.fn0 += __popc(~p0 & r0);
  k.tp1 += __popc(p1 & r1);
  k.fp1 += __popc(p1 & ~r1 & valid);
  k.fn1 += __popc(~p1 & r1);
}

__device__ __forceinline__ void pemp_count_byte(uint8_t p, uint8_t r, PempCounts& k) {
  const bool valid = r != 255;
  k.tp0 += 8u * (p == 0 && r == 0);
  k.fp0 += 8u * (p == 0 && r != 0 && valid);
  k.fn0 += 8u * (p != 0 && r == 0);
  k.tp1 += 8u * (p == 1 && r == 1);
  k.fp1 += 8u * (p == 1 && r != 1 && valid);
  k.fn1 += 8u * (p != 1 && r == 1);
}
