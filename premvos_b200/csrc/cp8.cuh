// Device helpers for CP8 activations (split-bf16 chunk planes, see CView in common.cuh).
#pragma once
#include "common.cuh"

namespace premvos {
namespace cp8 {

struct F8 { float v[8]; };

__device__ __forceinline__ F8 zero8() {
  F8 r;
#pragma unroll
  for (int j = 0; j < 8; j++) r.v[j] = 0.f;
  return r;
}
__device__ __forceinline__ F8 ld_chunk(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long elem /* multiple of 8 */) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi + elem);
  const uint4 l = *reinterpret_cast<const uint4*>(lo + elem);
  const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
  F8 r;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    r.v[2 * j] = __uint_as_float(hh[j] << 16) + __uint_as_float(ll[j] << 16);
    r.v[2 * j + 1] = __uint_as_float(hh[j] & 0xffff0000u) + __uint_as_float(ll[j] & 0xffff0000u);
  }
  return r;
}
__device__ __forceinline__ void st_chunk(__nv_bfloat16* hi, __nv_bfloat16* lo, long elem, const F8& f) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float x0 = f.v[2 * j], x1 = f.v[2 * j + 1];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi + elem) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(lo + elem) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

struct CV {  // device copy of a CView
  __nv_bfloat16* hi; __nv_bfloat16* lo; int N, H, W, chunks, c0, C;
};
inline CV dev(const CView& v) { return CV{v.hi, v.lo, v.N, v.H, v.W, v.chunks, v.c0, v.C}; }
__device__ __forceinline__ long cv_elem(const CV& v, int n, int chunk, int y, int x) {
  return ((((long)n * v.chunks + v.c0 + chunk) * v.H + y) * v.W + x) * 8;
}
inline unsigned blocks_for(long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

}  // namespace cp8
}  // namespace premvos
