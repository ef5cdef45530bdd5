// Shared host/device definitions for the collective kernels: control-region
// layout inside the symmetric heap, launch arguments, slicing helpers.
#pragma once
#include "reduce.cuh"

namespace ub {

// 16-byte load/store with a partial tail (only `bytes - off` bytes valid).
__device__ __forceinline__ uint4 load16_partial(const void* base, uint64_t off, uint64_t bytes) {
  const char* p = (const char*)base + off;
  if (off + 16 <= bytes) return ld_v4(p);
  alignas(16) unsigned char tmp[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tmp[i] = (off + i < bytes) ? ((const volatile unsigned char*)p)[i] : 0;
  return *reinterpret_cast<uint4*>(tmp);
}
__device__ __forceinline__ void store16_partial(void* base, uint64_t off, uint64_t bytes, const uint4& v) {
  char* p = (char*)base + off;
  if (off + 16 <= bytes) {
    st_v4(p, v);
    return;
  }
  const unsigned char* s = reinterpret_cast<const unsigned char*>(&v);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (off + i < bytes) ((unsigned char*)p)[i] = s[i];
}

}  // namespace ub
