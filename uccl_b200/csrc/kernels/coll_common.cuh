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

// Aligned 16-byte-unit copy of units [lo, hi) with 8 loads in flight per thread.
__device__ __forceinline__ void copy_units16(char* dst, const char* src, uint64_t lo, uint64_t hi) {
  constexpr int B = 8;
  uint64_t u = lo + threadIdx.x;
  for (; u + (uint64_t)(B - 1) * blockDim.x < hi; u += (uint64_t)B * blockDim.x) {
    uint4 v[B];
#pragma unroll
    for (int j = 0; j < B; ++j) v[j] = ld_v4(src + (u + (uint64_t)j * blockDim.x) * 16);
#pragma unroll
    for (int j = 0; j < B; ++j) st_v4(dst + (u + (uint64_t)j * blockDim.x) * 16, v[j]);
  }
  for (; u < hi; u += blockDim.x) st_v4(dst + u * 16, ld_v4(src + u * 16));
}

// 16-byte units [blo, bhi) of the current chunk (`cb` bytes) that block blockIdx.x owns.
// The boundaries are those of a FULL chunk (clamped to the current one), so a block touches the
// same range of the staging area in every chunk of a kernel.  The kernels only synchronise
// same-index blocks across ranks; if a short tail chunk were re-sliced, a faster peer block of
// another index could overwrite stage bytes that this block is still reading from the previous
// chunk (observed as corrupted results with 8 ranks and a message slightly larger than the stage).
__device__ __forceinline__ void chunk_slice(uint64_t msg_bytes, uint64_t chunk_bytes, uint64_t cb, uint64_t& blo,
                                            uint64_t& bhi, uint64_t gran = 1) {
  chunk_slice_hd(msg_bytes, chunk_bytes, cb, (int)gridDim.x, (int)blockIdx.x, blo, bhi, gran);
}

}  // namespace ub
