#include "compress.h"

#include "../kernels/launch.h"
#include "../kernels/prims.cuh"

namespace ub {

namespace {

__host__ __device__ inline uint64_t cmp_align(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct CmpLayout {
  uint64_t nblocks, meta_off, offs_off, raw_off, packed_off, rest_bytes;
};
__host__ __device__ inline CmpLayout cmp_layout(uint64_t count, int dtype) {
  CmpLayout l;
  l.nblocks = (count + kCmpBlock - 1) / kCmpBlock;
  l.rest_bytes = dtype == kBF16 ? 1 : 3;
  l.meta_off = 64;
  l.offs_off = cmp_align(l.meta_off + l.nblocks * 2, 16);
  l.raw_off = cmp_align(l.offs_off + (l.nblocks + 1) * 4, 16);
  l.packed_off = cmp_align(l.raw_off + l.nblocks * kCmpBlock * l.rest_bytes, 16);
  return l;
}

template <int DT>
__device__ __forceinline__ uint32_t load_elem(const void* src, uint64_t i) {
  if constexpr (DT == kBF16) return reinterpret_cast<const uint16_t*>(src)[i];
  else return reinterpret_cast<const uint32_t*>(src)[i];
}
template <int DT>
__device__ __forceinline__ uint32_t exponent_of(uint32_t v) {
  if constexpr (DT == kBF16) return (v >> 7) & 0xffu;
  else return (v >> 23) & 0xffu;
}

// pass 1: per block exponent range -> {min, width}; packed size in 32-bit words
template <int DT>
__global__ void __launch_bounds__(256) cmp_analyze_kernel(const void* src, uint64_t count, unsigned char* out) {
  const CmpLayout l = cmp_layout(count, DT);
  const uint64_t b = blockIdx.x;
  if (b >= l.nblocks) return;
  const uint64_t base = b * kCmpBlock;
  uint32_t mn = 255, mx = 0;
  for (int i = threadIdx.x; i < kCmpBlock; i += blockDim.x) {
    const uint64_t e = base + i;
    if (e < count) {
      const uint32_t x = exponent_of<DT>(load_elem<DT>(src, e));
      mn = min(mn, x);
      mx = max(mx, x);
    }
  }
  __shared__ uint32_t s_mn[8], s_mx[8];
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) s_mn[threadIdx.x >> 5] = mn, s_mx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mn = min(mn, s_mn[w]), mx = max(mx, s_mx[w]);
    const uint32_t range = mx >= mn ? mx - mn : 0;
    const uint32_t width = range == 0 ? 0 : 32 - __clz(range);
    out[l.meta_off + b * 2] = (unsigned char)mn;
    out[l.meta_off + b * 2 + 1] = (unsigned char)width;
    reinterpret_cast<uint32_t*>(out + l.offs_off)[b] = width * (kCmpBlock / 32);  // size, scanned in pass 2
  }
}

// pass 2: exclusive scan of the per-block packed sizes (single CTA; <= a few 100k blocks) + header
__global__ void __launch_bounds__(1024) cmp_scan_kernel(uint64_t count, int dtype, unsigned char* out) {
  const CmpLayout l = cmp_layout(count, dtype);
  uint32_t* offs = reinterpret_cast<uint32_t*>(out + l.offs_off);
  const uint64_t n = l.nblocks;
  const uint64_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint64_t lo = min(n, (uint64_t)threadIdx.x * per), hi = min(n, lo + per);
  uint64_t sum = 0;
  for (uint64_t i = lo; i < hi; ++i) sum += offs[i];
  __shared__ uint64_t s_part[1024];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t acc = 0;
    for (int t = 0; t < (int)blockDim.x; ++t) {
      const uint64_t v = s_part[t];
      s_part[t] = acc;
      acc += v;
    }
    offs[n] = (uint32_t)acc;  // total packed words
    CmpHeader* h = reinterpret_cast<CmpHeader*>(out);
    h->magic = kCmpMagic;
    h->dtype = (uint32_t)dtype;
    h->count = count;
    h->nblocks = n;
    h->meta_off = l.meta_off, h->offs_off = l.offs_off, h->raw_off = l.raw_off, h->packed_off = l.packed_off;
    h->total_bytes = l.packed_off + acc * 4;
  }
  __syncthreads();
  uint64_t acc = s_part[threadIdx.x];
  for (uint64_t i = lo; i < hi; ++i) {
    const uint32_t v = offs[i];
    offs[i] = (uint32_t)acc;
    acc += v;
  }
}

// pass 3: write mantissa byte planes and exponent bit planes
template <int DT>
__global__ void __launch_bounds__(256) cmp_pack_kernel(const void* src, uint64_t count, unsigned char* out) {
  const CmpLayout l = cmp_layout(count, DT);
  const uint64_t b = blockIdx.x;
  if (b >= l.nblocks) return;
  const uint64_t base = b * kCmpBlock;
  const uint32_t mn = out[l.meta_off + b * 2], width = out[l.meta_off + b * 2 + 1];
  uint32_t* packed = reinterpret_cast<uint32_t*>(out + l.packed_off) + reinterpret_cast<const uint32_t*>(out + l.offs_off)[b];
  unsigned char* raw = out + l.raw_off + b * kCmpBlock * l.rest_bytes;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kCmpBlock; i += blockDim.x) {  // i/32 is warp-uniform
    const uint64_t e = base + i;
    const uint32_t v = e < count ? load_elem<DT>(src, e) : 0;
    const uint32_t delta = e < count ? exponent_of<DT>(v) - mn : 0;
    if constexpr (DT == kBF16) {
      raw[i] = (unsigned char)(((v >> 8) & 0x80u) | (v & 0x7fu));  // sign + 7 mantissa bits
    } else {
      const uint32_t rest = ((v >> 8) & 0x800000u) | (v & 0x7fffffu);  // sign + 23 mantissa bits
      raw[i] = (unsigned char)rest;
      raw[kCmpBlock + i] = (unsigned char)(rest >> 8);
      raw[2 * kCmpBlock + i] = (unsigned char)(rest >> 16);
    }
    const int group = i >> 5;
    for (uint32_t j = 0; j < width; ++j) {
      const uint32_t plane = __ballot_sync(0xffffffffu, (delta >> j) & 1u);
      if (lane == 0) packed[(uint64_t)group * width + j] = plane;
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256) cmp_unpack_kernel(const unsigned char* in, void* dst, uint64_t count) {
  const CmpLayout l = cmp_layout(count, DT);
  const uint64_t b = blockIdx.x;
  if (b >= l.nblocks) return;
  const uint64_t base = b * kCmpBlock;
  const uint32_t mn = in[l.meta_off + b * 2], width = in[l.meta_off + b * 2 + 1];
  const uint32_t* packed =
      reinterpret_cast<const uint32_t*>(in + l.packed_off) + reinterpret_cast<const uint32_t*>(in + l.offs_off)[b];
  const unsigned char* raw = in + l.raw_off + b * kCmpBlock * l.rest_bytes;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kCmpBlock; i += blockDim.x) {
    const uint64_t e = base + i;
    const int group = i >> 5;
    uint32_t delta = 0;
    for (uint32_t j = 0; j < width; ++j) delta |= ((packed[(uint64_t)group * width + j] >> lane) & 1u) << j;
    if (e >= count) continue;
    const uint32_t ex = mn + delta;
    if constexpr (DT == kBF16) {
      const uint32_t r = raw[i];
      reinterpret_cast<uint16_t*>(dst)[e] = (uint16_t)(((r & 0x80u) << 8) | (ex << 7) | (r & 0x7fu));
    } else {
      const uint32_t r = (uint32_t)raw[i] | ((uint32_t)raw[kCmpBlock + i] << 8) | ((uint32_t)raw[2 * kCmpBlock + i] << 16);
      reinterpret_cast<uint32_t*>(dst)[e] = ((r & 0x800000u) << 8) | (ex << 23) | (r & 0x7fffffu);
    }
  }
}

}  // namespace

bool cmp_dtype_supported(int dtype) { return dtype == kBF16 || dtype == kF32; }

size_t cmp_bound(size_t count, int dtype) {
  const CmpLayout l = cmp_layout(count, dtype);
  return (size_t)(l.packed_off + l.nblocks * (uint64_t)kCmpBlock);  // 8 bits per exponent worst case
}

cudaError_t cmp_compress_async(const void* src, size_t count, int dtype, void* dst, cudaStream_t st) {
  if (!cmp_dtype_supported(dtype)) return cudaErrorInvalidValue;
  const CmpLayout l = cmp_layout(count, dtype);
  unsigned char* out = reinterpret_cast<unsigned char*>(dst);
  const int grid = (int)(l.nblocks ? l.nblocks : 1);
  if (dtype == kBF16) UB_LAUNCH((cmp_analyze_kernel<kBF16>), grid, 256, 0, st, src, (uint64_t)count, out);
  else UB_LAUNCH((cmp_analyze_kernel<kF32>), grid, 256, 0, st, src, (uint64_t)count, out);
  UB_LAUNCH((cmp_scan_kernel), 1, 1024, 0, st, (uint64_t)count, dtype, out);
  if (dtype == kBF16) UB_LAUNCH((cmp_pack_kernel<kBF16>), grid, 256, 0, st, src, (uint64_t)count, out);
  else UB_LAUNCH((cmp_pack_kernel<kF32>), grid, 256, 0, st, src, (uint64_t)count, out);
  return cudaGetLastError();
}

cudaError_t cmp_decompress_async(const void* src, void* dst, size_t count, int dtype, cudaStream_t st) {
  if (!cmp_dtype_supported(dtype)) return cudaErrorInvalidValue;
  const CmpLayout l = cmp_layout(count, dtype);
  const int grid = (int)(l.nblocks ? l.nblocks : 1);
  const unsigned char* in = reinterpret_cast<const unsigned char*>(src);
  if (dtype == kBF16) UB_LAUNCH((cmp_unpack_kernel<kBF16>), grid, 256, 0, st, in, dst, (uint64_t)count);
  else UB_LAUNCH((cmp_unpack_kernel<kF32>), grid, 256, 0, st, in, dst, (uint64_t)count);
  return cudaGetLastError();
}

}  // namespace ub
