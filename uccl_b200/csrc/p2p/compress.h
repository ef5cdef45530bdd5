// Lossless float compression hook for large transfers (role of the reference's DietGPU hook,
// p2p/rdma/compression.h:12-608: strategies none / split / encode for tensors > 2 MiB).
//
// Codec ("split + frame-of-reference bit planes"): a float is split into its exponent byte and the
// remaining sign/mantissa bytes.  Mantissas are incompressible and are stored as byte planes; the
// exponents of one 4096-element block span only a handful of values in real tensors, so they are
// stored as (exponent - block minimum) in w = ceil(log2(range+1)) bits, laid out as warp-ballot bit
// planes (one 32-bit word per bit per 32 elements -- pack and unpack are a __ballot_sync and a
// shift, no serial bit twiddling).  bf16 typically shrinks to ~0.8x, fp32 to ~0.9x.  On NVLink
// (>600 GB/s) this never pays for itself, which is why the hook is off by default
// (UCCL_B200_P2P_COMPRESS=none); it exists for API parity and for slower links.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ub {

constexpr int kCmpBlock = 4096;        // elements per block
constexpr uint32_t kCmpMagic = 0x55434d50u;  // "UCMP"

struct CmpHeader {       // first 64 bytes of a compressed buffer
  uint32_t magic;
  uint32_t dtype;        // ub::DType (kBF16 or kF32)
  uint64_t count;        // elements
  uint64_t nblocks;
  uint64_t total_bytes;  // whole compressed size including this header
  uint64_t meta_off, offs_off, raw_off, packed_off;
};
static_assert(sizeof(CmpHeader) == 64, "CmpHeader is 64 bytes");

bool cmp_dtype_supported(int dtype);
// worst-case compressed size (w = 8 everywhere) -- what the caller must allocate
size_t cmp_bound(size_t count, int dtype);
// Compress `count` elements at `src` into `dst` (capacity >= cmp_bound).  The compressed size is
// written to dst's header (total_bytes) on the device; read it back after the stream has run.
cudaError_t cmp_compress_async(const void* src, size_t count, int dtype, void* dst, cudaStream_t st);
// Decompress a buffer produced by cmp_compress_async; `count`/`dtype` must match the header.
cudaError_t cmp_decompress_async(const void* src, void* dst, size_t count, int dtype, cudaStream_t st);

}  // namespace ub
