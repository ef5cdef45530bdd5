// In-kernel bulk copy engine for the P2P transfer path (KV-cache / weight moves).
//
// The reference moves intra-node bytes with cudaMemcpyAsync striped over 4 streams
// (p2p/engine.cc:1710-1767) -- i.e. the copy engines.  Here each CTA runs a TMA pipeline:
// one elected thread issues cp.async.bulk global->shared (mbarrier tracked) a few stages
// ahead and cp.async.bulk shared->global behind, so a vector of (src, dst, bytes) blocks --
// e.g. all KV blocks of a request -- is moved by ONE launch, local<->peer in either
// direction, with no per-block launch or per-block stream/event bookkeeping.
#include "../kernels/launch.h"
#include "../kernels/prims.cuh"
#include "p2p_types.h"

namespace ub {

constexpr int kCopyThreads = 128;

__global__ void __launch_bounds__(kCopyThreads) p2p_copy_kernel(const __grid_constant__ P2PCopyBatch b) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kP2PStages];
  const int tid = threadIdx.x;
  const uint32_t chunk = b.chunk_bytes;

  if (tid == 0) {
    for (int s = 0; s < kP2PStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- warp 0 / lane 0: TMA pipeline over the 16-byte aligned bulk chunks
  if (tid == 0) {
    const uint32_t total = b.chunk_prefix[b.n];
    uint32_t issued = 0, stored = 0;
    uint32_t phase_bits = 0;  // per-stage parity
    // my chunks: blockIdx.x, blockIdx.x + gridDim.x, ...
    auto locate = [&](uint32_t g, const char*& src, char*& dst, uint32_t& bytes) {
      int e = 0;
      while (e + 1 < b.n && g >= b.chunk_prefix[e + 1]) ++e;
      const uint64_t off = (uint64_t)(g - b.chunk_prefix[e]) * chunk;
      const uint64_t bulk = b.e[e].bulk_bytes;
      bytes = (uint32_t)((bulk - off) < chunk ? (bulk - off) : chunk);
      src = b.e[e].src + off;
      dst = b.e[e].dst + off;
    };
    const uint32_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t mine = first < total ? (total - first + stride - 1) / stride : 0;
    while (stored < mine) {
      // keep up to kP2PStages - 1 loads in flight
      while (issued < mine && issued < stored + kP2PStages) {
        const int s = issued % kP2PStages;
        if (issued >= kP2PStages) tma_store_wait_read<kP2PStages - 1>();  // stage's previous store has read smem
        const char* src;
        char* dst;
        uint32_t bytes;
        locate(first + issued * stride, src, dst, bytes);
        mbar_expect_tx(&full[s], bytes);
        tma_load_1d(smem + (size_t)s * chunk, src, bytes, &full[s]);
        ++issued;
      }
      const int s = stored % kP2PStages;
      mbar_wait(&full[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      const char* src;
      char* dst;
      uint32_t bytes;
      locate(first + stored * stride, src, dst, bytes);
      tma_store_1d(dst, smem + (size_t)s * chunk, bytes);
      tma_store_commit();
      ++stored;
    }
    tma_store_wait<0>();
  } else if (tid >= 32) {
    // ---- other warps: unaligned entries and < 16-byte tails with plain loads/stores
    const int t = tid - 32, nt = kCopyThreads - 32;
    for (int e = 0; e < b.n; ++e) {
      const uint64_t bulk = b.e[e].bulk_bytes, bytes = b.e[e].bytes;
      if (bulk == bytes) continue;
      const char* src = b.e[e].src;
      char* dst = b.e[e].dst;
      for (uint64_t i = bulk + (uint64_t)blockIdx.x * nt + t; i < bytes; i += (uint64_t)gridDim.x * nt) dst[i] = src[i];
    }
  }
}

cudaError_t launch_p2p_copy(const P2PCopyBatch& b, int grid, cudaStream_t st) {
  const size_t smem = (size_t)kP2PStages * b.chunk_bytes;
  if (g_preload || smem > 48 * 1024) {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(p2p_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
      attr_done[dev & 63] = true;
    }
  }
  UB_LAUNCH((p2p_copy_kernel), grid, kCopyThreads, smem, st, b);
  return cudaGetLastError();
}

cudaError_t preload_p2p_kernels() {
  P2PCopyBatch b;
  memset(&b, 0, sizeof(b));
  b.chunk_bytes = 16384;
  return launch_p2p_copy(b, 1, 0);
}

}  // namespace ub
