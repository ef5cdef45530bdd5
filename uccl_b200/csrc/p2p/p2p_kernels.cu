// In-kernel bulk copy engine for the P2P transfer path (KV-cache / weight moves).
//
// The reference moves intra-node bytes with cudaMemcpyAsync striped over 4 streams
// (p2p/engine.cc:1710-1767) -- i.e. the copy engines.  Here each CTA runs four TMA pipelines
// (one elected lane per warp): cp.async.bulk global->shared (mbarrier tracked) a few stages
// ahead and cp.async.bulk shared->global behind, so a vector of (src, dst, bytes) blocks --
// e.g. all KV blocks of a request, thousands of them through a pinned descriptor table --
// is moved by ONE launch, local<->peer in either direction, with no per-block launch or
// per-block stream/event bookkeeping.
#include "../kernels/launch.h"
#include "../kernels/prims.cuh"
#include "p2p_types.h"

namespace ub {

constexpr int kCopyThreads = kP2PWarps * 32;

// Every warp of the CTA runs its own pipeline (lane 0 issues): chunk g of the batch belongs to pipeline
// g % (gridDim.x * kP2PWarps).  kP2PStages bulk loads are in flight per pipeline and every arrived stage is
// bulk-stored to its destination right away, so a CTA keeps kP2PWarps * kP2PStages chunks moving from four
// issuing threads (the first version had ONE issuing thread per CTA and lost to cudaMemcpyPeer).
__global__ void __launch_bounds__(kCopyThreads) p2p_copy_kernel(const __grid_constant__ P2PCopyBatch b) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kP2PWarps][kP2PStages];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t chunk = b.chunk_bytes;
  const P2PCopyEntry* ent = b.table ? b.table : b.e;
  const uint32_t* pfx = b.table ? b.table_prefix : b.chunk_prefix;

  if (tid == 0) {
    for (int w = 0; w < kP2PWarps; ++w)
      for (int s = 0; s < kP2PStages; ++s) mbar_init(&full[w][s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  if (lane == 0) {
    // ---- TMA pipeline of this warp over the 16-byte aligned bulk chunks
    unsigned char* my_smem = smem + (size_t)warp * kP2PStages * chunk;
    const uint32_t total = pfx[b.n];
    const uint32_t first = blockIdx.x * kP2PWarps + warp, stride = gridDim.x * kP2PWarps;
    const uint32_t mine = first < total ? (total - first + stride - 1) / stride : 0;
    // chunk index -> (entry, offset): the indices of one pipeline only grow, so each cursor walks forward
    struct Cursor {
      int e = 0;
    } ci, cs;
    auto locate = [&](Cursor& c, uint32_t g, const char*& src, char*& dst, uint32_t& bytes) {
      if (pfx[c.e + 1] <= g) {  // jump ahead: binary search for the entry that holds chunk g
        int lo = c.e + 1, hi = b.n - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (pfx[mid] <= g) lo = mid;
          else hi = mid - 1;
        }
        c.e = lo;
      }
      const P2PCopyEntry en = ent[c.e];
      const uint64_t off = (uint64_t)(g - pfx[c.e]) * chunk;
      bytes = (uint32_t)((en.bulk_bytes - off) < chunk ? (en.bulk_bytes - off) : chunk);
      src = en.src + off;
      dst = en.dst + off;
    };
    uint32_t issued = 0, stored = 0, phase_bits = 0;
    while (stored < mine) {
      // a stage is refilled one iteration after its store was committed, so wait_group.read 1 (everything but the newest
      // store has left shared memory) covers it without blocking on the store that was just issued
      while (issued < mine && issued < stored + kP2PStages - 1) {
        const int s = issued % kP2PStages;
        if (issued >= kP2PStages) tma_store_wait_read<1>();
        const char* src;
        char* dst;
        uint32_t bytes;
        locate(ci, first + issued * stride, src, dst, bytes);
        mbar_expect_tx(&full[warp][s], bytes);
        tma_load_1d(my_smem + (size_t)s * chunk, src, bytes, &full[warp][s]);
        ++issued;
      }
      const int s = stored % kP2PStages;
      mbar_wait(&full[warp][s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      const char* src;
      char* dst;
      uint32_t bytes;
      locate(cs, first + stored * stride, src, dst, bytes);
      tma_store_1d(dst, my_smem + (size_t)s * chunk, bytes);
      tma_store_commit();
      ++stored;
    }
    tma_store_wait<0>();
  } else if (b.any_tail) {
    // ---- the other lanes: unaligned entries and < 16-byte tails with plain loads/stores
    const int t = warp * 31 + (lane - 1), nt = kP2PWarps * 31;
    for (int e = 0; e < b.n; ++e) {
      const P2PCopyEntry en = ent[e];
      if (en.bulk_bytes == en.bytes) continue;
      for (uint64_t i = en.bulk_bytes + (uint64_t)blockIdx.x * nt + t; i < en.bytes; i += (uint64_t)gridDim.x * nt)
        en.dst[i] = en.src[i];
    }
  }
}

cudaError_t launch_p2p_copy(const P2PCopyBatch& b, int grid, cudaStream_t st) {
  const size_t smem = (size_t)kP2PWarps * kP2PStages * b.chunk_bytes;
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  if (g_preload || smem > 48 * 1024) {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(p2p_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
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
