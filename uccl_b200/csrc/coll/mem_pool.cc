// torch.cuda.MemPool backend: allocations come out of a communicator's symmetric heap, so every
// tensor created under the pool is peer-mapped (and multicast-bound) and takes the zero-copy paths
// of all collectives -- the torch-side analogue of ncclMemAlloc (reference:
// experimental/lite/nccl/nccl.cu:2384-2430).  Ranks need not allocate symmetrically: the
// kernels exchange buffer offsets in their entry barrier.
//
// Exported C symbols (torch.cuda.memory.CUDAPluggableAllocator signature):
//   void* uccl_b200_pool_malloc(size_t size, int device, cudaStream_t stream)
//   void  uccl_b200_pool_free(void* ptr, size_t size, int device, cudaStream_t stream)
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>

#include "../common/log.h"
#include "comm.h"

namespace ub {

namespace {
std::mutex g_pool_mu;
std::map<int, std::weak_ptr<Comm>> g_pool_by_device;  // device -> communicator serving the pool
thread_local std::weak_ptr<Comm> t_pool_override;       // virtual ranks: several comms on one device
struct PoolStats {
  uint64_t allocs = 0, frees = 0, bytes_live = 0, fallback_allocs = 0;
} g_stats;
struct Owner {
  std::weak_ptr<Comm> comm;
  size_t size = 0;
  bool heap = false;  // false: cudaMalloc fallback block
};
std::map<void*, Owner> g_owner;
}  // namespace

void pool_install(std::shared_ptr<Comm> c) {
  std::lock_guard<std::mutex> g(g_pool_mu);
  g_pool_by_device[c->device()] = c;
}
void pool_set_thread_comm(std::shared_ptr<Comm> c) { t_pool_override = c; }
void pool_clear_thread_comm() { t_pool_override.reset(); }
void pool_stats(uint64_t* allocs, uint64_t* frees, uint64_t* live, uint64_t* fallback) {
  std::lock_guard<std::mutex> g(g_pool_mu);
  *allocs = g_stats.allocs, *frees = g_stats.frees, *live = g_stats.bytes_live, *fallback = g_stats.fallback_allocs;
}

}  // namespace ub

extern "C" {

__attribute__((visibility("default"))) void* uccl_b200_pool_malloc(size_t size, int device, cudaStream_t) {
  using namespace ub;
  std::shared_ptr<Comm> c = t_pool_override.lock();
  if (!c) {
    std::lock_guard<std::mutex> g(g_pool_mu);
    auto it = g_pool_by_device.find(device);
    if (it != g_pool_by_device.end()) c = it->second.lock();
  }
  void* p = nullptr;
  if (c && c->device() == device) {
    try {
      p = c->alloc(size ? size : 1, 512);
    } catch (const std::exception& e) {
      UB_WARN("mem pool: symmetric heap exhausted (%s); falling back to cudaMalloc for %zu bytes", e.what(), size);
    }
  }
  std::lock_guard<std::mutex> g(g_pool_mu);
  if (p) {
    g_owner[p] = Owner{c, size, true};
  } else {
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    if (cudaMalloc(&p, size ? size : 1) != cudaSuccess) p = nullptr;
    if (prev != device && prev >= 0) cudaSetDevice(prev);
    if (p) {
      g_owner[p] = Owner{std::weak_ptr<Comm>(), size, false};
      ++g_stats.fallback_allocs;
    }
  }
  if (p) {
    ++g_stats.allocs;
    g_stats.bytes_live += size;
  }
  return p;
}

__attribute__((visibility("default"))) void uccl_b200_pool_free(void* ptr, size_t, int device, cudaStream_t) {
  using namespace ub;
  if (!ptr) return;
  std::shared_ptr<Comm> c;
  bool heap_block = false;
  {
    std::lock_guard<std::mutex> g(g_pool_mu);
    auto it = g_owner.find(ptr);
    if (it == g_owner.end()) return;
    c = it->second.comm.lock();
    heap_block = it->second.heap;
    ++g_stats.frees;
    g_stats.bytes_live -= it->second.size;
    g_owner.erase(it);
  }
  if (heap_block) {
    if (c) c->free(ptr);  // communicator already gone: its heap went with it
    return;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  cudaFree(ptr);
  if (prev != device && prev >= 0) cudaSetDevice(prev);
}
}
