// Symmetric-heap fabric: one VMM allocation per rank, mapped by every peer over
// NVLink, plus one NVLS multicast object bound over the same heap.  Every buffer
// inside the heap therefore has {local VA, N-1 peer VAs, multicast VA} at the same
// offset.  This single layer replaces the reference's per-op cudaIpc open/close
// (p2p/engine.cc:1732,1764), EP's IPC-shared cudaMalloc (ep/src/uccl_ep.cc:460-477)
// and lite's RegisteredMemory/CudaIpc/NVLS plumbing
// (experimental/lite/core/gpu_ipc_mem.cc:225-294,432-530).
//
// Three flavours share the interface:
//   * multi-process CUDA  : cuMemCreate(POSIX_FD) + fd passing + cuMulticast*
//   * single-process CUDA : all ranks in one process (ncclCommInitAll-style, and the
//                           1-GPU "virtual ranks" used by the GPU test-suite)
//   * host fake           : POSIX shm / malloc "heaps" so the whole control plane and the
//                           host reference collectives run on a GPU-less box (CI).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "bootstrap.h"

namespace ub {

constexpr int kFabricMaxRanks = 8;

class Fabric {
 public:
  static std::shared_ptr<Fabric> create(Bootstrap& bs, int device, size_t heap_bytes, size_t ctrl_bytes,
                                        bool host_fake);
  static std::vector<std::shared_ptr<Fabric>> create_local(const std::vector<int>& devices, size_t heap_bytes,
                                                           size_t ctrl_bytes, bool host_fake);
  ~Fabric();

  int rank() const { return rank_; }
  int nranks() const { return nranks_; }
  int device() const { return device_; }
  bool is_host() const { return host_; }
  bool single_process() const { return single_process_; }
  size_t heap_bytes() const { return heap_bytes_; }
  char* heap(int r) const { return heap_[r]; }
  char* local() const { return heap_[rank_]; }
  char* mc() const { return mc_; }
  bool has_multicast() const { return mc_ != nullptr; }
  bool contains(const void* p, size_t n) const {
    const char* c = (const char*)p;
    return c >= heap_[rank_] && c + n <= heap_[rank_] + heap_bytes_;
  }
  uint64_t offset_of(const void* p) const { return (uint64_t)((const char*)p - heap_[rank_]); }
  std::string describe() const;

 private:
  Fabric() = default;
  struct Shared;  // state shared by the ranks of a single-process world
  int rank_ = 0, nranks_ = 1, device_ = 0;
  bool host_ = false, single_process_ = false;
  size_t heap_bytes_ = 0;
  char* heap_[kFabricMaxRanks] = {nullptr};
  char* mc_ = nullptr;
  // ownership
  unsigned long long mem_handle_ = 0;
  unsigned long long peer_handles_[kFabricMaxRanks] = {0};
  unsigned long long mc_handle_ = 0;
  bool owns_mc_mapping_ = false;
  std::string shm_name_;
  std::shared_ptr<Shared> shared_;
};

}  // namespace ub
