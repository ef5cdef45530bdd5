#include "fabric.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <set>

#include "../common/log.h"
#include "../common/param.h"
#include "cu_api.h"

namespace ub {

UB_PARAM(NvlsEnable, "NVLS_ENABLE", 1)

namespace {

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

CUmemAllocationProp vmm_prop(int device, bool shareable) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = shareable ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : CU_MEM_HANDLE_TYPE_NONE;
  return prop;
}

bool device_supports_multicast(int device) {
  if (!ubParamNvlsEnable()) return false;
  if (!cu().MulticastCreate || !cu().MulticastAddDevice || !cu().MulticastBindMem) return false;
  CUdevice d;
  if (cu().DeviceGet(&d, device) != CUDA_SUCCESS) return false;
  int v = 0;
  if (cu().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d) != CUDA_SUCCESS) return false;
  return v != 0;
}

size_t heap_granularity(int device, int nranks_for_mc, bool want_mc, size_t size_hint) {
  CUmemAllocationProp prop = vmm_prop(device, true);
  size_t g = 2u << 20;
  UB_CU(cu().MemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (want_mc && cu().MulticastGetGranularity) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = nranks_for_mc;
    mp.size = round_up(size_hint, g);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (cu().MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g)
      g = mg;
  }
  return g;
}

void set_access(CUdeviceptr va, size_t size, const std::vector<int>& devices) {
  std::vector<CUmemAccessDesc> descs;
  std::set<int> uniq(devices.begin(), devices.end());
  for (int d : uniq) {
    CUmemAccessDesc a;
    memset(&a, 0, sizeof(a));
    a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    a.location.id = d;
    a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    descs.push_back(a);
  }
  UB_CU(cu().MemSetAccess(va, size, descs.data(), descs.size()));
}

}  // namespace

struct Fabric::Shared {
  // single-process world: rank 0's object owns the teardown of everything shared
  std::vector<CUmemGenericAllocationHandle> mem;
  std::vector<CUdeviceptr> va;
  CUmemGenericAllocationHandle mc = 0;
  CUdeviceptr mc_va = 0;
  size_t size = 0;
  std::vector<void*> host_heaps;
  bool host = false;
  ~Shared() {
    if (host) {
      for (void* p : host_heaps) free(p);
      return;
    }
    if (mc_va) {
      cu().MemUnmap(mc_va, size);
      cu().MemAddressFree(mc_va, size);
    }
    for (size_t i = 0; i < va.size(); ++i) {
      if (va[i]) {
        cu().MemUnmap(va[i], size);
        cu().MemAddressFree(va[i], size);
      }
      if (mem[i]) cu().MemRelease(mem[i]);
    }
    if (mc) cu().MemRelease(mc);
  }
};

std::vector<std::shared_ptr<Fabric>> Fabric::create_local(const std::vector<int>& devices, size_t heap_bytes,
                                                          size_t ctrl_bytes, bool host_fake) {
  const int n = (int)devices.size();
  UB_CHECK(n >= 1 && n <= kFabricMaxRanks, "local world size %d unsupported (1..%d)", n, kFabricMaxRanks);
  auto shared = std::make_shared<Shared>();
  shared->host = host_fake;
  std::vector<std::shared_ptr<Fabric>> out;
  if (host_fake) {
    shared->size = round_up(heap_bytes, 4096);
    for (int r = 0; r < n; ++r) {
      void* p = nullptr;
      UB_CHECK(posix_memalign(&p, 4096, shared->size) == 0, "host heap alloc failed");
      memset(p, 0, std::min(ctrl_bytes, shared->size));
      shared->host_heaps.push_back(p);
    }
    for (int r = 0; r < n; ++r) {
      std::shared_ptr<Fabric> f(new Fabric());
      f->rank_ = r;
      f->nranks_ = n;
      f->device_ = -1;
      f->host_ = true;
      f->single_process_ = true;
      f->heap_bytes_ = shared->size;
      for (int p = 0; p < n; ++p) f->heap_[p] = (char*)shared->host_heaps[p];
      f->shared_ = shared;
      out.push_back(f);
    }
    return out;
  }

  UB_CHECK(cu().ok, "CUDA driver VMM entry points unavailable");
  int prev_dev = 0;
  UB_CUDA(cudaGetDevice(&prev_dev));
  std::set<int> uniq(devices.begin(), devices.end());
  bool want_mc = n > 1 && (int)uniq.size() == n;
  for (int d : uniq) want_mc = want_mc && device_supports_multicast(d);
  // peer access between distinct devices
  for (int a : uniq)
    for (int b : uniq)
      if (a != b) {
        int can = 0;
        UB_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
        UB_CHECK(can, "device %d cannot access peer %d", a, b);
      }
  for (int d : uniq) {
    UB_CUDA(cudaSetDevice(d));
    UB_CUDA(cudaFree(0));
  }
  size_t gran = 0;
  for (int d : uniq) gran = std::max(gran, heap_granularity(d, n, want_mc, heap_bytes));
  const size_t size = round_up(heap_bytes, gran);
  shared->size = size;
  shared->mem.assign(n, 0);
  shared->va.assign(n, 0);
  for (int r = 0; r < n; ++r) {
    CUmemAllocationProp prop = vmm_prop(devices[r], false);
    UB_CU(cu().MemCreate(&shared->mem[r], size, &prop, 0));
    UB_CU(cu().MemAddressReserve(&shared->va[r], size, gran, 0, 0));
    UB_CU(cu().MemMap(shared->va[r], size, 0, shared->mem[r], 0));
    set_access(shared->va[r], size, devices);
    UB_CUDA(cudaSetDevice(devices[r]));
    UB_CUDA(cudaMemset((void*)shared->va[r], 0, std::min(ctrl_bytes, size)));
    UB_CUDA(cudaDeviceSynchronize());
  }
  if (want_mc) {
    try {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof(mp));
      mp.numDevices = n;
      mp.size = size;
      mp.handleTypes = 0;
      UB_CU(cu().MulticastCreate(&shared->mc, &mp));
      for (int r = 0; r < n; ++r) {
        CUdevice d;
        UB_CU(cu().DeviceGet(&d, devices[r]));
        UB_CU(cu().MulticastAddDevice(shared->mc, d));
      }
      for (int r = 0; r < n; ++r) UB_CU(cu().MulticastBindMem(shared->mc, 0, shared->mem[r], 0, size, 0));
      UB_CU(cu().MemAddressReserve(&shared->mc_va, size, gran, 0, 0));
      UB_CU(cu().MemMap(shared->mc_va, size, 0, shared->mc, 0));
      set_access(shared->mc_va, size, devices);
    } catch (const std::exception& e) {
      UB_WARN("NVLS multicast setup failed (%s); continuing with P2P only", e.what());
      if (shared->mc_va) {
        cu().MemAddressFree(shared->mc_va, size);
        shared->mc_va = 0;
      }
      if (shared->mc) {
        cu().MemRelease(shared->mc);
        shared->mc = 0;
      }
    }
  }
  for (int r = 0; r < n; ++r) {
    std::shared_ptr<Fabric> f(new Fabric());
    f->rank_ = r;
    f->nranks_ = n;
    f->device_ = devices[r];
    f->single_process_ = true;
    f->heap_bytes_ = size;
    for (int p = 0; p < n; ++p) f->heap_[p] = (char*)shared->va[p];
    f->mc_ = (char*)shared->mc_va;
    f->shared_ = shared;
    out.push_back(f);
  }
  UB_CUDA(cudaSetDevice(prev_dev));
  UB_INFO(SUB_FABRIC, "local world: %d ranks, heap %zu MiB each, multicast=%d", n, size >> 20,
          shared->mc_va ? 1 : 0);
  return out;
}

std::shared_ptr<Fabric> Fabric::create(Bootstrap& bs, int device, size_t heap_bytes, size_t ctrl_bytes,
                                       bool host_fake) {
  const int n = bs.nranks(), rank = bs.rank();
  UB_CHECK(n >= 1 && n <= kFabricMaxRanks, "world size %d unsupported (1..%d)", n, kFabricMaxRanks);
  std::shared_ptr<Fabric> f(new Fabric());
  f->rank_ = rank;
  f->nranks_ = n;
  f->device_ = device;
  f->host_ = host_fake;

  if (host_fake) {
    const size_t size = round_up(heap_bytes, 4096);
    f->heap_bytes_ = size;
    char name[96];
    auto mk = [&](int r) {
      snprintf(name, sizeof(name), "/ub_%016lx_%d", (unsigned long)bs.nonce(), r);
      return std::string(name);
    };
    f->shm_name_ = mk(rank);
    int fd = shm_open(f->shm_name_.c_str(), O_CREAT | O_RDWR | O_EXCL, 0600);
    UB_CHECK(fd >= 0, "shm_open(%s) failed: %s", f->shm_name_.c_str(), strerror(errno));
    UB_CHECK(ftruncate(fd, (off_t)size) == 0, "ftruncate failed: %s", strerror(errno));
    void* p = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    UB_CHECK(p != MAP_FAILED, "mmap failed: %s", strerror(errno));
    memset(p, 0, std::min(ctrl_bytes, size));
    f->heap_[rank] = (char*)p;
    bs.barrier();
    for (int r = 0; r < n; ++r) {
      if (r == rank) continue;
      int pfd = shm_open(mk(r).c_str(), O_RDWR, 0600);
      UB_CHECK(pfd >= 0, "shm_open(peer %d) failed: %s", r, strerror(errno));
      void* q = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, pfd, 0);
      ::close(pfd);
      UB_CHECK(q != MAP_FAILED, "mmap(peer %d) failed", r);
      f->heap_[r] = (char*)q;
    }
    bs.barrier();
    shm_unlink(f->shm_name_.c_str());  // mappings stay alive; name is gone even if we crash later
    return f;
  }

  UB_CHECK(cu().ok, "CUDA driver VMM entry points unavailable");
  UB_CUDA(cudaSetDevice(device));
  UB_CUDA(cudaFree(0));
  // consensus on devices / multicast capability
  struct Info {
    char busid[32];
    int mc_ok;
    int pid;
  } mine, all[kFabricMaxRanks];
  memset(&mine, 0, sizeof(mine));
  UB_CUDA(cudaDeviceGetPCIBusId(mine.busid, sizeof(mine.busid), device));
  mine.mc_ok = device_supports_multicast(device) ? 1 : 0;
  mine.pid = (int)getpid();
  bs.allgather(&mine, all, sizeof(Info));
  bool want_mc = n > 1;
  std::set<std::string> busids;
  for (int r = 0; r < n; ++r) {
    want_mc = want_mc && all[r].mc_ok;
    busids.insert(all[r].busid);
  }
  if ((int)busids.size() != n) want_mc = false;  // several ranks share a GPU (test mode)

  size_t gran = heap_granularity(device, n, want_mc, heap_bytes);
  uint64_t g64 = gran, gmax[kFabricMaxRanks];
  bs.allgather(&g64, gmax, sizeof(uint64_t));
  for (int r = 0; r < n; ++r) gran = std::max<size_t>(gran, gmax[r]);
  const size_t size = round_up(heap_bytes, gran);
  f->heap_bytes_ = size;

  CUmemAllocationProp prop = vmm_prop(device, true);
  CUmemGenericAllocationHandle mem = 0;
  UB_CU(cu().MemCreate(&mem, size, &prop, 0));
  f->mem_handle_ = mem;
  CUdeviceptr va = 0;
  UB_CU(cu().MemAddressReserve(&va, size, gran, 0, 0));
  UB_CU(cu().MemMap(va, size, 0, mem, 0));
  set_access(va, size, {device});
  f->heap_[rank] = (char*)va;
  UB_CUDA(cudaMemset((void*)va, 0, std::min(ctrl_bytes, size)));
  UB_CUDA(cudaDeviceSynchronize());

  if (n > 1) {
    int mem_fd = -1;
    UB_CU(cu().MemExportToShareableHandle(&mem_fd, mem, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    // multicast object: created by rank 0, fd shipped with the heap fds
    CUmemGenericAllocationHandle mc = 0;
    int mc_fd = -1;
    int mc_status = want_mc ? 1 : 0;
    if (want_mc && rank == 0) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof(mp));
      mp.numDevices = n;
      mp.size = size;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      CUresult r1 = cu().MulticastCreate(&mc, &mp);
      if (r1 == CUDA_SUCCESS)
        r1 = cu().MemExportToShareableHandle(&mc_fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
      if (r1 != CUDA_SUCCESS) {
        UB_WARN("cuMulticastCreate/export failed: %s; NVLS disabled", cu_errstr(r1));
        mc_status = 0;
        if (mc) cu().MemRelease(mc);
        mc = 0;
      }
    }
    bs.broadcast(&mc_status, sizeof(mc_status), 0);
    want_mc = mc_status != 0;
    std::vector<int> offer = {mem_fd, (want_mc && rank == 0) ? mc_fd : mem_fd};
    auto fds = bs.exchange_fds(offer);
    for (int r = 0; r < n; ++r) {
      if (r == rank) continue;
      CUmemGenericAllocationHandle h = 0;
      UB_CU(cu().MemImportFromShareableHandle(&h, (void*)(uintptr_t)fds[r][0],
                                              CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      f->peer_handles_[r] = h;
      CUdeviceptr pva = 0;
      UB_CU(cu().MemAddressReserve(&pva, size, gran, 0, 0));
      UB_CU(cu().MemMap(pva, size, 0, h, 0));
      set_access(pva, size, {device});
      f->heap_[r] = (char*)pva;
    }
    if (want_mc) {
      int ok = 1;
      if (rank != 0) {
        CUresult r2 = cu().MemImportFromShareableHandle(&mc, (void*)(uintptr_t)fds[0][1],
                                                        CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        if (r2 != CUDA_SUCCESS) ok = 0;
      }
      CUdevice d;
      if (ok && (cu().DeviceGet(&d, device) != CUDA_SUCCESS || cu().MulticastAddDevice(mc, d) != CUDA_SUCCESS))
        ok = 0;
      int oks[kFabricMaxRanks];
      bs.allgather(&ok, oks, sizeof(int));  // also the "all devices added" barrier
      for (int r = 0; r < n; ++r) ok = ok && oks[r];
      if (ok) {
        if (cu().MulticastBindMem(mc, 0, mem, 0, size, 0) != CUDA_SUCCESS) ok = 0;
        bs.allgather(&ok, oks, sizeof(int));
        for (int r = 0; r < n; ++r) ok = ok && oks[r];
      }
      if (ok) {
        CUdeviceptr mva = 0;
        if (cu().MemAddressReserve(&mva, size, gran, 0, 0) == CUDA_SUCCESS &&
            cu().MemMap(mva, size, 0, mc, 0) == CUDA_SUCCESS) {
          try {
            set_access(mva, size, {device});
            f->mc_ = (char*)mva;
            f->owns_mc_mapping_ = true;
          } catch (...) {
            ok = 0;
          }
        } else {
          ok = 0;
        }
        bs.allgather(&ok, oks, sizeof(int));
        for (int r = 0; r < n; ++r) ok = ok && oks[r];
        if (!ok) f->mc_ = nullptr;
      }
      f->mc_handle_ = mc;
      if (!ok) UB_WARN("NVLS multicast bind/map failed on some rank; continuing with P2P only");
    }
    for (auto& v : fds)
      for (int fd : v)
        if (fd >= 0) ::close(fd);
    ::close(mem_fd);
    if (mc_fd >= 0) ::close(mc_fd);
  }
  bs.barrier();
  UB_INFO(SUB_FABRIC, "rank %d/%d dev %d: heap %zu MiB, multicast=%d", rank, n, device, size >> 20,
          f->mc_ ? 1 : 0);
  return f;
}

Fabric::~Fabric() {
  if (single_process_) return;  // Shared dtor owns everything
  if (host_) {
    for (int r = 0; r < nranks_; ++r)
      if (heap_[r]) munmap(heap_[r], heap_bytes_);
    return;
  }
  if (!cu().ok) return;
  if (mc_ && owns_mc_mapping_) {
    cu().MemUnmap((CUdeviceptr)mc_, heap_bytes_);
    cu().MemAddressFree((CUdeviceptr)mc_, heap_bytes_);
  }
  if (mc_handle_) {
    CUdevice d;
    if (cu().MulticastUnbind && cu().DeviceGet(&d, device_) == CUDA_SUCCESS)
      cu().MulticastUnbind(mc_handle_, d, 0, heap_bytes_);
    cu().MemRelease(mc_handle_);
  }
  for (int r = 0; r < nranks_; ++r) {
    if (!heap_[r]) continue;
    cu().MemUnmap((CUdeviceptr)heap_[r], heap_bytes_);
    cu().MemAddressFree((CUdeviceptr)heap_[r], heap_bytes_);
    if (r != rank_ && peer_handles_[r]) cu().MemRelease(peer_handles_[r]);
  }
  if (mem_handle_) cu().MemRelease(mem_handle_);
}

std::string Fabric::describe() const {
  char b[256];
  snprintf(b, sizeof(b), "Fabric(rank=%d/%d, device=%d, heap=%zuMiB, multicast=%d, host=%d, single_process=%d)",
           rank_, nranks_, device_, heap_bytes_ >> 20, mc_ ? 1 : 0, host_ ? 1 : 0, single_process_ ? 1 : 0);
  return std::string(b);
}

}  // namespace ub
