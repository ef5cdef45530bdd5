#include "sm_partition.h"

#include "../fabric/cu_api.h"

namespace ub {

namespace {
struct DevScope {
  int prev = -1;
  explicit DevScope(int d) {
    cudaGetDevice(&prev);
    if (prev != d) cudaSetDevice(d);
  }
  ~DevScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
}  // namespace

bool SmPartition::supported(int device, std::string* why) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    (void)cudaGetLastError();
    if (why) *why = "no such CUDA device";
    return false;
  }
  if (!cu().green()) {
    if (why) *why = "the CUDA driver does not provide green contexts (needs 12.4 or newer)";
    return false;
  }
  return true;
}

int SmPartition::device_sm_count(int device) {
  std::string why;
  UB_CHECK(supported(device, &why), "SM partitions unavailable: %s", why.c_str());
  DevScope g(device);
  UB_CUDA(cudaFree(nullptr));  // the primary context must exist before its resources can be queried
  CUdevice dev;
  UB_CU(cu().DeviceGet(&dev, device));
  CUdevResource all;
  UB_CU(cu().DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  return (int)all.sm.smCount;
}

std::shared_ptr<SmPartition> SmPartition::make(int device, void* resource) {
  const CuApi& d = cu();
  auto* res = static_cast<CUdevResource*>(resource);
  CUdevice dev;
  UB_CU(d.DeviceGet(&dev, device));
  CUdevResourceDesc desc;
  UB_CU(d.DevResourceGenerateDesc(&desc, res, 1));
  CUgreenCtx g = nullptr;
  UB_CU(d.GreenCtxCreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  std::shared_ptr<SmPartition> p(new SmPartition());
  p->device_ = device;
  p->green_ = g;
  CUdevResource got;
  UB_CU(d.GreenCtxGetDevResource(g, &got, CU_DEV_RESOURCE_TYPE_SM));
  p->sm_count_ = (int)got.sm.smCount;
  return p;
}

SmPartition::Pair SmPartition::split(int device, int sm_count, bool fine_grained) {
  std::string why;
  UB_CHECK(supported(device, &why), "SM partitions unavailable: %s", why.c_str());
  const CuApi& d = cu();
  DevScope g(device);
  UB_CUDA(cudaFree(nullptr));
  CUdevice dev;
  UB_CU(d.DeviceGet(&dev, device));
  CUdevResource all;
  UB_CU(d.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  UB_CHECK(sm_count > 0 && sm_count <= (int)all.sm.smCount, "SM partition of %d SMs on a device with %u", sm_count,
           all.sm.smCount);
  CUdevResource grp, rem;
  unsigned int nb = 1;
  const unsigned int flags = fine_grained ? CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING : 0;
  UB_CU(d.DevSmResourceSplitByCount(&grp, &nb, &all, &rem, flags, (unsigned int)sm_count));
  UB_CHECK(nb == 1, "the device could not be split into a partition of %d SMs", sm_count);
  Pair out;
  out.part = make(device, &grp);
  if (rem.type == CU_DEV_RESOURCE_TYPE_SM && rem.sm.smCount > 0) out.rest = make(device, &rem);
  UB_INFO(SUB_UTIL, "SM partition on GPU %d: %d SMs (asked for %d), rest %d of %u", device, out.part->sm_count(),
              sm_count, out.rest ? out.rest->sm_count() : 0, all.sm.smCount);
  return out;
}

cudaStream_t SmPartition::stream(int priority) {
  std::lock_guard<std::mutex> l(mu_);
  for (auto& s : streams_)
    if (s.first == priority) return s.second;
  CUstream s = nullptr;
  UB_CU(cu().GreenCtxStreamCreate(&s, static_cast<CUgreenCtx>(green_), CU_STREAM_NON_BLOCKING, priority));
  streams_.emplace_back(priority, reinterpret_cast<cudaStream_t>(s));
  return reinterpret_cast<cudaStream_t>(s);
}

SmPartition::~SmPartition() {
  for (auto& s : streams_) {
    cudaStreamSynchronize(s.second);
    cudaStreamDestroy(s.second);
  }
  if (green_ && cu().GreenCtxDestroy) cu().GreenCtxDestroy(static_cast<CUgreenCtx>(green_));
  (void)cudaGetLastError();
}

}  // namespace ub
