// Lazy CUDA driver-API table resolved through cudaGetDriverEntryPoint, so the
// shared object has no link-time dependency on libcuda.so.1 (it must import on
// the GPU-less build box) and so VMM / multicast symbols are looked up at the
// version the installed driver really provides.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../common/log.h"

namespace ub {

struct CuApi {
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*CtxGetCurrent)(CUcontext*) = nullptr;
  CUresult (*CtxGetDevice)(CUdevice*) = nullptr;
  CUresult (*StreamGetCtx)(CUstream, CUcontext*) = nullptr;
  CUresult (*StreamWriteValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
  CUresult (*StreamWaitValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*,
                                          CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                        unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*,
                                           CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t,
                               size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*,
                                      CUmulticastGranularity_flags) = nullptr;
  // green contexts (SM partitions, common/sm_partition.cc): optional, driver >= 12.4
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*,
                                        unsigned int, unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  CUresult (*GreenCtxGetDevResource)(CUgreenCtx, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  bool ok = false;
  bool green() const {
    return DeviceGetDevResource && DevSmResourceSplitByCount && DevResourceGenerateDesc && GreenCtxCreate &&
           GreenCtxDestroy && GreenCtxGetDevResource && GreenCtxStreamCreate;
  }
};

inline const CuApi& cu() {
  static CuApi* api = [] {
    auto* a = new CuApi();
    auto get = [&](const char* name, void** fn) -> bool {
      cudaDriverEntryPointQueryResult st;
      cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st);
      if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || *fn == nullptr) {
        (void)cudaGetLastError();
        *fn = nullptr;
        return false;
      }
      return true;
    };
    bool ok = true;
#define UB_GET(field, name) ok &= get(name, (void**)&a->field)
    UB_GET(GetErrorString, "cuGetErrorString");
    UB_GET(DeviceGet, "cuDeviceGet");
    UB_GET(DeviceGetAttribute, "cuDeviceGetAttribute");
    UB_GET(CtxGetCurrent, "cuCtxGetCurrent");
    UB_GET(MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
    UB_GET(MemCreate, "cuMemCreate");
    UB_GET(MemRelease, "cuMemRelease");
    UB_GET(MemAddressReserve, "cuMemAddressReserve");
    UB_GET(MemAddressFree, "cuMemAddressFree");
    UB_GET(MemMap, "cuMemMap");
    UB_GET(MemUnmap, "cuMemUnmap");
    UB_GET(MemSetAccess, "cuMemSetAccess");
    UB_GET(MemExportToShareableHandle, "cuMemExportToShareableHandle");
    UB_GET(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
    a->ok = ok;
    // multicast is optional (absent on drivers < 12.1 / non-NVSwitch systems)
    get("cuCtxGetDevice", (void**)&a->CtxGetDevice);
    get("cuStreamGetCtx", (void**)&a->StreamGetCtx);
    get("cuStreamWriteValue64", (void**)&a->StreamWriteValue64);
    get("cuStreamWaitValue64", (void**)&a->StreamWaitValue64);
    get("cuMulticastCreate", (void**)&a->MulticastCreate);
    get("cuMulticastAddDevice", (void**)&a->MulticastAddDevice);
    get("cuMulticastBindMem", (void**)&a->MulticastBindMem);
    get("cuMulticastUnbind", (void**)&a->MulticastUnbind);
    get("cuMulticastGetGranularity", (void**)&a->MulticastGetGranularity);
    get("cuDeviceGetDevResource", (void**)&a->DeviceGetDevResource);
    get("cuDevSmResourceSplitByCount", (void**)&a->DevSmResourceSplitByCount);
    get("cuDevResourceGenerateDesc", (void**)&a->DevResourceGenerateDesc);
    get("cuGreenCtxCreate", (void**)&a->GreenCtxCreate);
    get("cuGreenCtxDestroy", (void**)&a->GreenCtxDestroy);
    get("cuGreenCtxGetDevResource", (void**)&a->GreenCtxGetDevResource);
    get("cuGreenCtxStreamCreate", (void**)&a->GreenCtxStreamCreate);
#undef UB_GET
    return a;
  }();
  return *api;
}

inline const char* cu_errstr(CUresult r) {
  const char* s = nullptr;
  if (cu().GetErrorString) cu().GetErrorString(r, &s);
  return s ? s : "unknown CUresult";
}

}  // namespace ub

#define UB_CU(call)                                                        \
  do {                                                                     \
    CUresult _r = (call);                                                  \
    if (_r != CUDA_SUCCESS) UB_THROW("%s failed: %s (%d)", #call, ::ub::cu_errstr(_r), (int)_r); \
  } while (0)
#define UB_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) UB_THROW("%s failed: %s", #call, cudaGetErrorString(_e));    \
  } while (0)
