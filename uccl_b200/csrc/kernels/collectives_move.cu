// Data-movement collectives (no arithmetic): AllGather, Broadcast, AllToAll(v).
#include "collectives_impl.cuh"
#include "ll_exchange.cuh"
#include "launch.h"
namespace ub {
cudaError_t launch_allgather(int mode, const DevComm& c, const CollArgs& a, int grid, int block, cudaStream_t st) {
  switch (mode) {
    case 0: UB_LAUNCH((ag_kernel<0>), grid, block, 0, st, c, a); break;
    case 1: UB_LAUNCH((ag_kernel<1>), grid, block, 0, st, c, a); break;
    case 2: UB_LAUNCH((ag_kernel<2>), grid, block, 0, st, c, a); break;
    case 3: UB_LAUNCH((xchg_ll_kernel<0, false>), grid, block, 0, st, c, a); break;
    case 4: UB_LAUNCH((xchg_ll_kernel<0, true>), grid, block, 0, st, c, a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t launch_broadcast(int mode, const DevComm& c, const CollArgs& a, int grid, int block, cudaStream_t st) {
  switch (mode) {
    case 0: UB_LAUNCH((bcast_kernel<0>), grid, block, 0, st, c, a); break;
    case 1: UB_LAUNCH((bcast_kernel<1>), grid, block, 0, st, c, a); break;
    case 2: UB_LAUNCH((bcast_kernel<2>), grid, block, 0, st, c, a); break;
    case 3: UB_LAUNCH((bcast_kernel<3>), grid, block, 0, st, c, a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t launch_alltoall(int mode, const DevComm& c, const CollArgs& a, int grid, int block, cudaStream_t st) {
  switch (mode) {
    case 0: UB_LAUNCH((a2a_kernel<0>), grid, block, 0, st, c, a); break;
    case 1: UB_LAUNCH((a2a_kernel<1>), grid, block, 0, st, c, a); break;
    case 2: UB_LAUNCH((xchg_ll_kernel<1, false>), grid, block, 0, st, c, a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t launch_alltoallv(const DevComm& c, const CollArgs& a, const A2AvArgs& v, int grid, int block,
                             cudaStream_t st) {
  UB_LAUNCH((a2av_kernel), grid, block, 0, st, c, a, v);
  return cudaGetLastError();
}
}  // namespace ub

namespace ub {
cudaError_t launch_sendrecv(const DevComm& c, const SendRecvArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(kSrTmaStages + kSrSendStages) * kSrTmaChunk;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(sendrecv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done[dev & 63] = true;
  }
  UB_LAUNCH((sendrecv_kernel), (a.npeers > 0 ? a.npeers : 1) * kSrBlocks, 512, smem, st, c, a);
  return cudaGetLastError();
}
__global__ void barrier_kernel(const __grid_constant__ DevComm c, int domain) {
  BlockSync s = sync_begin(c, domain, blockIdx.x);
  sync_barrier(c, s);
  sync_end(s);
}
cudaError_t launch_barrier(const DevComm& c, int domain, cudaStream_t st) {
  UB_LAUNCH((barrier_kernel), 1, 32, 0, st, c, domain);
  return cudaGetLastError();
}
}  // namespace ub
