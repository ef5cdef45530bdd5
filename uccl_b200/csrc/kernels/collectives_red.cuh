// Typed launch helpers for ReduceScatter / Reduce.
#pragma once
#include "launch.h"
#include "collectives_impl.cuh"
#include "ll_exchange.cuh"
namespace ub {
template <typename T, int OP>
cudaError_t launch_red_typed(int which, bool nvls, const DevComm& c, const CollArgs& a, int grid, int block,
                             cudaStream_t st) {
  if (nvls) {
    if constexpr (MmLdRed<T, OP>::ok) {
      if (which == 0) UB_LAUNCH((rs_kernel<T, OP, true>), grid, block, 0, st, c, a);
      else UB_LAUNCH((reduce_kernel<T, OP, true>), grid, block, 0, st, c, a);
      return cudaGetLastError();
    } else {
      return cudaErrorInvalidValue;
    }
  }
  if (which == 0) UB_LAUNCH((rs_kernel<T, OP, false>), grid, block, 0, st, c, a);
  else UB_LAUNCH((reduce_kernel<T, OP, false>), grid, block, 0, st, c, a);
  return cudaGetLastError();
}
template <typename T>
cudaError_t launch_red_ops(int which, int op, bool nvls, const DevComm& c, const CollArgs& a, int grid, int block,
                           cudaStream_t st) {
  switch (op) {
    case kSum: case kAvg: return launch_red_typed<T, kSum>(which, nvls, c, a, grid, block, st);
    case kProd: return launch_red_typed<T, kProd>(which, nvls, c, a, grid, block, st);
    case kMax: return launch_red_typed<T, kMax>(which, nvls, c, a, grid, block, st);
    case kMin: return launch_red_typed<T, kMin>(which, nvls, c, a, grid, block, st);
    default: return cudaErrorInvalidValue;
  }
}
template <typename T>
cudaError_t launch_rs_ll_ops(int op, const DevComm& c, const CollArgs& a, int grid, int block, cudaStream_t st) {
  switch (op) {
    case kSum: case kAvg: UB_LAUNCH((rs_ll_kernel<T, kSum>), grid, block, 0, st, c, a); break;
    case kProd: UB_LAUNCH((rs_ll_kernel<T, kProd>), grid, block, 0, st, c, a); break;
    case kMax: UB_LAUNCH((rs_ll_kernel<T, kMax>), grid, block, 0, st, c, a); break;
    case kMin: UB_LAUNCH((rs_ll_kernel<T, kMin>), grid, block, 0, st, c, a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
}  // namespace ub
