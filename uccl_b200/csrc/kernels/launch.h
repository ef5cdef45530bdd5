// Host-callable launchers of every kernel (implemented in the .cu translation units).
#pragma once
#include <cuda_runtime.h>

#include "types.h"

namespace ub {
// allreduce: algo ids = ArAlgo in allreduce_impl.cuh (same numbering as ArAlgoId in comm.h)
cudaError_t launch_allreduce_f(int algo, int dtype, int op, int out_dtype, const DevComm& c, const CollArgs& a,
                               int grid, int block, cudaStream_t st);
cudaError_t launch_allreduce_i(int algo, int dtype, int op, const DevComm& c, const CollArgs& a, int grid,
                               int block, cudaStream_t st);
cudaError_t launch_allreduce_x(int algo, int dtype, int op, const DevComm& c, const CollArgs& a, int grid,
                               int block, cudaStream_t st);
cudaError_t launch_allgather(int mode, const DevComm& c, const CollArgs& a, int grid, int block, cudaStream_t st);
cudaError_t launch_broadcast(int mode, const DevComm& c, const CollArgs& a, int grid, int block, cudaStream_t st);
cudaError_t launch_alltoall(int mode, const DevComm& c, const CollArgs& a, int grid, int block, cudaStream_t st);
cudaError_t launch_alltoallv(const DevComm& c, const CollArgs& a, const A2AvArgs& v, int grid, int block,
                             cudaStream_t st);
// which: 0 = reduce_scatter, 1 = reduce
cudaError_t launch_red_f(int which, int dtype, int op, bool nvls, const DevComm& c, const CollArgs& a, int grid,
                         int block, cudaStream_t st);
cudaError_t launch_red_i(int which, int dtype, int op, const DevComm& c, const CollArgs& a, int grid, int block,
                         cudaStream_t st);
cudaError_t launch_barrier(const DevComm& c, int domain, cudaStream_t st);
}  // namespace ub
