// AllReduce instantiations: fp64 and fp8 dtypes.
#include <type_traits>
#include "allreduce_impl.cuh"
namespace ub {
cudaError_t launch_allreduce_x(int algo, int dtype, int op, const DevComm& c, const CollArgs& a, int grid,
                               int block, cudaStream_t st) {
  switch (dtype) {
    case kF64: return launch_ar_ops<double>(algo, op, c, a, grid, block, st);
    case kF8E4M3: return launch_ar_ops<__nv_fp8_e4m3>(algo, op, c, a, grid, block, st);
    case kF8E5M2: return launch_ar_ops<__nv_fp8_e5m2>(algo, op, c, a, grid, block, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace ub
