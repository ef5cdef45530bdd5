// AllReduce instantiations: integer dtypes.
#include <type_traits>
#include "allreduce_impl.cuh"
namespace ub {
cudaError_t launch_allreduce_i(int algo, int dtype, int op, const DevComm& c, const CollArgs& a, int grid,
                               int block, cudaStream_t st) {
  switch (dtype) {
    case kI8: return launch_ar_ops<int8_t>(algo, op, c, a, grid, block, st);
    case kU8: return launch_ar_ops<uint8_t>(algo, op, c, a, grid, block, st);
    case kI32: return launch_ar_ops<int32_t>(algo, op, c, a, grid, block, st);
    case kU32: return launch_ar_ops<uint32_t>(algo, op, c, a, grid, block, st);
    case kI64: return launch_ar_ops<int64_t>(algo, op, c, a, grid, block, st);
    case kU64: return launch_ar_ops<uint64_t>(algo, op, c, a, grid, block, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace ub
