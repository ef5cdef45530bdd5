#include "collectives_red.cuh"
namespace ub {
cudaError_t launch_red_i(int which, int dtype, int op, const DevComm& c, const CollArgs& a, int grid, int block,
                         cudaStream_t st) {
  switch (dtype) {
    case kI8: return launch_red_ops<int8_t>(which, op, false, c, a, grid, block, st);
    case kU8: return launch_red_ops<uint8_t>(which, op, false, c, a, grid, block, st);
    case kI32: return launch_red_ops<int32_t>(which, op, false, c, a, grid, block, st);
    case kU32: return launch_red_ops<uint32_t>(which, op, false, c, a, grid, block, st);
    case kI64: return launch_red_ops<int64_t>(which, op, false, c, a, grid, block, st);
    case kU64: return launch_red_ops<uint64_t>(which, op, false, c, a, grid, block, st);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_rs_ll_i(int dtype, int op, const DevComm& c, const CollArgs& a, int grid, int block,
                           cudaStream_t st) {
  switch (dtype) {
    case kI8: return launch_rs_ll_ops<int8_t>(op, c, a, grid, block, st);
    case kU8: return launch_rs_ll_ops<uint8_t>(op, c, a, grid, block, st);
    case kI32: return launch_rs_ll_ops<int32_t>(op, c, a, grid, block, st);
    case kU32: return launch_rs_ll_ops<uint32_t>(op, c, a, grid, block, st);
    case kI64: return launch_rs_ll_ops<int64_t>(op, c, a, grid, block, st);
    case kU64: return launch_rs_ll_ops<uint64_t>(op, c, a, grid, block, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace ub
