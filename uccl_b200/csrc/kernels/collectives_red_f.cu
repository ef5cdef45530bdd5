#include "collectives_red.cuh"
namespace ub {
// which: 0 = reduce_scatter, 1 = reduce
cudaError_t launch_red_f(int which, int dtype, int op, bool nvls, const DevComm& c, const CollArgs& a, int grid,
                         int block, cudaStream_t st) {
  switch (dtype) {
    case kF32: return launch_red_ops<float>(which, op, nvls, c, a, grid, block, st);
    case kBF16: return launch_red_ops<__nv_bfloat16>(which, op, nvls, c, a, grid, block, st);
    case kF16: return launch_red_ops<__half>(which, op, nvls, c, a, grid, block, st);
    case kF64: return launch_red_ops<double>(which, op, nvls, c, a, grid, block, st);
    case kF8E4M3: return launch_red_ops<__nv_fp8_e4m3>(which, op, nvls, c, a, grid, block, st);
    case kF8E5M2: return launch_red_ops<__nv_fp8_e5m2>(which, op, nvls, c, a, grid, block, st);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_rs_ll_f(int dtype, int op, const DevComm& c, const CollArgs& a, int grid, int block,
                           cudaStream_t st) {
  switch (dtype) {
    case kF32: return launch_rs_ll_ops<float>(op, c, a, grid, block, st);
    case kBF16: return launch_rs_ll_ops<__nv_bfloat16>(op, c, a, grid, block, st);
    case kF16: return launch_rs_ll_ops<__half>(op, c, a, grid, block, st);
    case kF64: return launch_rs_ll_ops<double>(op, c, a, grid, block, st);
    case kF8E4M3: return launch_rs_ll_ops<__nv_fp8_e4m3>(op, c, a, grid, block, st);
    case kF8E5M2: return launch_rs_ll_ops<__nv_fp8_e5m2>(op, c, a, grid, block, st);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace ub
