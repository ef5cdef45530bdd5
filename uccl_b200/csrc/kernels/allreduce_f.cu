// AllReduce instantiations: floating dtypes (+ fused output-cast variants).
#include <type_traits>
#include "allreduce_impl.cuh"
namespace ub {
cudaError_t launch_allreduce_f(int algo, int dtype, int op, int out_dtype, const DevComm& c, const CollArgs& a,
                               int grid, int block, cudaStream_t st) {
  if (out_dtype == dtype) {
    switch (dtype) {
      case kF32: return launch_ar_ops<float>(algo, op, c, a, grid, block, st);
      case kBF16: return launch_ar_ops<__nv_bfloat16>(algo, op, c, a, grid, block, st);
      case kF16: return launch_ar_ops<__half>(algo, op, c, a, grid, block, st);
      default: return cudaErrorInvalidValue;
    }
  }
  // fused cast: sum/avg only, zero-copy and staged algorithms only
  if (op != kSum && op != kAvg) return cudaErrorInvalidValue;
  if (dtype == kF32 && out_dtype == kBF16) return launch_ar_typed<float, kSum, __nv_bfloat16>(algo, c, a, grid, block, st);
  if (dtype == kF32 && out_dtype == kF16) return launch_ar_typed<float, kSum, __half>(algo, c, a, grid, block, st);
  if (dtype == kBF16 && out_dtype == kF32) return launch_ar_typed<__nv_bfloat16, kSum, float>(algo, c, a, grid, block, st);
  if (dtype == kF16 && out_dtype == kF32) return launch_ar_typed<__half, kSum, float>(algo, c, a, grid, block, st);
  return cudaErrorInvalidValue;
}
}  // namespace ub
