// Placement probe: which SM does every CTA of a launch land on?  Used to verify SM partitions
// (common/sm_partition.h; the reference's experimental/misc/cuda_greenctx.cu prints the same list) and by
// benchmarks/sm_partition_bench.py.  Every CTA holds its SM for `hold_ns` so that the grid spreads over all the SMs
// the stream may use instead of draining through the first few.
#include "launch.h"

namespace ub {

__global__ void __launch_bounds__(32) smid_probe_kernel(int* out, unsigned long long hold_ns) {
  unsigned int smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) out[blockIdx.x] = (int)smid;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < hold_ns);
}

cudaError_t launch_smid_probe(int* out, int blocks, unsigned long long hold_ns, cudaStream_t st) {
  if (blocks <= 0) return cudaErrorInvalidValue;
  UB_LAUNCH((smid_probe_kernel), blocks, 32, 0, st, out, hold_ns);
  return cudaGetLastError();
}

}  // namespace ub
