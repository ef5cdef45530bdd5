// Kernels around the GPU->CPU command queue: throughput / latency microbenchmarks (role of the
// reference's ep/src/bench_kernel.cu:11,143 `gpu_issue_batched_commands`) and the tiny helpers the
// proxy launches.
#include "../kernels/launch.h"
#include "d2h_queue.cuh"
#include "proxy.h"

namespace ub {

__global__ void d2h_bench_kernel(const __grid_constant__ D2HQueueDev q, int per_thread) {
  for (int i = 0; i < per_thread; ++i) d2h_push(q, D2H_NOP, 0, 0, 0, 0, 0, (uint32_t)i);
}

// one thread: push a NOP, wait for the proxy's acknowledgement, repeat; accumulates round-trip time
__global__ void d2h_latency_kernel(const __grid_constant__ D2HQueueDev q, int iters, unsigned long long* total_ns) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long sum = 0;
  uint64_t acked = *q.ack;
  for (int i = 0; i < iters; ++i) {
    const uint64_t t0 = globaltimer_ns();
    d2h_push(q, D2H_NOP, 0, 0, 0, 0, 0, 0);
    ++acked;
    while (*q.ack < acked) {
    }
    sum += globaltimer_ns() - t0;
  }
  *total_ns = sum;
}

__global__ void d2h_issue_kernel(const __grid_constant__ D2HQueueDev q, uint32_t type, uint32_t dst_rank, uint32_t aux,
                                 uint64_t src_off, uint64_t dst_off, uint32_t bytes, uint32_t value) {
  if (threadIdx.x == 0 && blockIdx.x == 0) d2h_push(q, type, dst_rank, aux, src_off, dst_off, bytes, value);
}

__global__ void u64_add_kernel(uint64_t* p, uint64_t v) {
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

cudaError_t launch_d2h_bench(const D2HQueueDev& q, int blocks, int threads, int per_thread, cudaStream_t st) {
  UB_LAUNCH((d2h_bench_kernel), blocks, threads, 0, st, q, per_thread);
  return cudaGetLastError();
}
cudaError_t launch_d2h_latency(const D2HQueueDev& q, int iters, unsigned long long* total_ns, cudaStream_t st) {
  UB_LAUNCH((d2h_latency_kernel), 1, 32, 0, st, q, iters, total_ns);
  return cudaGetLastError();
}
cudaError_t launch_d2h_issue(const D2HQueueDev& q, uint32_t type, uint32_t dst_rank, uint32_t aux, uint64_t src_off,
                             uint64_t dst_off, uint32_t bytes, uint32_t value, cudaStream_t st) {
  UB_LAUNCH((d2h_issue_kernel), 1, 32, 0, st, q, type, dst_rank, aux, src_off, dst_off, bytes, value);
  return cudaGetLastError();
}
cudaError_t launch_u64_add(uint64_t* p, uint64_t v, cudaStream_t st) {
  UB_LAUNCH((u64_add_kernel), 1, 1, 0, st, p, v);
  return cudaGetLastError();
}

}  // namespace ub
