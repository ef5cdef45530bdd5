// NCCL-API GPU test: ncclCommInitAll over N ranks (real GPUs when available, otherwise virtual
// ranks on device 0), one thread per rank.  Covers AllReduce (cudaMalloc and ncclMemAlloc
// buffers), AllGather, ReduceScatter, Broadcast, grouped Send/Recv (ring + all-to-all pattern).
#include <cuda_runtime.h>
#include <nccl.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t ck_err_ = (x);                                                                 \
    if (ck_err_ != cudaSuccess) {                                                              \
      fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(ck_err_), __LINE__);                 \
      exit(2);                                                                           \
    }                                                                                    \
  } while (0)
#define NK(x)                                                                            \
  do {                                                                                   \
    ncclResult_t nk_res_ = (x);                                                                \
    if (nk_res_ != ncclSuccess) {                                                              \
      fprintf(stderr, "NCCL %s (%s) @%d\n", ncclGetErrorString(nk_res_), ncclGetLastError(nullptr), __LINE__); \
      exit(3);                                                                           \
    }                                                                                    \
  } while (0)

static std::atomic<int> g_fail{0};
#define EXPECT(c)                                                  \
  do {                                                             \
    if (!(c)) {                                                    \
      fprintf(stderr, "rank %d: FAILED %s @%d\n", r, #c, __LINE__); \
      g_fail++;                                                    \
    }                                                              \
  } while (0)

int main(int argc, char** argv) {
  setenv("UCCL_B200_TIMEOUT_MS", "8000", 0);
  setenv("UCCL_B200_MAX_CTAS", "4", 0);  // virtual ranks must be co-resident
  setenv("UCCL_B200_NCCL_HEAP_MB", "512", 0);
  setenv("UCCL_B200_NCCL_STAGE_MB", "16", 0);
  int n = argc > 1 ? atoi(argv[1]) : 4;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = ndev >= n ? i : 0;
  std::vector<ncclComm_t> comms(n);
  NK(ncclCommInitAll(comms.data(), n, devs.data()));
  std::vector<std::thread> ts;
  for (int r = 0; r < n; ++r)
    ts.emplace_back([&, r] {
      CK(cudaSetDevice(devs[r]));
      cudaStream_t st;
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      const size_t N = (1 << 20) + 3;
      std::vector<float> h(N), o(N);
      for (size_t i = 0; i < N; ++i) h[i] = (float)(i % 7) + r;
      float *d_in, *d_out;
      CK(cudaMalloc(&d_in, N * 4));
      CK(cudaMalloc(&d_out, N * 4));
      CK(cudaMemcpy(d_in, h.data(), N * 4, cudaMemcpyHostToDevice));
      CK(cudaStreamSynchronize(0));  // pageable H2D may still be in flight (null stream); `st` is non-blocking
      // 1. allreduce on plain buffers (staged + tail)
      NK(ncclAllReduce(d_in, d_out, N, ncclFloat, ncclSum, comms[r], st));
      CK(cudaStreamSynchronize(st));
      CK(cudaMemcpy(o.data(), d_out, N * 4, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < N; i += 4099) EXPECT(o[i] == n * (float)(i % 7) + n * (n - 1) / 2.0f);
      // 2. allreduce on ncclMemAlloc buffers (zero-copy path), small message (packet path)
      float* sym;
      NK(ncclMemAlloc((void**)&sym, 4096 * 4));
      CK(cudaMemcpy(sym, h.data(), 4096 * 4, cudaMemcpyHostToDevice));
      CK(cudaStreamSynchronize(0));  // pageable H2D may still be in flight (null stream); `st` is non-blocking
      NK(ncclAllReduce(sym, sym, 4096, ncclFloat, ncclMax, comms[r], st));
      CK(cudaStreamSynchronize(st));
      CK(cudaMemcpy(o.data(), sym, 4096 * 4, cudaMemcpyDeviceToHost));
      for (int i = 0; i < 4096; i += 97) EXPECT(o[i] == (float)(i % 7) + (n - 1));
      // 2b. PreMulSum (device-resident bf16-free path: float scalar on the host), allreduce + reduce_scatter
      {
        float scalar = 0.5f;
        ncclRedOp_t premul;
        NK(ncclRedOpCreatePreMulSum(&premul, &scalar, ncclFloat, ncclScalarHostImmediate, comms[r]));
        NK(ncclAllReduce(d_in, d_out, 4096, ncclFloat, premul, comms[r], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(o.data(), d_out, 4096 * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < 4096; i += 101) EXPECT(o[i] == 0.5f * (n * (float)(i % 7) + n * (n - 1) / 2.0f));
        NK(ncclReduceScatter(d_in, d_out, 1024, ncclFloat, premul, comms[r], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(o.data(), d_out, 1024 * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < 1024; i += 53) {
          const size_t gi = (size_t)r * 1024 + i;
          EXPECT(o[i] == 0.5f * (n * (float)(gi % 7) + n * (n - 1) / 2.0f));
        }
        NK(ncclRedOpDestroy(premul, comms[r]));
      }
      // 3. allgather / reduce_scatter
      const size_t P = 5000;
      float *ag_out, *rs_out;
      CK(cudaMalloc(&ag_out, P * n * 4));
      CK(cudaMalloc(&rs_out, P * 4));
      NK(ncclAllGather(d_in, ag_out, P, ncclFloat, comms[r], st));
      NK(ncclReduceScatter(ag_out, rs_out, P, ncclFloat, ncclSum, comms[r], st));
      CK(cudaStreamSynchronize(st));
      std::vector<float> g(P * n), rs(P);
      CK(cudaMemcpy(g.data(), ag_out, P * n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(rs.data(), rs_out, P * 4, cudaMemcpyDeviceToHost));
      for (int s = 0; s < n; ++s) EXPECT(g[s * P + 10] == (float)(10 % 7) + s);
      for (size_t i = 0; i < P; i += 13) EXPECT(rs[i] == n * ((float)(i % 7) + r));  // every rank holds the same gathered buffer
      // 4. broadcast from the last rank
      NK(ncclBroadcast(d_in, d_out, 1000, ncclFloat, n - 1, comms[r], st));
      CK(cudaStreamSynchronize(st));
      CK(cudaMemcpy(o.data(), d_out, 1000 * 4, cudaMemcpyDeviceToHost));
      EXPECT(o[5] == 5.0f + (n - 1));
      // 5. grouped send/recv: ring (send to r+1, recv from r-1) with a multi-chunk message
      const size_t M = (3 << 20) / 4 + 16;  // > one 512 KiB staging slot per block
      float *s_buf, *r_buf;
      CK(cudaMalloc(&s_buf, M * 4));
      CK(cudaMalloc(&r_buf, M * 4));
      std::vector<float> hs(M);
      for (size_t i = 0; i < M; ++i) hs[i] = (float)(r * 1000 + (i % 251));
      CK(cudaMemcpy(s_buf, hs.data(), M * 4, cudaMemcpyHostToDevice));
      CK(cudaStreamSynchronize(0));  // pageable H2D may still be in flight (null stream); `st` is non-blocking
      CK(cudaMemset(r_buf, 0, M * 4));
      CK(cudaStreamSynchronize(0));  // cudaMemset is asynchronous; `st` does not order with the null stream (no device-wide sync: virtual ranks share the GPU)
      NK(ncclGroupStart());
      NK(ncclSend(s_buf, M, ncclFloat, (r + 1) % n, comms[r], st));
      NK(ncclRecv(r_buf, M, ncclFloat, (r + n - 1) % n, comms[r], st));
      NK(ncclGroupEnd());
      CK(cudaStreamSynchronize(st));
      std::vector<float> hr(M);
      CK(cudaMemcpy(hr.data(), r_buf, M * 4, cudaMemcpyDeviceToHost));
      const int src = (r + n - 1) % n;
      for (size_t i = 0; i < M; i += 1009) {
        if (hr[i] != (float)(src * 1000 + (i % 251)) && g_fail.load() < 4)
          fprintf(stderr, "rank %d: ring recv [%zu] = %f, expected %f\n", r, i, hr[i], (float)(src * 1000 + (i % 251)));
        EXPECT(hr[i] == (float)(src * 1000 + (i % 251)));
      }
      EXPECT(hr[M - 1] == (float)(src * 1000 + ((M - 1) % 251)));
      // 6. all-to-all through grouped send/recv (incl. self), twice (sequence counters persist)
      const size_t C = 40000;
      float *a_in, *a_out;
      CK(cudaMalloc(&a_in, C * n * 4));
      CK(cudaMalloc(&a_out, C * n * 4));
      std::vector<float> ha(C * n);
      for (int rep = 0; rep < 2; ++rep) {
        for (int p = 0; p < n; ++p)
          for (size_t i = 0; i < C; ++i) ha[p * C + i] = (float)(rep * 7 + r * 100 + p);
        CK(cudaMemcpy(a_in, ha.data(), C * n * 4, cudaMemcpyHostToDevice));
        CK(cudaStreamSynchronize(0));  // pageable H2D may still be in flight (null stream); `st` is non-blocking
        NK(ncclGroupStart());
        for (int p = 0; p < n; ++p) {
          NK(ncclSend(a_in + p * C, C, ncclFloat, p, comms[r], st));
          NK(ncclRecv(a_out + p * C, C, ncclFloat, p, comms[r], st));
        }
        NK(ncclGroupEnd());
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(ha.data(), a_out, C * n * 4, cudaMemcpyDeviceToHost));
        for (int p = 0; p < n; ++p) {
          EXPECT(ha[p * C] == (float)(rep * 7 + p * 100 + r));
          EXPECT(ha[p * C + C - 1] == (float)(rep * 7 + p * 100 + r));
        }
      }
      NK(ncclMemFree(sym));
      cudaFree(d_in);
      cudaFree(d_out);
    });
  for (auto& t : ts) t.join();
  for (auto c : comms) ncclCommDestroy(c);
  if (g_fail.load()) {
    fprintf(stderr, "nccl_gpu_test: %d failures\n", g_fail.load());
    return 1;
  }
  printf("nccl_gpu_test: OK (%d ranks)\n", n);
  return 0;
}
