// nccl-tests style harness (all_reduce_perf / all_gather_perf / reduce_scatter_perf) against ANY
// libnccl-compatible library: link it with -luccl_b200_nccl (this repo) or -lnccl (baseline).
// One process, one thread per GPU (ncclCommInitAll), device-timed with CUDA events, max over
// ranks, data check against the closed-form expected value.
//   ./nccl_perf [-o allreduce|allgather|reducescatter] [-g ngpus] [-b 1K] [-e 1G] [-f 2] [-n iters] [-w warmup] [-c 1]
// (reference launchers: collective/rdma/run_nccl_test.sh:95-98, experimental/lite/scripts/run-nccl-tests.sh)
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CUDACHECK(x)                                                                      \
  do {                                                                                    \
    cudaError_t e = (x);                                                                  \
    if (e != cudaSuccess) {                                                               \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)
#define NCCLCHECK(x)                                                                      \
  do {                                                                                    \
    ncclResult_t r = (x);                                                                 \
    if (r != ncclSuccess) {                                                               \
      fprintf(stderr, "NCCL error %s at %s:%d\n", ncclGetErrorString(r), __FILE__, __LINE__); \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)

static size_t parse_size(const char* s) {
  char* end;
  double v = strtod(s, &end);
  if (*end == 'K' || *end == 'k') v *= 1024;
  if (*end == 'M' || *end == 'm') v *= 1024 * 1024;
  if (*end == 'G' || *end == 'g') v *= 1024.0 * 1024 * 1024;
  return (size_t)v;
}

struct Barrier {
  std::atomic<int> count{0}, gen{0};
  int n;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    int g = gen.load();
    if (count.fetch_add(1) + 1 == n) {
      count.store(0);
      gen.fetch_add(1);
    } else {
      while (gen.load() == g) std::this_thread::yield();
    }
  }
};

int main(int argc, char** argv) {
  std::string op = "allreduce";
  int ngpus = 0, iters = 20, warmup = 5, check = 1;
  size_t minb = 1024, maxb = 1ull << 30, factor = 2;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string a = argv[i];
    if (a == "-o") op = argv[i + 1];
    else if (a == "-g") ngpus = atoi(argv[i + 1]);
    else if (a == "-b") minb = parse_size(argv[i + 1]);
    else if (a == "-e") maxb = parse_size(argv[i + 1]);
    else if (a == "-f") factor = (size_t)atoi(argv[i + 1]);
    else if (a == "-n") iters = atoi(argv[i + 1]);
    else if (a == "-w") warmup = atoi(argv[i + 1]);
    else if (a == "-c") check = atoi(argv[i + 1]);
  }
  int ndev = 0;
  CUDACHECK(cudaGetDeviceCount(&ndev));
  if (ngpus <= 0 || ngpus > ndev) ngpus = ndev;
  std::vector<ncclComm_t> comms(ngpus);
  std::vector<int> devs(ngpus);
  for (int i = 0; i < ngpus; ++i) devs[i] = i;
  NCCLCHECK(ncclCommInitAll(comms.data(), ngpus, devs.data()));
  const int n = ngpus;
  printf("# %s  ngpus %d  float  sum\n# %12s %12s %10s %10s %10s %8s\n", op.c_str(), n, "size(B)", "count", "time(us)",
         "algbw", "busbw", "#wrong");
  std::vector<std::vector<float>> results;
  Barrier bar(n);
  std::vector<double> times(n);
  std::vector<long> wrong(n);
  for (size_t bytes = minb; bytes <= maxb; bytes *= factor) {
    const size_t count = bytes / sizeof(float);  // total elements of the "full" buffer
    const size_t per = std::max<size_t>(count / n, 1);
    std::vector<std::thread> ts;
    for (int r = 0; r < n; ++r)
      ts.emplace_back([&, r] {
        CUDACHECK(cudaSetDevice(r));
        cudaStream_t st;
        CUDACHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        float *in, *out;
        size_t in_n = op == "allgather" ? per : (op == "reducescatter" ? per * n : count);
        size_t out_n = op == "allgather" ? per * n : (op == "reducescatter" ? per : count);
        CUDACHECK(cudaMalloc(&in, in_n * sizeof(float)));
        CUDACHECK(cudaMalloc(&out, out_n * sizeof(float)));
        std::vector<float> h(in_n);
        for (size_t i = 0; i < in_n; ++i) h[i] = (float)((i % 13) + r);
        CUDACHECK(cudaMemcpy(in, h.data(), in_n * sizeof(float), cudaMemcpyHostToDevice));
        auto run = [&] {
          if (op == "allreduce") NCCLCHECK(ncclAllReduce(in, out, count, ncclFloat, ncclSum, comms[r], st));
          else if (op == "allgather") NCCLCHECK(ncclAllGather(in, out, per, ncclFloat, comms[r], st));
          else NCCLCHECK(ncclReduceScatter(in, out, per, ncclFloat, ncclSum, comms[r], st));
        };
        for (int i = 0; i < warmup; ++i) run();
        CUDACHECK(cudaStreamSynchronize(st));
        bar.wait();
        cudaEvent_t e0, e1;
        CUDACHECK(cudaEventCreate(&e0));
        CUDACHECK(cudaEventCreate(&e1));
        CUDACHECK(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; ++i) run();
        CUDACHECK(cudaEventRecord(e1, st));
        CUDACHECK(cudaStreamSynchronize(st));
        float ms = 0;
        CUDACHECK(cudaEventElapsedTime(&ms, e0, e1));
        times[r] = ms * 1e3 / iters;
        wrong[r] = 0;
        if (check) {
          std::vector<float> o(out_n);
          CUDACHECK(cudaMemcpy(o.data(), out, out_n * sizeof(float), cudaMemcpyDeviceToHost));
          for (size_t i = 0; i < out_n; ++i) {
            float exp;
            if (op == "allreduce") exp = (float)(n * (i % 13)) + n * (n - 1) / 2.0f;
            else if (op == "allgather") exp = (float)(((i % per) % 13) + (i / per));
            else exp = (float)(n * ((r * per + i) % 13)) + n * (n - 1) / 2.0f;
            if (std::fabs(o[i] - exp) > 1e-3f * std::fabs(exp) + 1e-3f) ++wrong[r];
          }
        }
        bar.wait();
        cudaFree(in);
        cudaFree(out);
        cudaStreamDestroy(st);
      });
    for (auto& t : ts) t.join();
    const double us = *std::max_element(times.begin(), times.end());
    long w = 0;
    for (long x : wrong) w += x;
    const size_t moved = op == "allreduce" ? count * sizeof(float) : per * n * sizeof(float);
    const double algbw = moved / us / 1e3;
    const double busbw = algbw * (op == "allreduce" ? 2.0 * (n - 1) / n : (double)(n - 1) / n);
    printf("  %12zu %12zu %10.1f %10.2f %10.2f %8ld\n", moved, moved / sizeof(float), us, algbw, busbw, w);
    fflush(stdout);
  }
  for (auto c : comms) ncclCommDestroy(c);
  return 0;
}
