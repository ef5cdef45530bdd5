// nccl-tests style harness, one PROCESS per rank (ncclCommInitRank), for the NCCL-API drop-in or stock NCCL.
//   nccl_perf_mp -n <ranks> [-o allreduce|allgather|reducescatter|broadcast|alltoall] [-b 1K] [-e 64M] [-f 4]
//                [-i iters] [-w warmup] [-g gpus (ranks use device rank % gpus; 0 = host backend)]
// Ranks are forked from one launcher on this machine; with UCCL_B200_LOCAL_SIZE=L the drop-in treats every L
// consecutive ranks as one box and runs the hierarchical (NVLink + datagram rail) algorithms -- the way to
// look at the multi-box code path on a single machine.  On a real cluster start one launcher per host with
// -r <first rank> -N <total ranks> and share the id file (-x path).
// Reference role: collective/rdma/run_nccl_test.sh (all_reduce_perf / alltoall_perf / sendrecv_perf with MPI).
#include <cuda_runtime.h>
#include <nccl.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" ncclResult_t ncclAllToAll(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) __attribute__((weak));

static int g_rank = 0;
#define NC(x)                                                                                                   \
  do {                                                                                                          \
    ncclResult_t r_ = (x);                                                                                      \
    if (r_ != ncclSuccess) {                                                                                    \
      fprintf(stderr, "rank %d: %s: %s (%s)\n", g_rank, #x, ncclGetErrorString(r_), ncclGetLastError(nullptr)); \
      _exit(2);                                                                                                 \
    }                                                                                                           \
  } while (0)
#define CU(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      fprintf(stderr, "rank %d: %s: %s\n", g_rank, #x, cudaGetErrorString(e_));   \
      _exit(3);                                                                   \
    }                                                                             \
  } while (0)

static size_t parse_size(const char* s) {
  char* e = nullptr;
  double v = strtod(s, &e);
  if (*e == 'K' || *e == 'k') v *= 1 << 10;
  if (*e == 'M' || *e == 'm') v *= 1 << 20;
  if (*e == 'G' || *e == 'g') v *= 1 << 30;
  return (size_t)v;
}

struct Shared {  // launcher <-> ranks: per-size timing (max over ranks is taken by rank 0)
  double us[64][64];
  volatile int arrived[64];
};

int main(int argc, char** argv) {
  int n = 2, gpus = 0, iters = 20, warmup = 5, first = 0, total = -1;
  size_t minb = 1 << 10, maxb = 64 << 20, factor = 4;
  std::string op = "allreduce", idfile;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string a = argv[i];
    const char* v = argv[i + 1];
    if (a == "-n") n = atoi(v);
    else if (a == "-g") gpus = atoi(v);
    else if (a == "-o") op = v;
    else if (a == "-b") minb = parse_size(v);
    else if (a == "-e") maxb = parse_size(v);
    else if (a == "-f") factor = (size_t)atoi(v);
    else if (a == "-i") iters = atoi(v);
    else if (a == "-w") warmup = atoi(v);
    else if (a == "-r") first = atoi(v);
    else if (a == "-N") total = atoi(v);
    else if (a == "-x") idfile = v;
  }
  if (total < 0) total = n;
  if (gpus == 0) setenv("UCCL_B200_HOST_FAKE", "1", 1);
  ncclUniqueId id;
  if (first == 0) {
    if (ncclGetUniqueId(&id) != ncclSuccess) return 1;
    if (!idfile.empty()) {
      FILE* f = fopen(idfile.c_str(), "wb");
      fwrite(&id, sizeof(id), 1, f);
      fclose(f);
    }
  } else {
    FILE* f = nullptr;
    for (int t = 0; t < 600 && !(f = fopen(idfile.c_str(), "rb")); ++t) usleep(100000);
    if (!f || fread(&id, sizeof(id), 1, f) != 1) return 1;
    fclose(f);
  }
  Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  memset(sh, 0, sizeof(*sh));
  std::vector<pid_t> pids;
  int local = 0;
  for (int r = 1; r < n; ++r) {
    pid_t p = fork();
    if (p == 0) {
      local = r;
      pids.clear();
      break;
    }
    pids.push_back(p);
  }
  const int rank = first + local;
  g_rank = rank;
  const bool host = gpus == 0;
  cudaStream_t st = nullptr;
  if (!host) {
    CU(cudaSetDevice(local % gpus));
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  }
  ncclComm_t comm;
  NC(ncclCommInitRank(&comm, total, id, rank));
  const size_t bufb = maxb * (size_t)((op == "allgather" || op == "reducescatter" || op == "alltoall") ? total : 1);
  void *in = nullptr, *out = nullptr;
  if (host) {
    in = calloc(1, bufb);
    out = calloc(1, bufb);
  } else {
    CU(cudaMalloc(&in, bufb));
    CU(cudaMalloc(&out, bufb));
    CU(cudaMemset(in, 0, bufb));
  }
  if (local == 0 && first == 0)
    printf("# %s, %d ranks (%s), sizes %zu..%zu\n#%12s %12s %10s %10s\n", op.c_str(), total, host ? "host backend" : "cuda", minb,
           maxb, "bytes", "time_us", "algbw_GB/s", "busbw_GB/s");
  int si = 0;
  for (size_t b = minb; b <= maxb; b *= factor, ++si) {
    const size_t cnt = b / 4;
    auto once = [&] {
      if (op == "allreduce") NC(ncclAllReduce(in, out, cnt, ncclFloat, ncclSum, comm, st));
      else if (op == "allgather") NC(ncclAllGather(in, out, cnt, ncclFloat, comm, st));
      else if (op == "reducescatter") NC(ncclReduceScatter(in, out, cnt, ncclFloat, ncclSum, comm, st));
      else if (op == "broadcast") NC(ncclBroadcast(in, out, cnt, ncclFloat, 0, comm, st));
      else if (op == "alltoall" && ncclAllToAll) NC(ncclAllToAll(in, out, cnt, ncclFloat, comm, st));
      else {
        fprintf(stderr, "unknown op %s\n", op.c_str());
        _exit(4);
      }
    };
    for (int i = 0; i < warmup; ++i) once();
    if (!host) CU(cudaStreamSynchronize(st));
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) once();
    if (!host) CU(cudaStreamSynchronize(st));
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
    sh->us[si][local] = us;
    __sync_synchronize();
    __sync_fetch_and_add(&sh->arrived[si], 1);
    if (local == 0) {
      while (sh->arrived[si] < n) usleep(50);
      double mx = 0;
      for (int r = 0; r < n; ++r) mx = mx > sh->us[si][r] ? mx : sh->us[si][r];
      double factor_bus = 1.0;
      if (op == "allreduce") factor_bus = 2.0 * (total - 1) / total;
      else if (op == "allgather" || op == "reducescatter" || op == "alltoall") factor_bus = (double)(total - 1) / total;
      const double bytes = (op == "allgather" || op == "reducescatter" || op == "alltoall") ? (double)b * total : (double)b;
      const double alg = bytes / mx * 1e-3;
      if (first == 0) printf("%13zu %12.1f %10.3f %10.3f\n", b, mx, alg, alg * factor_bus);
      fflush(stdout);
    }
  }
  NC(ncclCommDestroy(comm));
  if (local != 0) _exit(0);
  int bad = 0;
  for (pid_t p : pids) {
    int s = 0;
    waitpid(p, &s, 0);
    bad += !(WIFEXITED(s) && WEXITSTATUS(s) == 0);
  }
  return bad ? 1 : 0;
}
