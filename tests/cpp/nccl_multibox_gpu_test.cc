// NCCL API across "boxes" with device buffers: N processes, each a box of ONE GPU rank (UCCL_B200_LOCAL_SIZE=1,
// all on device 0), datagram rails between them.  Exercises the pinned-staging path of MultiComm.
//   nccl_multibox_gpu_test [nranks=2]
#include <cuda_runtime.h>
#include <nccl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

static int g_rank = -1;
#define CHECK(x)                                                                                                       \
  do {                                                                                                                 \
    ncclResult_t _r = (x);                                                                                             \
    if (_r != ncclSuccess) {                                                                                           \
      fprintf(stderr, "rank %d: %s failed: %s (%s)\n", g_rank, #x, ncclGetErrorString(_r), ncclGetLastError(nullptr)); \
      exit(2);                                                                                                         \
    }                                                                                                                  \
  } while (0)
#define CU(x)                                                                              \
  do {                                                                                     \
    cudaError_t _e = (x);                                                                  \
    if (_e != cudaSuccess) {                                                               \
      fprintf(stderr, "rank %d: %s: %s\n", g_rank, #x, cudaGetErrorString(_e));            \
      exit(4);                                                                             \
    }                                                                                      \
  } while (0)
#define EXPECT(c)                                                                           \
  do {                                                                                      \
    if (!(c)) {                                                                             \
      fprintf(stderr, "rank %d: expectation failed: %s (line %d)\n", g_rank, #c, __LINE__); \
      exit(3);                                                                              \
    }                                                                                       \
  } while (0)

static int run(int rank, int n, ncclUniqueId id) {
  g_rank = rank;
  CU(cudaSetDevice(0));
  cudaStream_t st;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  ncclComm_t comm;
  CHECK(ncclCommInitRank(&comm, n, id, rank));
  const size_t N = 1 << 20;
  std::vector<float> h(N);
  for (size_t i = 0; i < N; ++i) h[i] = (float)(i % 13) + rank;
  float *d_in, *d_out;
  CU(cudaMalloc(&d_in, N * 4));
  CU(cudaMalloc(&d_out, N * 4 * n));
  CU(cudaMemcpyAsync(d_in, h.data(), N * 4, cudaMemcpyHostToDevice, st));
  CHECK(ncclAllReduce(d_in, d_out, N, ncclFloat, ncclSum, comm, st));
  std::vector<float> r(N * n);
  CU(cudaMemcpyAsync(r.data(), d_out, N * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (size_t i = 0; i < N; i += 97) EXPECT(r[i] == (float)n * (i % 13) + n * (n - 1) / 2);
  CHECK(ncclAllGather(d_in, d_out, N, ncclFloat, comm, st));
  CU(cudaMemcpyAsync(r.data(), d_out, N * 4 * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int k = 0; k < n; ++k)
    for (size_t i = 0; i < N; i += 1013) EXPECT(r[k * N + i] == (float)(i % 13) + k);
  CHECK(ncclBroadcast(d_in, d_out, N, ncclFloat, n - 1, comm, st));
  CU(cudaMemcpyAsync(r.data(), d_out, N * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (size_t i = 0; i < N; i += 1013) EXPECT(r[i] == (float)(i % 13) + (n - 1));
  const int nxt = (rank + 1) % n, prv = (rank + n - 1) % n;
  CHECK(ncclGroupStart());
  CHECK(ncclSend(d_in, N, ncclFloat, nxt, comm, st));
  CHECK(ncclRecv(d_out, N, ncclFloat, prv, comm, st));
  CHECK(ncclGroupEnd());
  CU(cudaMemcpyAsync(r.data(), d_out, N * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (size_t i = 0; i < N; i += 1013) EXPECT(r[i] == (float)(i % 13) + prv);
  CHECK(ncclCommDestroy(comm));
  return 0;
}

// fork the ranks FIRST and hand them the id over pipes: ncclGetUniqueId starts the rendezvous relay thread, and a child
// forked from a multi-threaded parent inherits whatever locks that thread held at the fork (allocator, logger)
static ncclUniqueId fork_ranks_then_make_id(int n, pid_t* pids, int (*body)(int, int, ncclUniqueId)) {
  std::vector<int> wr(n, -1);
  for (int r = 1; r < n; ++r) {
    int fds[2];
    if (pipe(fds) != 0) _exit(4);
    pids[r] = fork();
    if (pids[r] == 0) {
      for (int q = 1; q < r; ++q) close(wr[q]);
      close(fds[1]);
      ncclUniqueId cid;
      if (read(fds[0], &cid, sizeof(cid)) != (ssize_t)sizeof(cid)) _exit(3);
      close(fds[0]);
      _exit(body(r, n, cid));
    }
    close(fds[0]);
    wr[r] = fds[1];
  }
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) {
    fprintf(stderr, "ncclGetUniqueId failed\n");
    exit(1);
  }
  for (int r = 1; r < n; ++r) {
    if (write(wr[r], &id, sizeof(id)) != (ssize_t)sizeof(id)) exit(1);
    close(wr[r]);
  }
  return id;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2;
  setenv("UCCL_B200_LOCAL_SIZE", "1", 1);
  setenv("UCCL_B200_NET_BIND_IP", "127.0.0.1", 1);
  setenv("UCCL_B200_NCCL_HEAP_MB", "512", 0);
  setenv("UCCL_B200_NCCL_STAGE_MB", "16", 0);
  std::vector<pid_t> pids(n, 0);
  ncclUniqueId id = fork_ranks_then_make_id(n, pids.data(), run);  // also: fork before any CUDA call in this process
  bool ok = run(0, n, id) == 0;
  for (int r = 1; r < n; ++r) {
    int s = 0;
    waitpid(pids[r], &s, 0);
    ok = ok && WIFEXITED(s) && WEXITSTATUS(s) == 0;
  }
  printf(ok ? "nccl_multibox_gpu_test: OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}
