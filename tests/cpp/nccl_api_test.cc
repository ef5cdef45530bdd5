// NCCL-API plumbing test (BASELINE config #1): two processes, world_size = 2, host backend.
// Exercises ncclGetUniqueId / ncclCommInitRank (TCP rendezvous + shm heaps), AllReduce,
// AllGather, ReduceScatter, Broadcast, Reduce, CommCount/UserRank, CommSplit, error paths.
#include <nccl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(x)                                                                              \
  do {                                                                                        \
    ncclResult_t _r = (x);                                                                    \
    if (_r != ncclSuccess) {                                                                  \
      fprintf(stderr, "rank %d: %s failed: %s (%s)\n", g_rank, #x, ncclGetErrorString(_r), ncclGetLastError(nullptr)); \
      exit(2);                                                                                \
    }                                                                                         \
  } while (0)
#define EXPECT(c)                                                         \
  do {                                                                    \
    if (!(c)) {                                                           \
      fprintf(stderr, "rank %d: expectation failed: %s (line %d)\n", g_rank, #c, __LINE__); \
      exit(3);                                                            \
    }                                                                     \
  } while (0)

// NCCL 2.28 entry points the drop-in exports; the system header of this image is 2.27
extern "C" {
ncclResult_t ncclAlltoAll(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
ncclResult_t ncclGather(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t ncclScatter(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t ncclCommRevoke(ncclComm_t, int);
}

static int g_rank = -1;

static int run(int rank, int n, ncclUniqueId id) {
  g_rank = rank;
  ncclComm_t comm;
  CHECK(ncclCommInitRank(&comm, n, id, rank));
  int cnt = 0, ur = -1, ver = 0;
  CHECK(ncclCommCount(comm, &cnt));
  CHECK(ncclCommUserRank(comm, &ur));
  CHECK(ncclGetVersion(&ver));
  EXPECT(cnt == n && ur == rank && ver >= 20000);

  const size_t N = 100003;
  std::vector<float> x(N), y(N, 0.f);
  for (size_t i = 0; i < N; ++i) x[i] = (float)(i % 97) + rank;
  CHECK(ncclAllReduce(x.data(), y.data(), N, ncclFloat, ncclSum, comm, nullptr));
  for (size_t i = 0; i < N; ++i) EXPECT(y[i] == 2.f * (i % 97) + 1.f);
  // in place, avg
  CHECK(ncclAllReduce(x.data(), x.data(), N, ncclFloat, ncclAvg, comm, nullptr));
  for (size_t i = 0; i < N; ++i) EXPECT(std::fabs(x[i] - ((i % 97) + 0.5f)) < 1e-5f);

  std::vector<int> g(n * 10), mine(10, rank + 1);
  CHECK(ncclAllGather(mine.data(), g.data(), 10, ncclInt32, comm, nullptr));
  for (int r = 0; r < n; ++r)
    for (int i = 0; i < 10; ++i) EXPECT(g[r * 10 + i] == r + 1);

  std::vector<double> rs_in(n * 7), rs_out(7);
  for (size_t i = 0; i < rs_in.size(); ++i) rs_in[i] = (double)i * (rank + 1);
  CHECK(ncclReduceScatter(rs_in.data(), rs_out.data(), 7, ncclDouble, ncclSum, comm, nullptr));
  for (int i = 0; i < 7; ++i) EXPECT(rs_out[i] == (double)(rank * 7 + i) * 3.0);

  std::vector<long long> b(33, rank == 1 ? 42 : -1);
  CHECK(ncclBroadcast(b.data(), b.data(), 33, ncclInt64, 1, comm, nullptr));
  for (auto v : b) EXPECT(v == 42);

  std::vector<float> red(5, (float)(rank + 2)), red_out(5, 0.f);
  CHECK(ncclReduce(red.data(), red_out.data(), 5, ncclFloat, ncclProd, 0, comm, nullptr));
  if (rank == 0)
    for (auto v : red_out) EXPECT(v == 6.f);

  // user-defined PreMulSum: 0.25 * (sum over ranks), fused into the reduction
  {
    float scalar = 0.25f;
    ncclRedOp_t premul;
    CHECK(ncclRedOpCreatePreMulSum(&premul, &scalar, ncclFloat, ncclScalarHostImmediate, comm));
    std::vector<float> pin(9, (float)(rank + 1)), pout(9, 0.f);
    CHECK(ncclAllReduce(pin.data(), pout.data(), 9, ncclFloat, premul, comm, nullptr));
    for (auto v : pout) EXPECT(v == 0.25f * 3.f);
    CHECK(ncclRedOpDestroy(premul, comm));
    EXPECT(ncclAllReduce(pin.data(), pout.data(), 9, ncclFloat, premul, comm, nullptr) == ncclInvalidArgument);
  }

  // grouped send/recv (ring step) and a lone send / recv pair
  {
    const int next = (rank + 1) % n, prev = (rank + n - 1) % n;
    std::vector<int> sbuf(5000, 1000 + rank), rbuf(5000, -1);
    CHECK(ncclGroupStart());
    CHECK(ncclSend(sbuf.data(), sbuf.size(), ncclInt32, next, comm, nullptr));
    CHECK(ncclRecv(rbuf.data(), rbuf.size(), ncclInt32, prev, comm, nullptr));
    CHECK(ncclGroupEnd());
    for (auto v : rbuf) EXPECT(v == 1000 + prev);
    float token = (float)rank, got = -1.f;
    if (rank == 0) {
      CHECK(ncclSend(&token, 1, ncclFloat, 1, comm, nullptr));
    } else if (rank == 1) {
      CHECK(ncclRecv(&got, 1, ncclFloat, 0, comm, nullptr));
      EXPECT(got == 0.f);
    }
  }

  // NCCL 2.28 additions: ncclAlltoAll (spelling), ncclGather / ncclScatter (rooted), ncclCommRevoke
  {
    std::vector<int> ain(n * 7), aout(n * 7, -1);
    for (int j = 0; j < n; ++j)
      for (int k = 0; k < 7; ++k) ain[j * 7 + k] = rank * 100 + j;
    CHECK(ncclAlltoAll(ain.data(), aout.data(), 7, ncclInt32, comm, nullptr));
    for (int i = 0; i < n; ++i) EXPECT(aout[i * 7] == i * 100 + rank && aout[i * 7 + 6] == i * 100 + rank);
    std::vector<double> gin(11, 10.0 + rank), gout(n * 11, -1.0);
    CHECK(ncclGather(gin.data(), gout.data(), 11, ncclFloat64, 1, comm, nullptr));
    if (rank == 1)
      for (int i = 0; i < n; ++i) EXPECT(gout[i * 11] == 10.0 + i && gout[i * 11 + 10] == 10.0 + i);
    std::vector<short> sin(n * 13), sout(13, -1);
    for (int j = 0; j < n; ++j)
      for (int k = 0; k < 13; ++k) sin[j * 13 + k] = (short)(500 + j);
    CHECK(ncclScatter(sin.data(), sout.data(), 13, ncclHalf, 0, comm, nullptr));  // 2-byte elements, moved as bytes
    for (auto v : sout) EXPECT(v == (short)(500 + rank));
    // in place on the root: recvbuff == sendbuff + root * count
    CHECK(ncclScatter(sin.data(), rank == 0 ? sin.data() : sout.data(), 13, ncclHalf, 0, comm, nullptr));
    EXPECT(ncclGather(gin.data(), gout.data(), 11, ncclFloat64, 9, comm, nullptr) == ncclInvalidArgument);
    EXPECT(ncclCommRevoke(comm, 1) == ncclInvalidArgument);
    CHECK(ncclCommRevoke(comm, 0));
  }

  // error paths
  EXPECT(ncclAllReduce(x.data(), y.data(), 4, (ncclDataType_t)99, ncclSum, comm, nullptr) == ncclInvalidArgument);
  EXPECT(ncclBroadcast(b.data(), b.data(), 1, ncclInt64, 7, comm, nullptr) == ncclInvalidArgument);

  // split into singleton communicators
  ncclComm_t sub;
  CHECK(ncclCommSplit(comm, rank, 0, &sub, nullptr));
  int sc = 0;
  CHECK(ncclCommCount(sub, &sc));
  EXPECT(sc == 1);
  float one = 3.f, out = 0.f;
  CHECK(ncclAllReduce(&one, &out, 1, ncclFloat, ncclSum, sub, nullptr));
  EXPECT(out == 3.f);
  CHECK(ncclCommDestroy(sub));
  CHECK(ncclCommDestroy(comm));
  return 0;
}

int main() {
  setenv("UCCL_B200_HOST_FAKE", "1", 1);
  setenv("UCCL_B200_TIMEOUT_MS", "30000", 0);
  // fork FIRST, then create the id in the parent and hand it over a pipe: ncclGetUniqueId starts the rendezvous relay
  // thread, and a child forked from a multi-threaded parent inherits whatever locks that thread held at the moment
  // of the fork (allocator, logger) -- under ASan the child's own relay then hung about one run in ten.
  int fds[2];
  if (pipe(fds) != 0) return 1;
  pid_t pid = fork();
  if (pid == 0) {
    close(fds[1]);
    ncclUniqueId cid;
    if (read(fds[0], &cid, sizeof(cid)) != (ssize_t)sizeof(cid)) _exit(3);
    close(fds[0]);
    _exit(run(1, 2, cid));
  }
  close(fds[0]);
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess || write(fds[1], &id, sizeof(id)) != (ssize_t)sizeof(id)) {
    fprintf(stderr, "ncclGetUniqueId failed\n");
    return 1;
  }
  close(fds[1]);
  int rc = run(0, 2, id);
  int st = 0;
  waitpid(pid, &st, 0);
  if (rc != 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) {
    fprintf(stderr, "FAILED (rank0 rc=%d, rank1 status=%d)\n", rc, st);
    return 1;
  }
  printf("nccl_api_test: OK\n");
  return 0;
}
