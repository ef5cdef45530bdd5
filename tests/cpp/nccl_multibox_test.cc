// NCCL API over a group that spans boxes: 4 processes = 2 "boxes" x 2 ranks (UCCL_B200_LOCAL_SIZE=2) on the
// host backend.  Inside a box the native communicator (shared-memory heap), between boxes the datagram
// rails; every collective is checked against its closed form.
#include <nccl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(x)                                                                                                       \
  do {                                                                                                                 \
    ncclResult_t _r = (x);                                                                                             \
    if (_r != ncclSuccess) {                                                                                           \
      fprintf(stderr, "rank %d: %s failed: %s (%s)\n", g_rank, #x, ncclGetErrorString(_r), ncclGetLastError(nullptr)); \
      exit(2);                                                                                                         \
    }                                                                                                                  \
  } while (0)
#define EXPECT(c)                                                                           \
  do {                                                                                      \
    if (!(c)) {                                                                             \
      fprintf(stderr, "rank %d: expectation failed: %s (line %d)\n", g_rank, #c, __LINE__); \
      exit(3);                                                                              \
    }                                                                                       \
  } while (0)

static int g_rank = -1;
// exported by the drop-in (and by newer NCCL), not declared in nccl.h 2.27
extern "C" ncclResult_t ncclAllToAll(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
extern "C" ncclResult_t ncclAllToAllv(const void*, const size_t*, const size_t*, void*, const size_t*, const size_t*,
                                      ncclDataType_t, ncclComm_t, cudaStream_t);

static int run(int rank, int n, ncclUniqueId id) {
  g_rank = rank;
  ncclComm_t comm;
  CHECK(ncclCommInitRank(&comm, n, id, rank));
  int cnt = 0, ur = -1;
  CHECK(ncclCommCount(comm, &cnt));
  CHECK(ncclCommUserRank(comm, &ur));
  EXPECT(cnt == n && ur == rank);
  const float tri = (float)(n * (n - 1) / 2);

  // all-reduce: a count that is not a multiple of the box size, in and out of place, sum / avg / max
  const size_t N = 200003;
  std::vector<float> x(N), y(N, 0.f);
  for (size_t i = 0; i < N; ++i) x[i] = (float)(i % 89) + rank;
  CHECK(ncclAllReduce(x.data(), y.data(), N, ncclFloat, ncclSum, comm, nullptr));
  for (size_t i = 0; i < N; ++i) EXPECT(y[i] == (float)n * (i % 89) + tri);
  CHECK(ncclAllReduce(x.data(), x.data(), N, ncclFloat, ncclAvg, comm, nullptr));
  for (size_t i = 0; i < N; ++i) EXPECT(std::fabs(x[i] - ((i % 89) + tri / n)) < 1e-4f);
  std::vector<int> mi(777, rank * 10), mo(777, -1);
  CHECK(ncclAllReduce(mi.data(), mo.data(), 777, ncclInt32, ncclMax, comm, nullptr));
  for (auto v : mo) EXPECT(v == (n - 1) * 10);
  {  // bf16: 1, 2, 3, 4 -> 10 (0x4120) and max -> 4 (0x4080); exercises the vectorised rail reduction
    const uint16_t enc[4] = {0x3F80, 0x4000, 0x4040, 0x4080};
    std::vector<uint16_t> bi(4099, enc[rank]), bo(4099, 0);
    CHECK(ncclAllReduce(bi.data(), bo.data(), bi.size(), ncclBfloat16, ncclSum, comm, nullptr));
    for (auto v : bo) EXPECT(v == 0x4120);
    CHECK(ncclAllReduce(bi.data(), bo.data(), bi.size(), ncclBfloat16, ncclMax, comm, nullptr));
    for (auto v : bo) EXPECT(v == 0x4080);
  }
  {
    float scalar = 0.5f;
    ncclRedOp_t premul;
    CHECK(ncclRedOpCreatePreMulSum(&premul, &scalar, ncclFloat, ncclScalarHostImmediate, comm));
    std::vector<float> pin(11, (float)(rank + 1)), pout(11, 0.f);
    CHECK(ncclAllReduce(pin.data(), pout.data(), 11, ncclFloat, premul, comm, nullptr));
    for (auto v : pout) EXPECT(v == 0.5f * (tri + n));
    CHECK(ncclRedOpDestroy(premul, comm));
  }

  std::vector<int> g(n * 1000), mine(1000, rank + 1);
  CHECK(ncclAllGather(mine.data(), g.data(), 1000, ncclInt32, comm, nullptr));
  for (int r = 0; r < n; ++r)
    for (int i = 0; i < 1000; ++i) EXPECT(g[r * 1000 + i] == r + 1);

  std::vector<double> rs_in(n * 513), rs_out(513);
  for (size_t i = 0; i < rs_in.size(); ++i) rs_in[i] = (double)i * (rank + 1);
  CHECK(ncclReduceScatter(rs_in.data(), rs_out.data(), 513, ncclDouble, ncclSum, comm, nullptr));
  for (int i = 0; i < 513; ++i) EXPECT(rs_out[i] == (double)(rank * 513 + i) * (tri + n));

  for (int root : {3, 0, 1}) {
    std::vector<long long> b(100001, rank == root ? 4242 + root : -1);
    CHECK(ncclBroadcast(b.data(), b.data(), b.size(), ncclInt64, root, comm, nullptr));
    for (auto v : b) EXPECT(v == 4242 + root);
  }

  std::vector<float> red(50, (float)(rank + 1)), red_out(50, 0.f);
  CHECK(ncclReduce(red.data(), red_out.data(), 50, ncclFloat, ncclProd, 2, comm, nullptr));
  if (rank == 2)
    for (auto v : red_out) EXPECT(v == 24.f);

  std::vector<int> a_in(n * 300), a_out(n * 300, -1);
  for (int d = 0; d < n; ++d)
    for (int i = 0; i < 300; ++i) a_in[d * 300 + i] = 1000 * rank + 10 * d + (i % 7);
  CHECK(ncclAllToAll(a_in.data(), a_out.data(), 300, ncclInt32, comm, nullptr));
  for (int s = 0; s < n; ++s)
    for (int i = 0; i < 300; ++i) EXPECT(a_out[s * 300 + i] == 1000 * s + 10 * rank + (i % 7));

  {  // variable splits: rank s sends (s + 2 d) % 5 + 1 values to rank d
    std::vector<size_t> sc(n), sd(n), rc(n), rd(n);
    size_t st = 0, rt = 0;
    for (int d = 0; d < n; ++d) {
      sc[d] = (size_t)((rank + 2 * d) % 5 + 1) * 100, sd[d] = st, st += sc[d];
      rc[d] = (size_t)((d + 2 * rank) % 5 + 1) * 100, rd[d] = rt, rt += rc[d];
    }
    std::vector<float> vin(st), vout(rt, -1.f);
    for (int d = 0; d < n; ++d)
      for (size_t i = 0; i < sc[d]; ++i) vin[sd[d] + i] = (float)(100 * rank + d);
    CHECK(ncclAllToAllv(vin.data(), sc.data(), sd.data(), vout.data(), rc.data(), rd.data(), ncclFloat, comm, nullptr));
    for (int s = 0; s < n; ++s)
      for (size_t i = 0; i < rc[s]; ++i) EXPECT(vout[rd[s] + i] == (float)(100 * s + rank));
  }

  // grouped send/recv: the box-mate (native kernel path) and the rail-mate (datagram path) in one group
  {
    const int mate = rank ^ 1, rail = (rank + 2) % n;
    std::vector<int> s1(70000, 100 + rank), r1(70000, -1), s2(70000, 200 + rank), r2(70000, -1);
    CHECK(ncclGroupStart());
    CHECK(ncclSend(s1.data(), s1.size(), ncclInt32, mate, comm, nullptr));
    CHECK(ncclRecv(r1.data(), r1.size(), ncclInt32, mate, comm, nullptr));
    CHECK(ncclSend(s2.data(), s2.size(), ncclInt32, rail, comm, nullptr));
    CHECK(ncclRecv(r2.data(), r2.size(), ncclInt32, rail, comm, nullptr));
    CHECK(ncclGroupEnd());
    for (auto v : r1) EXPECT(v == 100 + mate);
    for (auto v : r2) EXPECT(v == 200 + rail);
    // a peer on another rail of another box is refused with a clear error, not a hang
    const int diag = (rank + 2) % n ^ 1;
    EXPECT(ncclSend(s1.data(), 1, ncclInt32, diag, comm, nullptr) != ncclSuccess);
    EXPECT(strstr(ncclGetLastError(comm), "not routed") != nullptr);
  }
  {  // split into rails (one member per box -> still spans boxes) and into boxes (plain NVLink communicators)
    ncclComm_t rail_comm = nullptr, box_comm = nullptr;
    CHECK(ncclCommSplit(comm, rank % 2, rank, &rail_comm, nullptr));
    CHECK(ncclCommSplit(comm, rank / 2, rank, &box_comm, nullptr));
    int c1 = 0, c2 = 0, r1 = -1, r2 = -1;
    CHECK(ncclCommCount(rail_comm, &c1));
    CHECK(ncclCommCount(box_comm, &c2));
    CHECK(ncclCommUserRank(rail_comm, &r1));
    CHECK(ncclCommUserRank(box_comm, &r2));
    EXPECT(c1 == 2 && c2 == 2 && r1 == rank / 2 && r2 == rank % 2);
    std::vector<float> v(3000, (float)rank), o(3000, -1.f);
    CHECK(ncclAllReduce(v.data(), o.data(), v.size(), ncclFloat, ncclSum, rail_comm, nullptr));
    for (auto e : o) EXPECT(e == (float)(rank % 2) * 2 + 2);          // ranks {l, l+2}
    CHECK(ncclAllReduce(v.data(), o.data(), v.size(), ncclFloat, ncclSum, box_comm, nullptr));
    for (auto e : o) EXPECT(e == (float)(rank / 2) * 4 + 1);          // ranks {2k, 2k+1}
    CHECK(ncclCommDestroy(rail_comm));
    CHECK(ncclCommDestroy(box_comm));
    // members per box must be equal: {0,1,2} vs {3} is refused
    ncclComm_t bad = nullptr;
    EXPECT(ncclCommSplit(comm, rank == 3 ? 1 : 0, rank, &bad, nullptr) == (rank == 3 ? ncclSuccess : ncclInvalidUsage));
    if (bad) CHECK(ncclCommDestroy(bad));
  }
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

int main() {
  setenv("UCCL_B200_HOST_FAKE", "1", 1);
  setenv("UCCL_B200_LOCAL_SIZE", "2", 1);
  setenv("UCCL_B200_NET_BIND_IP", "127.0.0.1", 1);
  setenv("UCCL_B200_NET_PATHS", "2", 0);
  setenv("UCCL_B200_MN_PIPELINE_BYTES", "65536", 0);  // the 200003-float all-reduce runs as a 7-block pipeline
  setenv("UCCL_B200_TIMEOUT_MS", "30000", 0);
  const int n = 4;
  pid_t pids[4] = {0};
  ncclUniqueId id = fork_ranks_then_make_id(n, pids, run);
  int rc = run(0, n, id);
  bool ok = rc == 0;
  for (int r = 1; r < n; ++r) {
    int st = 0;
    waitpid(pids[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) {
      fprintf(stderr, "rank %d exited with status %d\n", r, st);
      ok = false;
    }
  }
  if (!ok) {
    fprintf(stderr, "FAILED\n");
    return 1;
  }
  printf("nccl_multibox_test: OK\n");
  return 0;
}
