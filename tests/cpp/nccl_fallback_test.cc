// The drop-in's optional forwarding to a dlopen'd libnccl (csrc/coll/nccl_fallback.cc), against tests/cpp/fake_nccl.cc:
// two host-mode ranks; broadcast >= 1 KiB and all send/recv are configured to go to the "real" library, everything else
// must stay native.  argv[1] = path of the fake library.
#include <dlfcn.h>
#include <nccl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static int g_fail = 0;
#define EXPECT(c)                                             \
  do {                                                        \
    if (!(c)) {                                               \
      fprintf(stderr, "FAILED %s @%d\n", #c, __LINE__);       \
      ++g_fail;                                               \
    }                                                         \
  } while (0)
#define CHECK(x) EXPECT((x) == ncclSuccess)

static int run(int rank, int n, ncclUniqueId id, const char* lib) {
  ncclComm_t comm = nullptr;
  CHECK(ncclCommInitRank(&comm, n, id, rank));
  if (!comm) return 1;
  // native: all-reduce is not in the forward list
  std::vector<float> x(1000, (float)(rank + 1)), y(1000, 0.f);
  CHECK(ncclAllReduce(x.data(), y.data(), x.size(), ncclFloat, ncclSum, comm, nullptr));
  EXPECT(y[0] == 3.f && y[999] == 3.f);
  // native: a broadcast below the size threshold
  std::vector<unsigned char> s(256, (unsigned char)(10 + rank)), r(256, 0);
  CHECK(ncclBroadcast(s.data(), r.data(), s.size(), ncclChar, 1, comm, nullptr));
  EXPECT(r[0] == 11 && r[255] == 11);
  // forwarded: a broadcast at / above the threshold carries the fake library's marker
  std::vector<unsigned char> bs(4096, 1), br(4096, 0);
  CHECK(ncclBroadcast(bs.data(), br.data(), bs.size(), ncclChar, 1, comm, nullptr));
  EXPECT(br[0] == 0xB1 && br[4095] == 0xB1);
  // forwarded: grouped send/recv ring (send/recv ignore the size threshold)
  std::vector<unsigned char> ps(64, 7), pr(64, 0);
  CHECK(ncclGroupStart());
  CHECK(ncclSend(ps.data(), ps.size(), ncclChar, (rank + 1) % n, comm, nullptr));
  CHECK(ncclRecv(pr.data(), pr.size(), ncclChar, (rank + n - 1) % n, comm, nullptr));
  CHECK(ncclGroupEnd());
  EXPECT(pr[0] == 0x50 + (rank + n - 1) % n);
  // a user-created PreMulSum operator only exists in the drop-in: it stays native even for a listed operation
  int counts[8] = {0};
  void* h = dlopen(lib, RTLD_NOW | RTLD_NOLOAD);
  EXPECT(h != nullptr);
  if (h) {
    auto fn = (void (*)(int*))dlsym(h, "fake_nccl_counts");
    EXPECT(fn != nullptr);
    if (fn) fn(counts);
  }
  EXPECT(counts[0] == 0 && counts[2] == 1 && counts[5] == 1 && counts[6] == 1 && counts[7] == 1);
  CHECK(ncclCommDestroy(comm));
  return g_fail ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  setenv("UCCL_B200_HOST_FAKE", "1", 1);
  setenv("UCCL_B200_TIMEOUT_MS", "30000", 0);
  setenv("UCCL_B200_NCCL_FALLBACK_LIB", argv[1], 1);
  setenv("UCCL_B200_NCCL_FALLBACK_OPS", "broadcast, send_recv", 1);
  setenv("UCCL_B200_NCCL_FALLBACK_MIN_BYTES", "1024", 1);
  int fds[2];
  if (pipe(fds) != 0) return 1;
  pid_t pid = fork();  // before the id exists (no threads yet)
  if (pid == 0) {
    close(fds[1]);
    ncclUniqueId cid;
    if (read(fds[0], &cid, sizeof(cid)) != (ssize_t)sizeof(cid)) _exit(3);
    _exit(run(1, 2, cid, argv[1]));
  }
  close(fds[0]);
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess || write(fds[1], &id, sizeof(id)) != (ssize_t)sizeof(id)) return 1;
  int rc = run(0, 2, id, argv[1]);
  int st = 0;
  waitpid(pid, &st, 0);
  // a misspelt operation is a configuration error, not a silent no-op
  pid_t bad = fork();
  if (bad == 0) {
    setenv("UCCL_B200_NCCL_FALLBACK_OPS", "brodcast", 1);
    ncclUniqueId id2;
    ncclComm_t c2 = nullptr;
    if (ncclGetUniqueId(&id2) != ncclSuccess) _exit(5);
    _exit(ncclCommInitRank(&c2, 1, id2, 0) == ncclSuccess ? 1 : 0);
  }
  int st2 = 0;
  waitpid(bad, &st2, 0);
  if (rc != 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0 || !WIFEXITED(st2) || WEXITSTATUS(st2) != 0) {
    fprintf(stderr, "FAILED (rank0 rc=%d, rank1 status=%d, bad-config status=%d)\n", rc, st, st2);
    return 1;
  }
  printf("nccl_fallback_test: OK\n");
  return 0;
}
