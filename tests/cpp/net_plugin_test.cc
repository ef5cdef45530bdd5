// Drives the NCCL net plugin (ncclNet v8 vtable) the way NCCL's proxy does: dlopen, init, listen on one
// "rank", non-blocking connect/accept loops, isend/irecv/test of several sizes, close.
//   net_plugin_test <path to libnccl-net-uccl_b200.so>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

extern "C" {
typedef int ncclResult_t;
typedef void (*logger_t)(int level, unsigned long flags, const char* file, int line, const char* fmt, ...);
typedef struct {
  char* name;
  char* pciPath;
  uint64_t guid;
  int ptrSupport, regIsGlobal, speed, port;
  float latency;
  int maxComms, maxRecvs, netDeviceType, netDeviceVersion;
} props_t;
typedef struct {
  const char* name;
  ncclResult_t (*init)(logger_t);
  ncclResult_t (*devices)(int*);
  ncclResult_t (*getProperties)(int, props_t*);
  ncclResult_t (*listen)(int, void*, void**);
  ncclResult_t (*connect)(int, void*, void**, void**);
  ncclResult_t (*accept)(void*, void**, void**);
  ncclResult_t (*regMr)(void*, void*, size_t, int, void**);
  ncclResult_t (*regMrDmaBuf)(void*, void*, size_t, int, uint64_t, int, void**);
  ncclResult_t (*deregMr)(void*, void*);
  ncclResult_t (*isend)(void*, void*, int, int, void*, void**);
  ncclResult_t (*irecv)(void*, int, void**, int*, int*, void**, void**);
  ncclResult_t (*iflush)(void*, int, void**, int*, void**, void**);
  ncclResult_t (*test)(void*, int*, int*);
  ncclResult_t (*closeSend)(void*);
  ncclResult_t (*closeRecv)(void*);
  ncclResult_t (*closeListen)(void*);
  ncclResult_t (*getDeviceMr)(void*, void*, void**);
  ncclResult_t (*irecvConsumed)(void*, int, void*);
} net_t;
}

#define REQUIRE(c)                                      \
  do {                                                  \
    if (!(c)) {                                         \
      fprintf(stderr, "FAILED %s @%d\n", #c, __LINE__); \
      exit(1);                                          \
    }                                                   \
  } while (0)

static void logger(int level, unsigned long, const char*, int, const char* fmt, ...) {
  if (level > 3) return;
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}

int main(int argc, char** argv) {
  REQUIRE(argc > 1);
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) fprintf(stderr, "%s\n", dlerror());
  REQUIRE(lib);
  net_t* net = (net_t*)dlsym(lib, "ncclNetPlugin_v8");
  REQUIRE(net && !strcmp(net->name, "uccl_b200"));
  REQUIRE(net->init(logger) == 0);
  int ndev = 0;
  REQUIRE(net->devices(&ndev) == 0 && ndev >= 1);
  props_t pr;
  REQUIRE(net->getProperties(0, &pr) == 0);
  REQUIRE(pr.name && pr.ptrSupport == 1 && pr.maxRecvs == 1 && pr.speed > 0);
  printf("dev0 %s speed %d Mb/s\n", pr.name, pr.speed);

  char handle[128];
  void* lcomm = nullptr;
  REQUIRE(net->listen(0, handle, &lcomm) == 0 && lcomm);
  void *scomm = nullptr, *rcomm = nullptr;
  char hcopy[128];
  memcpy(hcopy, handle, sizeof(handle));  // the handle travels to the peer through NCCL's bootstrap
  for (int i = 0; i < 200000 && (!scomm || !rcomm); ++i) {
    if (!scomm) REQUIRE(net->connect(0, hcopy, &scomm, nullptr) == 0);
    if (!rcomm) REQUIRE(net->accept(lcomm, &rcomm, nullptr) == 0);
    if (!scomm || !rcomm) std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
  REQUIRE(scomm && rcomm);
  void *mh_s = nullptr, *mh_r = nullptr;
  const int sizes[] = {0, 8, 4096, 65536, 1 << 20, (8 << 20) + 13};
  std::vector<char> sbuf(9 << 20), rbuf(9 << 20);
  for (size_t i = 0; i < sbuf.size(); ++i) sbuf[i] = (char)(i * 131 + 7);
  REQUIRE(net->regMr(scomm, sbuf.data(), sbuf.size(), 1, &mh_s) == 0);
  REQUIRE(net->regMr(rcomm, rbuf.data(), rbuf.size(), 1, &mh_r) == 0);
  REQUIRE(net->regMr(rcomm, rbuf.data(), rbuf.size(), 2 /*CUDA*/, &mh_r) != 0);  // host pointers only
  REQUIRE(net->regMr(rcomm, rbuf.data(), rbuf.size(), 1, &mh_r) == 0);
  for (int sz : sizes) {
    memset(rbuf.data(), 0, (size_t)sz + 1);
    // NCCL posts receives of the maximum size and learns the real size from test()
    void *rreq = nullptr, *sreq = nullptr;
    void* rptr = rbuf.data();
    int rsz = (int)rbuf.size(), tag = 0;
    REQUIRE(net->irecv(rcomm, 1, &rptr, &rsz, &tag, &mh_r, &rreq) == 0 && rreq);
    REQUIRE(net->isend(scomm, sbuf.data(), sz, 0, mh_s, &sreq) == 0 && sreq);
    int sd = 0, rd = 0, got = -1, sgot = -1;
    for (long spin = 0; spin < 400000000L && !(sd && rd); ++spin) {
      if (!sd) REQUIRE(net->test(sreq, &sd, &sgot) == 0);
      if (!rd) REQUIRE(net->test(rreq, &rd, &got) == 0);
    }
    REQUIRE(sd && rd);
    REQUIRE(got == sz && sgot == sz);
    REQUIRE(memcmp(rbuf.data(), sbuf.data(), (size_t)sz) == 0);
    void* freq = (void*)1;
    REQUIRE(net->iflush(rcomm, 1, &rptr, &rsz, &mh_r, &freq) == 0 && freq == nullptr);
  }
  // a pipeline of 8 outstanding requests (NCCL keeps up to NCCL_NET_MAX_REQUESTS in flight)
  {
    void *rq[8], *sq[8];
    const int sz = 300000;
    for (int i = 0; i < 8; ++i) {
      void* rptr = rbuf.data() + (size_t)i * sz;
      int rsz = sz, tag = 0;
      REQUIRE(net->irecv(rcomm, 1, &rptr, &rsz, &tag, &mh_r, &rq[i]) == 0);
    }
    for (int i = 0; i < 8; ++i) REQUIRE(net->isend(scomm, sbuf.data() + (size_t)i * sz, sz, 0, mh_s, &sq[i]) == 0);
    for (int i = 0; i < 8; ++i) {
      int d = 0, n = 0;
      while (!d) REQUIRE(net->test(rq[i], &d, &n) == 0);
      REQUIRE(n == sz);
      d = 0;
      while (!d) REQUIRE(net->test(sq[i], &d, &n) == 0);
    }
    REQUIRE(memcmp(rbuf.data(), sbuf.data(), (size_t)8 * sz) == 0);
  }
  REQUIRE(net->deregMr(scomm, mh_s) == 0 && net->deregMr(rcomm, mh_r) == 0);
  REQUIRE(net->closeSend(scomm) == 0 && net->closeRecv(rcomm) == 0 && net->closeListen(lcomm) == 0);
  printf("PASS\n");
  return 0;
}
