// uccl_engine_* C API (NIXL-plugin surface) in host mode: two engines in one process,
// connect/accept, registration, one-sided write/read against a prepared descriptor, vector write,
// two-sided send/recv, transfer status polling, notifications.
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "uccl_engine.h"

static int g_fail = 0;
#define EXPECT(c)                                              \
  do {                                                         \
    if (!(c)) {                                                \
      fprintf(stderr, "FAILED %s @%d\n", #c, __LINE__);        \
      ++g_fail;                                                \
    }                                                          \
  } while (0)

static bool wait_done(uccl_conn_t* c, uint64_t tid) {
  for (int i = 0; i < 200000; ++i) {
    if (uccl_engine_xfer_status(c, tid)) return true;
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
  return false;
}

int main() {
  uccl_engine_t* a = uccl_engine_create_on(-1, 2);
  uccl_engine_t* b = uccl_engine_create_on(-1, 2);
  EXPECT(a && b);
  char* md = nullptr;
  EXPECT(uccl_engine_get_metadata(b, &md) == 0 && md != nullptr);
  // metadata is "ip:port?gpu" style text: parse ip and port
  std::string m(md);
  free(md);
  const size_t colon = m.find(':');
  EXPECT(colon != std::string::npos);
  const std::string ip = m.substr(0, colon);
  const int port = atoi(m.c_str() + colon + 1);
  EXPECT(port > 0);
  uccl_conn_t* ca = uccl_engine_connect(a, ip.c_str(), -1, port);
  EXPECT(ca != nullptr);
  char ipbuf[64];
  int rgpu = 0;
  uccl_conn_t* cb = uccl_engine_accept(b, ipbuf, sizeof(ipbuf), &rgpu);
  EXPECT(cb != nullptr && rgpu == -1);
  EXPECT(uccl_engine_conn_is_local(ca));

  std::vector<unsigned char> src(100000), dst(100000, 0), back(100000, 0);
  for (size_t i = 0; i < src.size(); ++i) src[i] = (unsigned char)(i * 7);
  uccl_mr_t mra = 0, mrb = 0;
  EXPECT(uccl_engine_reg(a, (uintptr_t)src.data(), src.size(), mra) == 0);
  EXPECT(uccl_engine_reg(b, (uintptr_t)dst.data(), dst.size(), mrb) == 0);
  char fifo[UCCL_ENGINE_DESC_BYTES];
  EXPECT(uccl_engine_prepare_fifo(b, mrb, dst.data(), dst.size(), fifo) == 0);
  uint64_t tid = 0;
  EXPECT(uccl_engine_write(ca, mra, src.data(), src.size(), fifo, &tid) == 0 && wait_done(ca, tid));
  EXPECT(memcmp(src.data(), dst.data(), src.size()) == 0);
  EXPECT(uccl_engine_read(ca, mra, back.data(), back.size(), fifo, &tid) == 0 && wait_done(ca, tid));
  EXPECT(memcmp(back.data(), dst.data(), back.size()) == 0);
  // a sub-window: update_fifo narrows the descriptor to [addr, addr + size)
  memset(dst.data(), 0, dst.size());
  char sub[UCCL_ENGINE_DESC_BYTES];
  memcpy(sub, fifo, sizeof(sub));
  EXPECT(uccl_engine_update_fifo(sub, (uint64_t)(uintptr_t)(dst.data() + 1000), 500) == 0);
  EXPECT(uccl_engine_write(ca, mra, src.data(), 500, sub, &tid) == 0 && wait_done(ca, tid));
  EXPECT(dst[999] == 0 && dst[1000] == src[0] && dst[1499] == src[499] && dst[1500] == 0);
  // vector write of two blocks
  std::vector<unsigned char> d2(64, 0), d3(32, 0);
  uccl_mr_t m2 = 0, m3 = 0;
  uccl_engine_reg(b, (uintptr_t)d2.data(), d2.size(), m2);
  uccl_engine_reg(b, (uintptr_t)d3.data(), d3.size(), m3);
  char f2[UCCL_ENGINE_DESC_BYTES], f3[UCCL_ENGINE_DESC_BYTES];
  uccl_engine_prepare_fifo(b, m2, d2.data(), d2.size(), f2);
  uccl_engine_prepare_fifo(b, m3, d3.data(), d3.size(), f3);
  EXPECT(uccl_engine_write_vector(ca, {mra, mra}, {src.data(), src.data() + 64}, {64, 32},
                                  {std::string(f2, sizeof(f2)), std::string(f3, sizeof(f3))}, 2, &tid) == 0 &&
         wait_done(ca, tid));
  EXPECT(memcmp(d2.data(), src.data(), 64) == 0 && memcmp(d3.data(), src.data() + 64, 32) == 0);
  // the reference header's spellings (FifoItem by value / by reference, FIFO_SIZE buffers, string GPU identity)
  {
    char fbuf[FIFO_SIZE];
    EXPECT(uccl_engine_prepare_fifo(b, mrb, dst.data(), dst.size(), fbuf) == 0);
    FifoItem item;
    deserialize_fifo_item(fbuf, &item);
    memset(dst.data(), 0, dst.size());
    EXPECT(uccl_engine_update_fifo(item, (uint64_t)(uintptr_t)(dst.data() + 64), 128) == 0);
    EXPECT(uccl_engine_write(ca, mra, src.data(), 128, item, &tid) == 0 && wait_done(ca, tid));
    EXPECT(dst[63] == 0 && dst[64] == src[0] && dst[191] == src[127] && dst[192] == 0);
    std::vector<unsigned char> rb(128, 0);
    EXPECT(uccl_engine_read(ca, mra, rb.data(), 128, item, &tid) == 0 && wait_done(ca, tid));
    EXPECT(memcmp(rb.data(), src.data(), 128) == 0);
    FifoItem i2, i3;
    deserialize_fifo_item(f2, &i2);
    deserialize_fifo_item(f3, &i3);
    memset(d2.data(), 0, d2.size());
    memset(d3.data(), 0, d3.size());
    EXPECT(uccl_engine_write_vector(ca, {mra, mra}, {src.data() + 1, src.data() + 100}, {64, 32}, std::vector<FifoItem>{i2, i3}, 2,
                                    &tid) == 0 && wait_done(ca, tid));
    EXPECT(memcmp(d2.data(), src.data() + 1, 64) == 0 && memcmp(d3.data(), src.data() + 100, 32) == 0);
    char sbuf[FIFO_SIZE];
    serialize_fifo_item(i2, sbuf);
    EXPECT(memcmp(sbuf, f2, FIFO_SIZE) == 0);
    uccl_conn_t* c2 = uccl_engine_connect(a, ip.c_str(), "0", port);  // GPU identity as a string
    EXPECT(c2 != nullptr);
    char ipbuf2[64];
    int g2 = 0;
    uccl_conn_t* cb2 = uccl_engine_accept(b, ipbuf2, sizeof(ipbuf2), &g2);
    EXPECT(cb2 != nullptr);
    uccl_engine_conn_destroy(c2);
    uccl_engine_conn_destroy(cb2);
  }
  // two-sided
  std::vector<unsigned char> r(4096, 0);
  std::thread rx([&] { EXPECT(uccl_engine_recv(cb, mrb, r.data(), r.size()) == 0); });
  EXPECT(uccl_engine_send(ca, mra, src.data(), r.size(), &tid) == 0 && wait_done(ca, tid));
  rx.join();
  EXPECT(memcmp(r.data(), src.data(), r.size()) == 0);
  // notifications
  notify_msg_t n;
  memset(&n, 0, sizeof(n));
  snprintf(n.name, sizeof(n.name), "agentA");
  snprintf(n.msg, sizeof(n.msg), "xfer-complete:7");
  EXPECT(uccl_engine_send_notif(ca, &n) == 0);
  bool got = false;
  for (int i = 0; i < 2000 && !got; ++i) {
    for (auto& x : uccl_engine_get_notifs())
      if (strstr(x.msg, "xfer-complete:7")) got = true;
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  EXPECT(got);
  uccl_engine_mr_destroy(a, mra);
  uccl_engine_conn_destroy(ca);
  uccl_engine_conn_destroy(cb);
  uccl_engine_destroy(a);
  uccl_engine_destroy(b);
  if (g_fail) {
    fprintf(stderr, "uccl_engine_test: %d failures\n", g_fail);
    return 1;
  }
  printf("uccl_engine_test: OK\n");
  return 0;
}
