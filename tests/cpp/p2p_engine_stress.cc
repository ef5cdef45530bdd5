// Concurrency stress of the P2P engine in host mode (two endpoints in one process), meant to run under
// ThreadSanitizer: a two-sided stream, one-sided vector writes / reads, notifications and connection churn all at the
// same time on separate connections, each driven by its own pair of threads while the two engine threads do the
// matching.  Also runs in normal builds as a functional test (payload patterns are checked).
//   g++ -std=c++17 -O1 -g -fsanitize=thread -Iuccl_b200/csrc -Iuccl_b200/csrc/p2p -I/usr/local/cuda/include \
//       tests/cpp/p2p_engine_stress.cc tests/cpp/p2p_kernel_stub.cc uccl_b200/csrc/p2p/endpoint.cc -lcudart -lpthread
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "endpoint.h"

using namespace ub;

static std::atomic<int> g_fail{0};
#define EXPECT(c)                                                  \
  do {                                                             \
    if (!(c)) {                                                    \
      std::fprintf(stderr, "FAILED %s @%d\n", #c, __LINE__);       \
      ++g_fail;                                                    \
    }                                                              \
  } while (0)

struct Pair {
  uint64_t a = 0, b = 0;  // the connection as seen by the initiator / by the target
};

static Pair link(Endpoint& A, Endpoint& B) {
  Pair p;
  std::string ip;
  uint16_t port = 0;
  int gpu = 0;
  EXPECT(Endpoint::parse_metadata(B.get_metadata(), &ip, &port, &gpu));
  EXPECT(A.connect(ip, gpu, port, &p.a));
  std::string rip;
  int rgpu = 0;
  EXPECT(B.accept(&rip, &rgpu, &p.b, 20000));
  return p;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? std::atoi(argv[1]) : 150;
  Endpoint A(-1, 2), B(-1, 2);
  Pair two = link(A, B), one = link(A, B), note = link(A, B);

  // ---- two-sided stream A -> B, messages of varying size, strictly ordered per connection
  std::thread tx([&] {
    std::vector<unsigned char> buf(1 << 16);
    for (int i = 0; i < iters; ++i) {
      const size_t n = 64 + (size_t)(i * 977) % (buf.size() - 64);
      for (size_t k = 0; k < n; ++k) buf[k] = (unsigned char)(i + k);
      uint64_t tid = 0;
      EXPECT(A.send_async(two.a, {buf.data()}, {n}, &tid));
      EXPECT(A.wait(tid, 30000));
    }
  });
  std::thread rx([&] {
    std::vector<unsigned char> buf(1 << 16);
    for (int i = 0; i < iters; ++i) {
      const size_t n = 64 + (size_t)(i * 977) % (buf.size() - 64);
      std::memset(buf.data(), 0, n);
      uint64_t tid = 0;
      EXPECT(B.recv_async(two.b, {buf.data()}, {n}, &tid));
      EXPECT(B.wait(tid, 30000));
      bool ok = true;
      for (size_t k = 0; k < n && ok; ++k) ok = buf[k] == (unsigned char)(i + k);
      EXPECT(ok);
    }
  });

  // ---- one-sided: A writes vectors into windows B describes, then reads them back
  std::thread os([&] {
    std::vector<unsigned char> w0(5000), w1(3000), src(8000), back(8000);
    uint64_t m0 = 0, m1 = 0;
    EXPECT(B.reg(w0.data(), w0.size(), &m0) && B.reg(w1.data(), w1.size(), &m1));
    XferDesc d0, d1;
    EXPECT(B.describe(w0.data(), w0.size(), &d0) && B.describe(w1.data(), w1.size(), &d1));
    for (int i = 0; i < iters; ++i) {
      for (size_t k = 0; k < src.size(); ++k) src[k] = (unsigned char)(3 * i + k);
      uint64_t tid = 0;
      EXPECT(A.write_async(one.a, {src.data(), src.data() + 5000}, {5000, 3000}, {d0, d1}, &tid));
      EXPECT(A.wait(tid, 30000));
      EXPECT(A.read_async(one.a, {back.data(), back.data() + 5000}, {5000, 3000}, {d0, d1}, &tid));
      EXPECT(A.wait(tid, 30000));
      EXPECT(std::memcmp(back.data(), src.data(), src.size()) == 0);
    }
    EXPECT(B.dereg(m0) && B.dereg(m1));
  });

  // ---- notifications A -> B while B drains them
  std::atomic<int> seen{0};
  std::thread ntx([&] {
    for (int i = 0; i < iters; ++i) EXPECT(A.send_notif(note.a, "n" + std::to_string(i)));
  });
  std::thread nrx([&] {
    const auto t0 = std::chrono::steady_clock::now();
    while (seen < iters && std::chrono::steady_clock::now() - t0 < std::chrono::seconds(60)) {
      seen += (int)B.get_notifs().size();
      std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
  });

  // ---- connection churn: connect / accept / use once / remove, while everything above is running
  std::thread churn([&] {
    unsigned char x[256], y[256];
    for (int i = 0; i < iters / 5 + 1; ++i) {
      Pair p = link(A, B);
      std::memset(x, i, sizeof(x));
      uint64_t ts = 0, tr = 0;
      EXPECT(B.recv_async(p.b, {y}, {sizeof(y)}, &tr));
      EXPECT(A.send_async(p.a, {x}, {sizeof(x)}, &ts));
      EXPECT(A.wait(ts, 30000) && B.wait(tr, 30000));
      EXPECT(std::memcmp(x, y, sizeof(x)) == 0);
      EXPECT(A.remove_remote_endpoint(p.a));
      EXPECT(B.remove_remote_endpoint(p.b));
    }
  });

  tx.join(), rx.join(), os.join(), ntx.join(), nrx.join(), churn.join();
  EXPECT(seen == iters);
  const P2PStats sa = A.stats(), sb = B.stats();
  EXPECT(sa.transfers > 0 && sb.transfers > 0);
  if (g_fail) {
    std::fprintf(stderr, "p2p_engine_stress: %d failures\n", g_fail.load());
    return 1;
  }
  std::printf("p2p_engine_stress: OK (%d iterations)\n", iters);
  return 0;
}
