// UkNetComm (ukernel plans over the datagram transport) with 3 ranks in ONE process -- one engine, one receiver
// thread and one caller thread per rank -- so that ThreadSanitizer sees the receiver / caller hand-over of every
// collective.  Built against the sanitizer variant of the core by scripts/sanitize_host.sh.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <memory>
#include <thread>
#include <vector>

#include "kernels/types.h"
#include "net/net_engine.h"
#include "ukernel/uk_net.h"

using namespace ub;

static std::atomic<int> g_fail{0};
#define EXPECT(c)                                              \
  do {                                                         \
    if (!(c)) {                                                \
      std::fprintf(stderr, "FAILED %s @%d\n", #c, __LINE__);   \
      ++g_fail;                                                \
    }                                                          \
  } while (0)

int main(int argc, char** argv) {
  const int n = 3, iters = argc > 1 ? std::atoi(argv[1]) : 20;
  net::EngineConfig ec = net::EngineConfig::from_env();
  ec.paths = 4;
  std::vector<std::shared_ptr<net::Engine>> eng;
  std::vector<uint32_t> lid;
  for (int r = 0; r < n; ++r) {
    eng.push_back(std::make_shared<net::Engine>(ec));
    lid.push_back(eng[r]->listen());
  }
  // full mesh: the lower rank connects, the higher rank accepts (in rank order, so accepts match connects)
  std::vector<std::vector<uint32_t>> flows(n, std::vector<uint32_t>(n, 0));
  for (int a = 0; a < n; ++a)
    for (int b = a + 1; b < n; ++b) {
      flows[a][b] = eng[a]->connect("127.0.0.1", eng[b]->port(), lid[b], 20000);
      flows[b][a] = eng[b]->accept(lid[b], 20000);
      EXPECT(flows[a][b] != 0 && flows[b][a] != 0);
    }
  std::vector<std::thread> ts;
  for (int r = 0; r < n; ++r)
    ts.emplace_back([&, r] {
      UkNetConfig cfg;
      cfg.nlanes = 2, cfg.tile_bytes = 8192;
      UkNetComm c(r, n, eng[r], flows[r], cfg);
      const size_t count = 6000;
      std::vector<float> x(count), y(count), g(count * n), rs(count / n);
      for (int it = 0; it < iters; ++it) {
        for (size_t i = 0; i < count; ++i) x[i] = (float)(r + 1) * (float)(1 + i % 5) + (float)it;
        c.all_reduce(x.data(), y.data(), count, kF32, kSum, it % 2 ? UkAlgo::Ring : UkAlgo::FullMesh);
        bool ok = true;
        for (size_t i = 0; i < count && ok; ++i)
          ok = std::fabs(y[i] - ((float)(n * (n + 1) / 2) * (float)(1 + i % 5) + (float)(n * it))) < 1e-2f;
        EXPECT(ok);
        c.all_gather(x.data(), g.data(), count, kF32);
        EXPECT(g[(size_t)((r + 1) % n) * count] == (float)((r + 1) % n + 1) + (float)it);
        c.broadcast(x.data(), y.data(), count, kF32, it % n);
        EXPECT(y[1] == (float)(it % n + 1) * 2.f + (float)it);
        c.reduce_scatter(x.data(), rs.data(), count / n, kF32, kMax);
        EXPECT(rs[0] == (float)n * (float)(1 + (r * (count / n)) % 5) + (float)it);
        c.barrier();
      }
      EXPECT(c.stats().ops >= (uint64_t)iters * 5);
    });
  for (auto& t : ts) t.join();
  for (auto& e : eng) e->shutdown(200);
  if (g_fail) {
    std::fprintf(stderr, "uk_net_stress: %d failures\n", g_fail.load());
    return 1;
  }
  std::printf("uk_net_stress: OK (%d ranks x %d iterations)\n", n, iters);
  return 0;
}
