// Thread-per-rank stress of the host backend (the code path every CPU test and the GPU-less CI use): native
// collectives of `Comm` in host mode plus the ukernel communicator (task FIFOs drained by host worker threads),
// 4 ranks in ONE process so that ThreadSanitizer sees every cross-rank access.  Built against the sanitizer variant
// of the core by scripts/sanitize_host.sh.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "coll/comm.h"
#include "ukernel/uk_comm.h"

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
  const int n = 4, iters = argc > 1 ? std::atoi(argv[1]) : 30;
  CommConfig cfg;
  cfg.host_fake = true;
  cfg.heap_bytes = 160ull << 20;
  cfg.stage_bytes = 1ull << 20;
  std::vector<int> devs(n, -1);
  auto comms = Comm::create_local(devs, cfg);
  EXPECT((int)comms.size() == n);
  std::vector<std::thread> ts;
  for (int r = 0; r < n; ++r)
    ts.emplace_back([&, r] {
      auto& c = comms[r];
      const size_t count = 5000 + 16 * r * 0;  // same on every rank
      std::vector<float> x(count), y(count), g(count * n), a2a_in(count / n * n), a2a_out(count / n * n);
      UkCommConfig ucfg;
      ucfg.nlanes = 2, ucfg.tile_bytes = 4096, ucfg.staging_bytes = 64 << 10;
      UkComm uk(c, ucfg);
      for (int it = 0; it < iters; ++it) {
        for (size_t i = 0; i < count; ++i) x[i] = (float)(r + 1) + (float)(i % 7) + (float)it;
        c->allreduce(x.data(), y.data(), count, kF32, kSum, nullptr);
        bool ok = true;
        for (size_t i = 0; i < count && ok; ++i)
          ok = std::fabs(y[i] - ((float)(n * (n + 1) / 2) + (float)n * ((float)(i % 7) + (float)it))) < 1e-3f;
        EXPECT(ok);
        c->allgather(x.data(), g.data(), count, kF32, nullptr);
        EXPECT(g[(size_t)((r + 1) % n) * count] == (float)((r + 1) % n + 1) + (float)it);
        for (size_t i = 0; i < a2a_in.size(); ++i) a2a_in[i] = (float)(r * 1000) + (float)(i / (count / n));
        c->alltoall(a2a_in.data(), a2a_out.data(), count / n, kF32, nullptr);
        for (int s = 0; s < n; ++s) EXPECT(a2a_out[(size_t)s * (count / n)] == (float)(s * 1000 + r));
        c->broadcast(x.data(), y.data(), count, kF32, it % n, nullptr);
        EXPECT(y[3] == (float)(it % n + 1) + 3.f + (float)it);
        // grouped send/recv ring
        std::vector<float> sb(257, (float)r), rb(257, -1.f);
        c->group_p2p({{true, sb.data(), sb.size() * 4, (r + 1) % n}, {false, rb.data(), rb.size() * 4, (r + n - 1) % n}}, nullptr);
        EXPECT(rb[0] == (float)((r + n - 1) % n) && rb[256] == rb[0]);
        c->barrier(nullptr);
        // ukernel: planned all-reduce (ring and full mesh alternate) + all-to-all on the worker threads
        for (size_t i = 0; i < count; ++i) x[i] = (float)(r + 1);
        const uint64_t t1 = uk.all_reduce(x.data(), y.data(), count, kF32, kSum, it % 2 ? UkAlgo::Ring : UkAlgo::FullMesh, nullptr);
        uk.wait(t1);
        EXPECT(y[0] == (float)(n * (n + 1) / 2) && y[count - 1] == y[0]);
        const uint64_t t2 = uk.all_to_all(a2a_in.data(), a2a_out.data(), count / n, kF32, nullptr);
        uk.wait(t2);
        for (int s = 0; s < n; ++s) EXPECT(a2a_out[(size_t)s * (count / n)] == (float)(s * 1000 + r));
        uk.wait(uk.barrier(nullptr));
      }
      uk.stop();
    });
  for (auto& t : ts) t.join();
  if (g_fail) {
    std::fprintf(stderr, "host_world_stress: %d failures\n", g_fail.load());
    return 1;
  }
  std::printf("host_world_stress: OK (%d ranks x %d iterations)\n", n, iters);
  return 0;
}
