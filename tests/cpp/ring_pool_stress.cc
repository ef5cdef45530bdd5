// Lock-free building blocks under ThreadSanitizer: SPSC ring (one producer, one consumer), MPMC ring (4 x 4) and the
// thread-cached SharedPool on top of it.  Checks that every item is delivered exactly once and -- per producer --
// in order.  (Reference role: include/util/jring.h / shared_pool.h and util_test.cc.)
//   g++ -std=c++17 -O1 -g -fsanitize=thread -Iuccl_b200/csrc tests/cpp/ring_pool_stress.cc -lpthread
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "common/pool.h"
#include "common/ring.h"

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
  const uint64_t N = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 200000;
  {  // ---- SPSC
    SpscRing<uint64_t> r(64);
    std::thread prod([&] {
      for (uint64_t i = 0; i < N;)
        if (r.push(i)) ++i;
    });
    uint64_t expect = 0, v = 0;
    while (expect < N)
      if (r.pop(&v)) {
        EXPECT(v == expect);
        ++expect;
      }
    prod.join();
    EXPECT(r.size() == 0);
  }
  {  // ---- MPMC: P producers each push (id << 32 | seq); consumers check per-producer order and the total
    constexpr int P = 4, C = 4;
    MpmcRing<uint64_t> r(128);
    std::atomic<uint64_t> popped{0}, sum{0};
    std::vector<std::thread> ts;
    for (int p = 0; p < P; ++p)
      ts.emplace_back([&, p] {
        for (uint64_t i = 0; i < N / P;)
          if (r.push(((uint64_t)p << 32) | i)) ++i;
          else std::this_thread::yield();
      });
    for (int c = 0; c < C; ++c)
      ts.emplace_back([&] {
        uint64_t last[P];
        bool seen[P] = {false};
        uint64_t v;
        while (popped.load(std::memory_order_relaxed) < (N / P) * P) {
          if (!r.pop(&v)) {
            std::this_thread::yield();
            continue;
          }
          const int p = (int)(v >> 32);
          const uint64_t s = v & 0xffffffffu;
          if (seen[p]) EXPECT(s > last[p]);  // one consumer sees any producer's items in increasing order
          seen[p] = true, last[p] = s;
          sum.fetch_add(s, std::memory_order_relaxed);
          popped.fetch_add(1, std::memory_order_relaxed);
        }
      });
    for (auto& t : ts) t.join();
    const uint64_t per = N / P;
    EXPECT(popped == per * P && sum == P * (per * (per - 1) / 2));
    uint64_t v;
    EXPECT(!r.pop(&v));
  }
  {  // ---- SharedPool: tokens circulate between threads; none is lost or duplicated
    constexpr int T = 6;
    constexpr uint64_t TOK = 256;
    SharedPool<uint64_t, 16> pool(1024);
    for (uint64_t i = 0; i < TOK; ++i) EXPECT(pool.release_global(i));
    std::vector<std::atomic<int>> owner(TOK);
    for (auto& o : owner) o = 0;
    std::vector<std::thread> ts;
    for (int t = 0; t < T; ++t)
      ts.emplace_back([&] {
        std::vector<uint64_t> mine;
        for (uint64_t it = 0; it < N / 20; ++it) {
          uint64_t v;
          if (mine.size() < 8 && pool.get(&v)) {
            EXPECT(owner[v].fetch_add(1) == 0);  // nobody else holds it
            mine.push_back(v);
          } else if (!mine.empty()) {
            v = mine.back();
            mine.pop_back();
            EXPECT(owner[v].fetch_sub(1) == 1);
            pool.put(v);
          }
        }
        for (uint64_t v : mine) {
          EXPECT(owner[v].fetch_sub(1) == 1);
          pool.release_global(v);  // hand back to the shared ring so that the final count is exact
        }
      });
    for (auto& t : ts) t.join();
    for (auto& o : owner) EXPECT(o == 0);
  }
  if (g_fail) {
    std::fprintf(stderr, "ring_pool_stress: %d failures\n", g_fail.load());
    return 1;
  }
  std::printf("ring_pool_stress: OK\n");
  return 0;
}
