// Cheap clocks, completion contexts and wrap-around sequence numbers
// (roles: include/util/util.h TSC helpers + PollCtx, UINT_CSN :310, include/util/timer.h).
#pragma once
#include <time.h>

#include <atomic>
#include <cstdint>
#include <thread>
#if defined(__x86_64__)
#include <x86intrin.h>
#endif

namespace ub {

inline uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

inline uint64_t rdtsc() {
#if defined(__x86_64__)
  return __rdtsc();
#else
  return now_ns();
#endif
}

// Calibrated TSC -> ns conversion (one 10 ms calibration per process).
inline double tsc_ghz() {
  static double ghz = [] {
    const uint64_t t0 = now_ns(), c0 = rdtsc();
    while (now_ns() - t0 < 10000000ull) {
    }
    const uint64_t t1 = now_ns(), c1 = rdtsc();
    return (double)(c1 - c0) / (double)(t1 - t0);
  }();
  return ghz;
}
inline uint64_t tsc_to_ns(uint64_t cycles) { return (uint64_t)((double)cycles / tsc_ghz()); }

// Completion flag handed from a worker to a poller (reference: PollCtx + fences, transport.h:1215-1222).
struct PollCtx {
  std::atomic<bool> done{false};
  std::atomic<uint64_t> bytes{0};
  uint64_t t_start_ns = 0;
  void signal(uint64_t n = 0) {
    bytes.store(n, std::memory_order_relaxed);
    done.store(true, std::memory_order_release);
  }
  bool poll() const { return done.load(std::memory_order_acquire); }
  void wait() const {
    uint32_t spins = 0;
    while (!poll())
      if ((++spins & 0xff) == 0) std::this_thread::yield();
  }
};

// Sequence number on B bits with wrap-around comparisons (like TCP serial arithmetic).
template <unsigned B>
struct SeqNo {
  static_assert(B >= 2 && B <= 32, "bits");
  static constexpr uint32_t kMask = (B == 32) ? 0xffffffffu : ((1u << B) - 1u);
  static constexpr uint32_t kHalf = 1u << (B - 1);
  uint32_t v = 0;
  SeqNo() = default;
  explicit SeqNo(uint32_t x) : v(x & kMask) {}
  SeqNo operator+(uint32_t d) const { return SeqNo(v + d); }
  SeqNo& operator++() {
    v = (v + 1) & kMask;
    return *this;
  }
  // signed distance this - o in (-2^(B-1), 2^(B-1)]
  int32_t diff(const SeqNo& o) const {
    uint32_t d = (v - o.v) & kMask;
    return d >= kHalf ? (int32_t)d - (int32_t)(kMask)-1 : (int32_t)d;
  }
  bool operator==(const SeqNo& o) const { return v == o.v; }
  bool operator<(const SeqNo& o) const { return diff(o) < 0; }
  bool operator<=(const SeqNo& o) const { return diff(o) <= 0; }
  bool operator>(const SeqNo& o) const { return diff(o) > 0; }
};

}  // namespace ub
