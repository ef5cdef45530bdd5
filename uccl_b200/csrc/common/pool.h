// Object pool with per-thread caches in front of a shared MPMC ring
// (role of the reference's SharedPool, include/util/shared_pool.h): the hot path touches only
// the thread-local stack; refills/flushes move half a cache at a time.
#pragma once
#include <vector>

#include "ring.h"

namespace ub {

template <typename T, size_t kCache = 32>
class SharedPool {
 public:
  explicit SharedPool(size_t capacity) : global_(capacity) {}
  ~SharedPool() = default;
  // seed / return to the shared ring directly (any thread)
  bool release_global(const T& v) { return global_.push(v); }
  void put(const T& v) {
    Cache& c = cache();
    if (c.n == kCache) {
      // flush the older half
      size_t moved = 0;
      for (size_t i = 0; i < kCache / 2; ++i)
        if (global_.push(c.items[i])) ++moved;
        else break;
      for (size_t i = moved; i < c.n; ++i) c.items[i - moved] = c.items[i];
      c.n -= moved;
      if (c.n == kCache) return;  // pool over capacity: drop
    }
    c.items[c.n++] = v;
  }
  bool get(T* out) {
    Cache& c = cache();
    if (c.n == 0) {
      for (size_t i = 0; i < kCache / 2; ++i) {
        T v;
        if (!global_.pop(&v)) break;
        c.items[c.n++] = v;
      }
      if (c.n == 0) return false;
    }
    *out = c.items[--c.n];
    return true;
  }
  size_t global_size() const { return global_.size_approx(); }

 private:
  struct Cache {
    T items[kCache];
    size_t n = 0;
  };
  Cache& cache() {
    thread_local std::vector<std::pair<const void*, Cache>> caches;
    for (auto& kv : caches)
      if (kv.first == this) return kv.second;
    caches.emplace_back(this, Cache());
    return caches.back().second;
  }
  MpmcRing<T> global_;
};

}  // namespace ub
