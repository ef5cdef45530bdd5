// Bounded lock-free rings: SPSC (one producer thread, one consumer thread) and MPMC
// (any number of each).  Same role as the reference's jring (include/util/jring.h) used for
// engine work queues; implementation here is a per-cell sequence design (Vyukov style) so that
// producers and consumers never share a CAS target with each other.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <new>

namespace ub {

constexpr size_t kCacheLine = 64;

inline size_t round_pow2(size_t n) {
  size_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

template <typename T>
class SpscRing {
 public:
  explicit SpscRing(size_t capacity) : cap_(round_pow2(capacity)), mask_(cap_ - 1), buf_(new T[cap_]) {}
  bool push(const T& v) {
    const size_t h = head_.load(std::memory_order_relaxed);
    if (h - tail_cache_ >= cap_) {
      tail_cache_ = tail_.load(std::memory_order_acquire);
      if (h - tail_cache_ >= cap_) return false;
    }
    buf_[h & mask_] = v;
    head_.store(h + 1, std::memory_order_release);
    return true;
  }
  bool pop(T* out) {
    const size_t t = tail_.load(std::memory_order_relaxed);
    if (t == head_cache_) {
      head_cache_ = head_.load(std::memory_order_acquire);
      if (t == head_cache_) return false;
    }
    *out = buf_[t & mask_];
    tail_.store(t + 1, std::memory_order_release);
    return true;
  }
  size_t size() const { return head_.load(std::memory_order_acquire) - tail_.load(std::memory_order_acquire); }
  size_t capacity() const { return cap_; }

 private:
  const size_t cap_, mask_;
  std::unique_ptr<T[]> buf_;
  alignas(kCacheLine) std::atomic<size_t> head_{0};
  size_t tail_cache_ = 0;
  alignas(kCacheLine) std::atomic<size_t> tail_{0};
  size_t head_cache_ = 0;
};

template <typename T>
class MpmcRing {
 public:
  explicit MpmcRing(size_t capacity) : cap_(round_pow2(capacity < 2 ? 2 : capacity)), mask_(cap_ - 1), cells_(new Cell[cap_]) {
    for (size_t i = 0; i < cap_; ++i) cells_[i].seq.store(i, std::memory_order_relaxed);
  }
  bool push(const T& v) {
    size_t pos = enq_.load(std::memory_order_relaxed);
    for (;;) {
      Cell& c = cells_[pos & mask_];
      const size_t seq = c.seq.load(std::memory_order_acquire);
      const intptr_t dif = (intptr_t)seq - (intptr_t)pos;
      if (dif == 0) {
        if (enq_.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) {
          c.val = v;
          c.seq.store(pos + 1, std::memory_order_release);
          return true;
        }
      } else if (dif < 0) {
        return false;  // full
      } else {
        pos = enq_.load(std::memory_order_relaxed);
      }
    }
  }
  bool pop(T* out) {
    size_t pos = deq_.load(std::memory_order_relaxed);
    for (;;) {
      Cell& c = cells_[pos & mask_];
      const size_t seq = c.seq.load(std::memory_order_acquire);
      const intptr_t dif = (intptr_t)seq - (intptr_t)(pos + 1);
      if (dif == 0) {
        if (deq_.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed)) {
          *out = c.val;
          c.seq.store(pos + cap_, std::memory_order_release);
          return true;
        }
      } else if (dif < 0) {
        return false;  // empty
      } else {
        pos = deq_.load(std::memory_order_relaxed);
      }
    }
  }
  // bulk helpers (best effort, return how many were moved)
  size_t push_bulk(const T* v, size_t n) {
    size_t i = 0;
    while (i < n && push(v[i])) ++i;
    return i;
  }
  size_t pop_bulk(T* v, size_t n) {
    size_t i = 0;
    while (i < n && pop(&v[i])) ++i;
    return i;
  }
  size_t capacity() const { return cap_; }
  size_t size_approx() const {
    const size_t e = enq_.load(std::memory_order_relaxed), d = deq_.load(std::memory_order_relaxed);
    return e >= d ? e - d : 0;
  }

 private:
  struct Cell {
    std::atomic<size_t> seq;
    T val;
  };
  const size_t cap_, mask_;
  std::unique_ptr<Cell[]> cells_;
  alignas(kCacheLine) std::atomic<size_t> enq_{0};
  alignas(kCacheLine) std::atomic<size_t> deq_{0};
};

}  // namespace ub
