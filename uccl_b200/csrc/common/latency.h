// Latency histogram with bounded relative error (log2 major buckets x 16 linear minor buckets),
// percentile queries and a one-line summary -- role of the reference's Latency class
// (include/util/latency.h:21) and the per-op/per-verb latency macros of the EP proxy.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace ub {

class LatencyHist {
 public:
  LatencyHist() : buckets_(64 * kMinor, 0) {}
  void record(uint64_t v) {
    ++count_;
    sum_ += (double)v;
    if (v < min_) min_ = v;
    if (v > max_) max_ = v;
    ++buckets_[index(v)];
  }
  uint64_t count() const { return count_; }
  double mean() const { return count_ ? sum_ / (double)count_ : 0.0; }
  uint64_t min() const { return count_ ? min_ : 0; }
  uint64_t max() const { return max_; }
  // value v such that `p` percent of the samples are <= v (upper edge of the bucket)
  uint64_t percentile(double p) const {
    if (!count_) return 0;
    const uint64_t target = (uint64_t)((p / 100.0) * (double)count_ + 0.5);
    uint64_t acc = 0;
    for (size_t i = 0; i < buckets_.size(); ++i) {
      acc += buckets_[i];
      if (acc >= target && buckets_[i]) return std::min(upper_edge(i), max_);
    }
    return max_;
  }
  void merge(const LatencyHist& o) {
    for (size_t i = 0; i < buckets_.size(); ++i) buckets_[i] += o.buckets_[i];
    count_ += o.count_;
    sum_ += o.sum_;
    if (o.count_) {
      if (o.min_ < min_) min_ = o.min_;
      if (o.max_ > max_) max_ = o.max_;
    }
  }
  void reset() { *this = LatencyHist(); }
  std::string summary(const char* unit = "ns") const {
    char b[256];
    snprintf(b, sizeof(b), "n=%lu mean=%.1f%s min=%lu p50=%lu p90=%lu p99=%lu p99.9=%lu max=%lu", (unsigned long)count_,
             mean(), unit, (unsigned long)min(), (unsigned long)percentile(50), (unsigned long)percentile(90),
             (unsigned long)percentile(99), (unsigned long)percentile(99.9), (unsigned long)max_);
    return std::string(b);
  }

 private:
  static constexpr int kMinorBits = 4, kMinor = 1 << kMinorBits;
  static size_t index(uint64_t v) {
    if (v < (uint64_t)kMinor) return (size_t)v;
    const int msb = 63 - __builtin_clzll(v);
    const int shift = msb - kMinorBits;
    return (size_t)(((msb - kMinorBits + 1) << kMinorBits) + ((v >> shift) & (kMinor - 1)));
  }
  static uint64_t upper_edge(size_t i) {
    if (i < (size_t)kMinor) return (uint64_t)i;
    const int major = (int)(i >> kMinorBits), minor = (int)(i & (kMinor - 1));
    const int shift = major - 1;
    return ((((uint64_t)kMinor + minor + 1) << shift)) - 1;
  }
  std::vector<uint64_t> buckets_;
  uint64_t count_ = 0, min_ = ~0ull, max_ = 0;
  double sum_ = 0;
};

}  // namespace ub
