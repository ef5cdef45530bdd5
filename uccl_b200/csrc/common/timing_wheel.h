// Timing wheel: O(1) scheduling of many deadlines with a fixed granularity (packet pacing, timers).
//
// Role in the reference: the Carousel-style pacing wheel of the RDMA transport (collective/rdma/timing_wheel.h,
// bypassed there by default, `transport_config.h:96-97`).  Slots are `granularity_ns` wide; an entry lands in the
// slot of its deadline (deadlines in the past go to the next slot to fire; deadlines beyond the horizon are
// clamped to the last slot and re-checked when it fires).  `advance(now)` returns everything whose time has come,
// in deadline order per slot.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace ub {

template <typename T>
class TimingWheel {
 public:
  TimingWheel(uint64_t granularity_ns, size_t slots, uint64_t now_ns)
      : gran_(granularity_ns ? granularity_ns : 1), wheel_(slots ? slots : 1), cur_(now_ns / gran_) {}

  uint64_t granularity_ns() const { return gran_; }
  uint64_t horizon_ns() const { return gran_ * wheel_.size(); }
  size_t size() const { return count_; }
  bool empty() const { return count_ == 0; }

  void insert(uint64_t deadline_ns, T item) {
    uint64_t tick = deadline_ns / gran_;
    if (tick <= cur_) tick = cur_ + 1;                                   // already due: next slot to fire
    if (tick >= cur_ + wheel_.size()) tick = cur_ + wheel_.size() - 1;   // beyond the horizon: re-armed on expiry
    wheel_[tick % wheel_.size()].push_back(Entry{deadline_ns, std::move(item)});
    ++count_;
  }

  // Moves the wheel to `now_ns`; appends every entry whose deadline has passed to `out` and returns their number.
  size_t advance(uint64_t now_ns, std::vector<T>* out) {
    const uint64_t target = now_ns / gran_;
    size_t fired = 0;
    // never walk more than one full revolution: older slots have been visited already
    uint64_t from = cur_ + 1;
    if (target >= cur_ + wheel_.size()) from = target - wheel_.size() + 1;
    for (uint64_t t = from; t <= target && count_ > 0; ++t) {
      auto& slot = wheel_[t % wheel_.size()];
      if (slot.empty()) continue;
      std::stable_sort(slot.begin(), slot.end(), [](const Entry& a, const Entry& b) { return a.deadline < b.deadline; });
      std::vector<Entry> keep;
      for (auto& e : slot) {
        if (e.deadline <= now_ns) {
          out->push_back(std::move(e.item));
          ++fired;
          --count_;
        } else {
          keep.push_back(std::move(e));  // clamped far deadline: stays for a later revolution
        }
      }
      slot.clear();
      cur_ = t;  // re-inserting below must see the slot being processed as the past
      for (auto& e : keep) {
        --count_;
        insert(e.deadline, std::move(e.item));
      }
    }
    if (target > cur_) cur_ = target;
    return fired;
  }

 private:
  struct Entry {
    uint64_t deadline;
    T item;
  };
  uint64_t gran_;
  std::vector<std::vector<Entry>> wheel_;
  uint64_t cur_;  // last tick that has been processed
  size_t count_ = 0;
};

}  // namespace ub
