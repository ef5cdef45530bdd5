// EQDS-style receiver-driven credit pacer (Olteanu et al., NSDI'22).
// Role in the reference: include/cc/eqds.h:63 + collective/rdma/eqds.{h,cc} (pacer thread granting
// pull credits over a credit QP).  Here: a pure scheduler object -- senders register demand,
// the receiver grants credits at its ingress line rate, round-robin among active senders, with an
// "idle" list for senders whose demand is satisfied (they are re-activated by new demand).
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <unordered_map>
#include <vector>

namespace ub {
namespace cc {

struct EqdsConfig {
  double link_gbps = 7200.0;      // receiver ingress (900 GB/s)
  uint32_t credit_bytes = 65536;  // bytes granted per credit
  uint32_t max_backlog_credits = 64;  // outstanding un-used credit per sender
};

class EqdsPacer {
 public:
  explicit EqdsPacer(const EqdsConfig& c = EqdsConfig()) : cfg_(c) {}
  // sender announces `bytes` more to send (speculative first window is the caller's business)
  void add_demand(uint32_t sender, uint64_t bytes) {
    auto& s = senders_[sender];
    const bool was_idle = s.demand == 0;
    s.demand += bytes;
    if (was_idle && !s.queued) {
      active_.push_back(sender);
      s.queued = true;
    }
  }
  // the sender's flow is gone (closed or failed): forget its demand so it neither blocks the grant loop nor
  // keeps the engine thread in its pacing nap
  void remove(uint32_t sender) {
    if (senders_.erase(sender) == 0) return;
    for (auto it = active_.begin(); it != active_.end();) it = (*it == sender) ? active_.erase(it) : std::next(it);
  }
  // sender consumed previously granted credit
  void on_data(uint32_t sender, uint64_t bytes) {
    auto it = senders_.find(sender);
    if (it == senders_.end()) return;
    it->second.unused = it->second.unused > bytes ? it->second.unused - bytes : 0;
  }
  // advance the pacer clock to `now_us`; returns the (sender, bytes) grants issued in this tick
  std::vector<std::pair<uint32_t, uint32_t>> tick(double now_us) {
    std::vector<std::pair<uint32_t, uint32_t>> grants;
    if (last_us_ < 0) last_us_ = now_us;
    budget_bytes_ += (now_us - last_us_) * cfg_.link_gbps * 1e3 / 8.0;
    last_us_ = now_us;
    // never bank more than a handful of credits of idle time
    budget_bytes_ = std::min(budget_bytes_, (double)cfg_.credit_bytes * 8);
    size_t stalled = 0;
    while (budget_bytes_ >= cfg_.credit_bytes && !active_.empty() && stalled < active_.size()) {
      const uint32_t id = active_.front();
      active_.pop_front();
      auto& s = senders_[id];
      if (s.demand == 0) {
        s.queued = false;  // goes idle
        stalled = 0;
        continue;
      }
      if (s.unused >= (uint64_t)cfg_.max_backlog_credits * cfg_.credit_bytes) {
        active_.push_back(id);  // has plenty of unused credit: skip this round
        ++stalled;
        continue;
      }
      const uint32_t g = (uint32_t)std::min<uint64_t>(cfg_.credit_bytes, s.demand);
      s.demand -= g;
      s.unused += g;
      s.granted += g;
      budget_bytes_ -= cfg_.credit_bytes;
      grants.emplace_back(id, g);
      stalled = 0;
      if (s.demand > 0) active_.push_back(id);
      else s.queued = false;
    }
    return grants;
  }
  uint64_t granted(uint32_t sender) const {
    auto it = senders_.find(sender);
    return it == senders_.end() ? 0 : it->second.granted;
  }
  size_t active_senders() const { return active_.size(); }

 private:
  struct Sender {
    uint64_t demand = 0, unused = 0, granted = 0;
    bool queued = false;
  };
  EqdsConfig cfg_;
  std::unordered_map<uint32_t, Sender> senders_;
  std::deque<uint32_t> active_;
  double last_us_ = -1.0, budget_bytes_ = 0.0;
};

}  // namespace cc
}  // namespace ub
