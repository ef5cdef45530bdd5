// Timely: RTT-gradient rate control (Mittal et al., SIGCOMM'15).
// Role in the reference: include/cc/timely.h:48 (default sender CC of the RDMA transport).
// On one NVSwitch node there is no congested fabric to react to; the class is kept as a
// reusable pacing policy (e.g. for rate-limiting background KV moves) and for API parity.
#pragma once
#include <algorithm>
#include <cstdint>

namespace ub {
namespace cc {

struct TimelyConfig {
  double min_rtt_us = 2.0;       // propagation floor (NVLink peer round trip is ~2 us)
  double t_low_us = 6.0;         // below: additive increase regardless of gradient
  double t_high_us = 60.0;       // above: multiplicative decrease regardless of gradient
  double ewma_alpha = 0.46;
  double beta = 0.26;            // multiplicative decrease factor
  double add_step_gbps = 5.0;    // additive increase step
  double min_rate_gbps = 1.0;
  double link_gbps = 7200.0;     // 900 GB/s
  int hai_threshold = 5;         // consecutive "good" RTTs before hyper-active increase
};

class Timely {
 public:
  explicit Timely(const TimelyConfig& c = TimelyConfig()) : cfg_(c), rate_(c.link_gbps) {}
  double rate_gbps() const { return rate_; }
  double rtt_diff_us() const { return rtt_diff_; }
  // feed one RTT sample; returns the new sending rate
  double on_rtt(double rtt_us) {
    if (prev_rtt_ < 0) {
      prev_rtt_ = rtt_us;
      return rate_;
    }
    const double new_diff = rtt_us - prev_rtt_;
    prev_rtt_ = rtt_us;
    rtt_diff_ = (1.0 - cfg_.ewma_alpha) * rtt_diff_ + cfg_.ewma_alpha * new_diff;
    const double gradient = rtt_diff_ / cfg_.min_rtt_us;
    if (rtt_us < cfg_.t_low_us) {
      increase();
    } else if (rtt_us > cfg_.t_high_us) {
      rate_ *= (1.0 - cfg_.beta * (1.0 - cfg_.t_high_us / rtt_us));
      good_ = 0;
    } else if (gradient <= 0) {
      increase();
    } else {
      rate_ *= (1.0 - cfg_.beta * std::min(gradient, 1.0));
      good_ = 0;
    }
    rate_ = std::max(cfg_.min_rate_gbps, std::min(rate_, cfg_.link_gbps));
    return rate_;
  }
  // inter-packet gap for pacing a chunk of `bytes`
  double pacing_delay_us(uint64_t bytes) const { return (double)bytes * 8.0 / (rate_ * 1e3); }

 private:
  void increase() {
    ++good_;
    const int n = good_ >= cfg_.hai_threshold ? 5 : 1;  // HAI mode
    rate_ += n * cfg_.add_step_gbps;
  }
  TimelyConfig cfg_;
  double rate_;
  double prev_rtt_ = -1.0, rtt_diff_ = 0.0;
  int good_ = 0;
};

}  // namespace cc
}  // namespace ub
