// Swift: delay-target congestion window (Kumar et al., SIGCOMM'20).
// Role in the reference: include/cc/swift.h:34.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>

namespace ub {
namespace cc {

struct SwiftConfig {
  double base_target_us = 8.0;   // base fabric delay target
  double hop_scale_us = 1.0;     // per-hop addition
  double fs_range_us = 20.0;     // flow-scaling range
  double fs_min_cwnd = 0.1, fs_max_cwnd = 256.0;
  double ai = 1.0;               // additive increase per RTT (packets)
  double beta = 0.8;             // multiplicative decrease gain
  double max_mdf = 0.5;          // max multiplicative decrease per RTT
  double min_cwnd = 0.01, max_cwnd = 1024.0;
};

class Swift {
 public:
  explicit Swift(const SwiftConfig& c = SwiftConfig()) : cfg_(c), cwnd_(16.0) {
    // flow scaling: target grows as cwnd shrinks (alpha/sqrt(cwnd) + beta form)
    const double a = 1.0 / std::sqrt(cfg_.fs_min_cwnd) - 1.0 / std::sqrt(cfg_.fs_max_cwnd);
    fs_alpha_ = cfg_.fs_range_us / a;
    fs_beta_ = -fs_alpha_ / std::sqrt(cfg_.fs_max_cwnd);
  }
  double cwnd() const { return cwnd_; }
  double target_delay_us(int hops = 1) const {
    double fs = fs_alpha_ / std::sqrt(std::max(cwnd_, cfg_.fs_min_cwnd)) + fs_beta_;
    fs = std::max(0.0, std::min(fs, cfg_.fs_range_us));
    return cfg_.base_target_us + hops * cfg_.hop_scale_us + fs;
  }
  // one ACK carrying a fabric delay sample; `acked` packets; `now_us` for the once-per-RTT decrease guard
  double on_ack(double delay_us, double acked, double now_us, double rtt_us, int hops = 1) {
    const double target = target_delay_us(hops);
    if (delay_us < target) {
      if (cwnd_ >= 1.0) cwnd_ += cfg_.ai * acked / cwnd_;
      else cwnd_ += cfg_.ai * acked;
    } else if (now_us - last_decrease_us_ >= rtt_us) {
      const double f = std::max(1.0 - cfg_.beta * (delay_us - target) / delay_us, 1.0 - cfg_.max_mdf);
      cwnd_ *= f;
      last_decrease_us_ = now_us;
    }
    cwnd_ = std::max(cfg_.min_cwnd, std::min(cwnd_, cfg_.max_cwnd));
    return cwnd_;
  }
  void on_retransmit_timeout() { cwnd_ = cfg_.min_cwnd; }
  // cwnd < 1 is realised by pacing one packet every rtt / cwnd
  double pacing_delay_us(double rtt_us) const { return cwnd_ < 1.0 ? rtt_us / cwnd_ : 0.0; }

 private:
  SwiftConfig cfg_;
  double cwnd_;
  double fs_alpha_ = 0, fs_beta_ = 0;
  double last_decrease_us_ = -1e18;
};

}  // namespace cc
}  // namespace ub
