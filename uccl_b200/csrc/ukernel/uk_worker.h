// Host side of the ukernel worker: owns the per-lane FIFOs, launches / stops the persistent
// kernel, produces tasks and observes completion.
//
// Two execution modes share the FIFO protocol byte for byte:
//   * device: `uk_worker_kernel` (uk_worker_kernel.cu) on a private high-priority stream;
//   * host:   one CPU thread per lane runs the same task interpreter over host memory -- the
//             backend used by the GPU-less CI (role of the reference's MockDeviceBackend,
//             experimental/ukernel/src/ccl/test/common/backend_test_utils.h:240).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "uk_task.h"

namespace ub {

struct UkWorkerStats {
  uint64_t pushed = 0, copies = 0, reduces = 0, signals = 0, waits = 0, bytes = 0;
};

class UkWorker {
 public:
  // device < 0 selects host mode.  idle_us: the device worker quits after this long without work and
  // is relaunched on demand (-1: UCCL_B200_UK_IDLE_US, default 200 ms; 0: stay resident).
  UkWorker(int device, int nlanes, uint64_t timeout_ms = 20000, int64_t idle_us = -1);
  ~UkWorker();
  UkWorker(const UkWorker&) = delete;

  // wait_for: the worker kernel starts after this event (e.g. a cross-rank barrier that publishes
  // "every rank has initialised its counters")
  void start(cudaEvent_t wait_for = nullptr);
  void stop();  // pushes UK_EXIT to every lane and joins / synchronises
  bool running() const { return running_; }
  bool is_host() const { return device_ < 0; }
  int device() const { return device_; }
  int nlanes() const { return nlanes_; }

  // Enqueue one task on a lane; returns the lane ticket (number of tasks pushed so far).
  // Blocks while the ring is full.
  uint64_t push(int lane, UkTask t);
  uint64_t pushed(int lane) const { return lanes_[lane].pushed.load(std::memory_order_acquire); }
  uint64_t kernel_launches() const { return launches_.load(); }
  uint64_t completed(int lane) const;
  bool done(int lane, uint64_t ticket) const { return completed(lane) >= ticket; }
  void wait(int lane, uint64_t ticket, double timeout_s = 30.0) const;
  void wait_all(double timeout_s = 30.0) const;
  // device address of the lane's completion counter (cuStreamWaitValue64 target); host mode: host address
  uint64_t* done_counter(int lane) const { return lanes_[lane].done_dev; }
  uint32_t error() const;
  UkWorkerStats stats() const { return stats_; }

 private:
  struct Lane {
    UkTask* ring = nullptr;         // pinned host memory
    UkTask* ring_dev = nullptr;     // its device alias
    uint64_t* done_host = nullptr;  // pinned
    uint64_t* done_host_dev = nullptr;
    uint64_t* done_dev = nullptr;   // device memory (host mode: == done_host)
    std::atomic<uint64_t> pushed{0};
    uint64_t start = 0;
  };
  void host_lane_loop(int lane);
  bool pending() const;
  void ensure_live_locked();  // (re)launch the kernel if it has quit and tasks are waiting
  void kick();
  void watchdog_loop();
  int device_;
  int nlanes_;
  uint64_t timeout_ns_;
  bool running_ = false;
  std::unique_ptr<Lane[]> lanes_;
  uint64_t idle_ns_ = 0;
  uint32_t* votes_dev_ = nullptr;
  mutable std::mutex launch_mu_;
  bool kernel_live_ = false;  // guarded by launch_mu_
  std::atomic<bool> live_hint_{false};
  std::atomic<uint64_t> launches_{0};
  std::thread watchdog_;
  std::atomic<bool> wd_stop_{false};
  uint32_t push_tick_ = 0;
  uint32_t* err_host_ = nullptr;
  uint32_t* err_dev_ = nullptr;
  cudaStream_t stream_ = nullptr;
  std::vector<std::thread> host_threads_;
  UkWorkerStats stats_;
};

cudaError_t launch_uk_worker(const UkWorkerArgs& w, cudaStream_t st);

}  // namespace ub
