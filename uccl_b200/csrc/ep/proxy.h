// CPU proxy: consumes D2HCmd from a GPU->CPU queue and executes them (copy-engine writes into any
// peer's heap, ordered remote atomics, notifications).  Reference role: ep/src/proxy.cpp (RDMA
// posting threads); here the "network" is cudaMemcpyAsync over peer-mapped VAs.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../coll/comm.h"
#include "../common/latency.h"
#include "d2h_queue.h"
#include "proxy_link.h"

namespace ub {

struct ProxyStats {
  uint64_t cmds = 0, nops = 0, writes = 0, atomics = 0, notifies = 0, bytes = 0;
  double avg_handle_us = 0;  // mean CPU time per command
};

class Proxy {
 public:
  Proxy(std::shared_ptr<Comm> comm, uint32_t capacity = 4096);
  ~Proxy();
  Proxy(const Proxy&) = delete;

  D2HQueueDev queue() const { return dev_; }
  void start();
  void stop();
  bool running() const { return running_; }
  // wait until every command pushed so far (as seen by the device head) is consumed and its copies done
  void drain(double timeout_s = 30.0);
  uint64_t consumed() const { return __atomic_load_n(tail_, __ATOMIC_ACQUIRE); }
  std::vector<std::pair<uint32_t, uint32_t>> poll_notifications();
  ProxyStats stats() const;

  // Groups that span boxes: commands whose destination rank (a GLOBAL rank, box-major) lives on another box are
  // forwarded over the datagram transport to the rail-mate proxy there, which applies them to the heap of the
  // final local rank over NVLink -- the reference's "proxy posts RDMA" half.  flows[k] = flow to box k's rail-mate.
  void attach_link(std::shared_ptr<net::Engine> engine, std::vector<uint32_t> flows, int box, int nboxes, int local_size);
  ProxyLink* link() const { return link_.get(); }

  // microbenchmarks (ep/src/bench_kernel.cu role): returns {commands/s} resp. {mean round trip in us}
  double bench_throughput(int blocks, int threads, int per_thread, cudaStream_t st);
  double bench_latency(int iters, cudaStream_t st);
  // test helper: one thread issues a WRITE (+ optional ATOMIC) from a kernel
  void issue_from_device(uint32_t type, int dst_rank, uint32_t aux, uint64_t src_off, uint64_t dst_off, uint32_t bytes,
                         uint32_t value, cudaStream_t st);

 private:
  void loop();
  void handle(const D2HCmd& c);
  std::shared_ptr<Comm> comm_;
  D2HQueueDev dev_{};
  D2HCmd* ring_ = nullptr;     // host alias
  uint64_t* tail_ = nullptr;   // host alias (pinned)
  uint64_t* ack_ = nullptr;
  uint32_t capacity_ = 0;
  std::thread th_;
  std::atomic<bool> stop_{false};
  bool running_ = false;
  cudaStream_t stream_ = nullptr;
  mutable std::mutex mu_;
  std::deque<std::pair<uint32_t, uint32_t>> notifs_;
  ProxyStats stats_;
  double handle_us_sum_ = 0;
  std::unique_ptr<ProxyLink> link_;
  int box_ = 0, local_size_ = 0;
  char* bounce_ = nullptr;  // pinned staging for outbound remote WRITEs
  size_t bounce_cap_ = 0;
  cudaStream_t link_stream_ = nullptr;  // inbound remote WRITE / ATOMIC are applied on this stream (in order)
};

cudaError_t launch_d2h_bench(const D2HQueueDev& q, int blocks, int threads, int per_thread, cudaStream_t st);
cudaError_t launch_d2h_latency(const D2HQueueDev& q, int iters, unsigned long long* total_ns, cudaStream_t st);
cudaError_t launch_d2h_issue(const D2HQueueDev& q, uint32_t type, uint32_t dst_rank, uint32_t aux, uint64_t src_off,
                             uint64_t dst_off, uint32_t bytes, uint32_t value, cudaStream_t st);
cudaError_t launch_u64_add(uint64_t* p, uint64_t v, cudaStream_t st);

}  // namespace ub
