// ukernel plans over a MESSAGE transport: the executor for ranks that are not load/store reachable (another
// box).  Send(peer, dst-in-peer's-buffer) becomes a two-part message {header: buffer, offset, lane | payload}
// on the datagram transport; a receiver thread places every payload at the address the header names as it
// arrives and bumps the (peer, lane) arrival counter that Recv ops wait on -- one-sided plan semantics on top
// of two-sided messaging.
//
// Reference role: experimental/ukernel/src/transport/adapter/{tcp,uccl}_adapter.cc + communicator.cc (Send /
// Recv over sockets or the UCCL p2p engine with a host bounce pool) under the same CCL planner.
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../net/net_engine.h"
#include "uk_plan.h"

namespace ub {

struct UkNetConfig {
  int nlanes = 2;
  uint64_t tile_bytes = 1 << 20;
  int timeout_ms = 60000;
};

class UkNetComm {
 public:
  // flows[p]: an established flow of `engine` to rank p, dedicated to this communicator (flows[rank] unused)
  UkNetComm(int rank, int nranks, std::shared_ptr<net::Engine> engine, std::vector<uint32_t> flows, const UkNetConfig& cfg);
  ~UkNetComm();
  UkNetComm(const UkNetComm&) = delete;

  int rank() const { return rank_; }
  int nranks() const { return n_; }
  // blocking collectives on host memory (the caller stages device data)
  void all_reduce(const void* in, void* out, size_t count, int dtype, int op, UkAlgo algo = UkAlgo::Auto);
  void all_to_all(const void* in, void* out, size_t count_per_peer, int dtype);
  void all_gather(const void* in, void* out, size_t count_per_rank, int dtype);
  void reduce_scatter(const void* in, void* out, size_t recv_count, int dtype, int op);
  void broadcast(const void* in, void* out, size_t count, int dtype, int root);
  void barrier();
  struct Stats {
    uint64_t ops = 0, sends = 0, recvs = 0, bytes_sent = 0, parked_headers = 0;
  };
  Stats stats() const { return stats_; }

 private:
  struct WireHdr {  // 32 bytes, precedes every payload
    uint32_t magic, op_seq;
    uint32_t buf, lane;
    uint64_t off, bytes;
  };
  struct Peer {  // receive state machine of one peer (lives until process exit: the engine may still hold pointers)
    uint32_t flow = 0;
    WireHdr hdr{};
    net::Request* hdr_req = nullptr;
    net::Request* pay_req = nullptr;
    bool parked = false;
    std::atomic<bool> closed{false};  // the peer closed its side (normal at teardown; an error only if a Recv still needs it)
  };
  void run(const UkPlan& plan, char* in, char* out, int dtype, int op);
  void receiver();
  char* base(int buf) const;

  int rank_, n_;
  std::shared_ptr<net::Engine> eng_;
  UkNetConfig cfg_;
  std::vector<Peer*> peers_;
  std::thread rx_;
  std::atomic<bool> stop_{false};
  std::atomic<uint32_t> cur_seq_{0};
  std::atomic<int> rx_error_{0};
  uint64_t ext_[3] = {0, 0, 0};  // bytes of each buffer that a Recv of the op in flight may write (header bounds check)
  char* bases_[3] = {nullptr, nullptr, nullptr};  // In / Out / Scratch of the op in flight (valid while cur_seq_ names it)
  std::vector<char> scratch_;
  std::vector<std::atomic<uint64_t>> arrived_;  // [peer * nlanes + lane]
  std::vector<uint64_t> expected_;              // Recv ordinals consumed so far, same indexing
  Stats stats_;
};

}  // namespace ub
