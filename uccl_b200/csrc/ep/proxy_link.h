// ProxyLink: the network half of the CPU proxy -- what the reference's EP proxy does with RDMA verbs
// (ep/src/proxy.cpp + rdma.cpp: post a WRITE for every D2H command whose destination is on another node, then
// an atomic on the same QP so that the signal is ordered after the data).
//
// Here a link is one datagram flow per remote box, between rail-mates (the proxies of the same local rank).
// `put` + `add` issued by one proxy towards one box are applied by the remote proxy in issue order (flow FIFO):
// put-with-signal semantics.  The remote side applies a WRITE to the heap of ANY local rank of its box (the
// second, NVLink hop of DeepEP's internode scheme) through the callbacks it was constructed with.
#pragma once
#include <atomic>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../net/net_engine.h"

namespace ub {

struct ProxyLinkStats {
  uint64_t puts = 0, adds = 0, notifies = 0, bytes_out = 0;
  uint64_t applied_writes = 0, applied_adds = 0, applied_notifies = 0, bytes_in = 0;
};

class ProxyLink {
 public:
  using WriteFn = std::function<void(int dst_local, uint64_t dst_off, const void* data, uint32_t bytes)>;
  using AddFn = std::function<void(int dst_local, uint64_t dst_off, uint64_t value)>;
  using NotifyFn = std::function<void(int src_box, uint32_t a, uint32_t b)>;
  // flows[k]: established flow to the rail-mate proxy of box k (flows[box] unused)
  ProxyLink(int box, int nboxes, std::shared_ptr<net::Engine> engine, std::vector<uint32_t> flows, WriteFn w, AddFn a,
            NotifyFn n = nullptr);
  ~ProxyLink();
  ProxyLink(const ProxyLink&) = delete;

  int box() const { return box_; }
  int nboxes() const { return n_; }
  // `src` is copied before the call returns (it usually is a bounce buffer that is reused immediately)
  void put(int dst_box, int dst_local, uint64_t dst_off, const void* src, uint32_t bytes);
  void add(int dst_box, int dst_local, uint64_t dst_off, uint64_t value);
  void notify(int dst_box, uint32_t a, uint32_t b);
  // every message issued so far has been acknowledged by the transport
  void flush(int timeout_ms = 30000);
  ProxyLinkStats stats() const;

  // convenience target for host memory (tests, CPU-only runs): one flat heap per local rank
  static WriteFn host_write(std::vector<char*> heaps, uint64_t heap_bytes);
  static AddFn host_add(std::vector<char*> heaps, uint64_t heap_bytes);

 private:
  struct Hdr {  // 32 bytes
    uint32_t magic, kind;
    uint32_t dst_local, bytes;
    uint64_t off, value;
  };
  struct Pending {
    Hdr hdr;
    std::vector<char> payload;
    net::Request *h = nullptr, *p = nullptr;
  };
  struct Peer {
    uint32_t flow = 0;
    Hdr hdr{};
    std::vector<char> buf;
    net::Request *hdr_req = nullptr, *pay_req = nullptr;
    bool closed = false;
  };
  void post(int dst_box, const Hdr& h, const void* payload);
  void reap(bool all, int timeout_ms);
  void receiver();

  int box_, n_;
  std::shared_ptr<net::Engine> eng_;
  std::vector<Peer*> peers_;
  WriteFn write_;
  AddFn add_;
  NotifyFn notify_;
  std::thread rx_;
  std::atomic<bool> stop_{false};
  mutable std::mutex mu_;
  std::deque<std::unique_ptr<Pending>> pending_;
  ProxyLinkStats st_;
};

}  // namespace ub
