// Out-of-band bootstrap for one node: a TCP rendezvous root (allgather / barrier /
// broadcast of small blobs) plus SCM_RIGHTS file-descriptor passing over abstract
// unix sockets (needed to share cuMem VMM / multicast handles between processes).
//
// Reference behaviour this replaces: lite's TcpBootstrap (experimental/lite/core/bootstrap.cc),
// the unix-socket fd server of gpu_ipc_mem.cc:264,319-321, and the socket helpers of
// include/util/util.h:73-287.  Design is ours: a relay thread in the id-creating
// process, star topology (control traffic is tiny on one node).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace ub {

struct UniqueId {
  char data[128];
};

class Bootstrap {
 public:
  // Creates a rendezvous root (listening socket + relay thread) in this process.
  // `ip`: address the relay listens on (default: UCCL_B200_BOOTSTRAP_IP, else 127.0.0.1)
  static UniqueId create_id(const char* ip = nullptr);
  Bootstrap(const UniqueId& id, int rank, int nranks);
  ~Bootstrap();
  Bootstrap(const Bootstrap&) = delete;

  int rank() const { return rank_; }
  int nranks() const { return nranks_; }
  uint64_t nonce() const { return nonce_; }

  // out must hold nranks*bytes
  void allgather(const void* in, void* out, size_t bytes);
  void barrier();
  void broadcast(void* buf, size_t bytes, int root);

  // Every rank offers `fds` (same count on every rank); returns result[peer][i] =
  // a local duplicate of peer's fds[i] (result[rank][i] = dup of own).
  std::vector<std::vector<int>> exchange_fds(const std::vector<int>& fds);

 private:
  void send_all(const void* p, size_t n);
  void recv_all(void* p, size_t n);
  int rank_, nranks_;
  uint64_t nonce_;
  int sock_ = -1;       // connection to the relay
  int uds_listen_ = -1; // abstract unix socket for fd passing
  uint32_t fd_round_ = 0;
};

}  // namespace ub
