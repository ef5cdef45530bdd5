// MultiComm: collectives for a rank group that spans several boxes, in C++ -- what the NCCL-API drop-in
// uses when `ncclCommInitRank` is called with more ranks than one NVLink domain holds.
//
// Same rail-aligned hierarchy as `uccl_b200/parallel/multinode.py`: ranks [k*L, (k+1)*L) form box k with a
// native `Comm` (NVLink kernels over the symmetric heap); local rank l of every box forms rail l over the
// multipath datagram transport (`csrc/net`).  all_reduce = NVLink reduce-scatter -> ring all-reduce of the
// shard along the rail (host staged) -> NVLink all-gather; the other collectives follow the same shape.
//
// Reference role: experimental/lite's inter-node path (nccl.cu:337-347,574-700: D2H to pinned staging +
// ibv_post_send) and NCCL's own net transport under the UCCL plugin.  Calls are host-synchronous across the
// network phase (the device work before and after is stream ordered), which is also how the reference's lite
// inter-node send/recv behaves.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "../net/net_engine.h"
#include "comm.h"

namespace ub {

class MultiComm {
 public:
  // ranks [k*local_size, (k+1)*local_size) share a box.  `id` is consumed by the global bootstrap.
  static std::shared_ptr<MultiComm> create(const UniqueId& id, int rank, int nranks, int local_size, int device,
                                           const CommConfig& cfg);
  ~MultiComm();

  int rank() const { return rank_; }
  int nranks() const { return nranks_; }
  int local_rank() const { return lrank_; }
  int local_size() const { return L_; }
  int node() const { return node_; }
  int nnodes() const { return N_; }
  bool is_host() const { return local_->is_host(); }
  int device() const { return local_->device(); }
  const std::shared_ptr<Comm>& local() const { return local_; }
  std::string describe() const;

  void allreduce(const void* in, void* out, size_t count, int dtype, int op, cudaStream_t stream, float scale = 1.0f);
  void allgather(const void* in, void* out, size_t count_per_rank, int dtype, cudaStream_t stream);
  void reduce_scatter(const void* in, void* out, size_t recv_count, int dtype, int op, cudaStream_t stream);
  void broadcast(const void* in, void* out, size_t count, int dtype, int root, cudaStream_t stream);
  void reduce(const void* in, void* out, size_t count, int dtype, int op, int root, cudaStream_t stream);
  void alltoall(const void* in, void* out, size_t count_per_peer, int dtype, cudaStream_t stream);
  // variable splits through the equal-split two-hop path: chunks are padded to the global maximum
  void alltoallv(const void* in, const size_t* send_counts, const size_t* send_displs, void* out, const size_t* recv_counts,
                 const size_t* recv_displs, int dtype, cudaStream_t stream);
  void barrier(cudaStream_t stream);
  // grouped point-to-point: peers inside the box go to the native kernel, rail peers to the transport;
  // peers on another rail of another box are not routed (error)
  void group_p2p(const std::vector<Comm::P2pOp>& ops, cudaStream_t stream);

 private:
  MultiComm() = default;
  // staging: device <-> host for the network phase (identity on the host backend)
  char* host_stage(size_t bytes, int slot);
  void* scratch(size_t bytes, int slot);  // symmetric-heap scratch (collective growth)
  void to_host(void* host, const void* dev, size_t bytes, cudaStream_t s);
  void to_dev(void* dev, const void* host, size_t bytes, cudaStream_t s);
  void copy_dd(void* dst, const void* src, size_t bytes, cudaStream_t s);
  void zero(void* p, size_t bytes, cudaStream_t s);
  void sync(cudaStream_t s);
  // rail primitives on host memory
  void rail_sendrecv(const void* sbuf, size_t sbytes, int to_node, void* rbuf, size_t rbytes, int from_node);
  void rail_allreduce(void* buf, size_t count, int dtype, int op);
  void rail_allgather(void* buf, size_t bytes_per_node);      // buf = [N][bytes], own block filled
  void rail_reduce_scatter(void* buf, size_t count_per_node, int dtype, int op, void* out);
  void rail_broadcast(void* buf, size_t bytes, int root_node);
  void rail_alltoall(const void* in, void* out, size_t bytes_per_node);
  void rail_barrier();
  // all-reduce of `count` elements viewed as [L][per] (zero padded) as a 3-stage pipeline over column blocks
  void allreduce_pipelined(const char* in, char* out, size_t count, size_t per, int dtype, int op, float scale, cudaStream_t st);
  void wait_req(net::Request* r, const char* what);

  int rank_ = 0, nranks_ = 1, L_ = 1, N_ = 1, node_ = 0, lrank_ = 0;
  std::shared_ptr<Comm> local_;
  std::unique_ptr<net::Engine> engine_;
  std::vector<uint32_t> rail_;  // flow to the same local rank of node k (own entry unused)
  int timeout_ms_ = 120000;
  size_t small_bytes_ = 16u << 10;     // messages up to this size take the two-phase latency path (UCCL_B200_MN_SMALL_BYTES)
  size_t scratch_cap_ = 32u << 20;     // upper size of one scratch buffer: large all_gather / reduce_scatter / all_to_all run in chunks
  size_t pipeline_bytes_ = 8u << 20;  // shard bytes above which all-reduce is pipelined (UCCL_B200_MN_PIPELINE_BYTES)
  struct HostBuf {
    char* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
  };
  HostBuf hbuf_[6];
  struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
  };
  DevBuf dbuf_[12];
};

}  // namespace ub
