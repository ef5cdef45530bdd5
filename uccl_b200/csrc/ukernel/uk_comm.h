// ukernel communicator: plans (uk_plan.h) executed by the persistent worker (uk_worker.h) over
// the symmetric heap of a `Comm`.  Collectives are enqueued without launching anything; user
// streams are ordered against the worker with stream memory operations
// (cuStreamWriteValue64 "inputs ready" -> WAIT task, SIGNAL task -> cuStreamWaitValue64).
//
// Reference role: experimental/ukernel/src/ccl/executor.cc + transport/communicator.cc + the
// torch extension py/ukernel_ccl.cpp (all_reduce / all_to_all_single / barrier).
#pragma once
#include <map>
#include <memory>

#include "../coll/comm.h"
#include "uk_plan.h"
#include "uk_worker.h"

namespace ub {

struct UkCommConfig {
  int nlanes = 4;
  uint64_t tile_bytes = 1 << 20;
  uint64_t staging_bytes = 32 << 20;  // per direction; larger messages are processed in segments
};

class UkComm {
 public:
  UkComm(std::shared_ptr<Comm> comm, const UkCommConfig& cfg);  // collective over the Comm's ranks
  ~UkComm();
  UkComm(const UkComm&) = delete;

  int rank() const { return comm_->rank(); }
  int nranks() const { return comm_->nranks(); }
  int nlanes() const { return cfg_.nlanes; }
  UkWorker& worker() { return *worker_; }

  // Asynchronous collectives; return an operation ticket for wait()/test().  `stream`: the user's
  // stream that produced `in` and will consume `out` (ignored by the host backend).
  // `symmetric`: in/out are heap tensors at the SAME heap offsets on every rank (allocated in the same
  // order everywhere) and may be used in place by the peers; otherwise the data is staged.
  uint64_t all_reduce(const void* in, void* out, size_t count, int dtype, int op, UkAlgo algo, cudaStream_t stream,
                      bool symmetric = false);
  uint64_t all_to_all(const void* in, void* out, size_t count_per_peer, int dtype, cudaStream_t stream,
                      bool symmetric = false);
  uint64_t all_gather(const void* in, void* out, size_t count_per_rank, int dtype, cudaStream_t stream,
                      bool symmetric = false);
  uint64_t reduce_scatter(const void* in, void* out, size_t recv_count, int dtype, int op, cudaStream_t stream);
  uint64_t broadcast(const void* in, void* out, size_t count, int dtype, int root, cudaStream_t stream);
  uint64_t barrier(cudaStream_t stream);
  // Executes a user-authored plan of THIS rank (uk_check_bounds must accept it; every rank calls with its own plan
  // of the same program).  In / Out that are symmetric heap tensors are used in place, anything else is staged
  // through the heap (then both must fit the staging buffers).
  uint64_t run_custom(const UkPlan& plan, const void* in, uint64_t in_bytes, void* out, uint64_t out_bytes, int dtype,
                      int op, cudaStream_t stream, bool symmetric = false);
  uint64_t scratch_capacity() const { return region_bytes_[1]; }
  bool test(uint64_t ticket);
  void wait(uint64_t ticket, double timeout_s = 60.0);
  void stop();

  struct Stats {
    uint64_t ops = 0, segments = 0, tasks = 0, zero_copy_ops = 0, stream_ordered_ops = 0;
  };
  Stats stats() const { return stats_; }

 private:
  struct Bufs {  // resolution of plan buffer names for one segment
    char* in;
    char* out;  // local addresses; both inside the symmetric heap
  };
  void begin_op(cudaStream_t stream);
  uint64_t end_op(cudaStream_t stream);
  void lane_barrier();
  void run_plan(const UkPlan& plan, const Bufs& b, int dtype, int op);
  void push(int lane, const UkTask& t);
  void copy_sliced(char* dst, const char* src, uint64_t bytes);
  char* local(const Bufs& b, const UkRef& r) const;
  char* remote(const char* local_ptr, int peer) const;  // the peer's mapping of the peer's own copy of this buffer
  void resolve_offsets();

  std::shared_ptr<Comm> comm_;
  UkCommConfig cfg_;
  std::unique_ptr<UkWorker> worker_;
  char* ctrl_ = nullptr;     // flags[nlanes][nranks] | lane_sync | ready
  char* scratch_ = nullptr;
  char* stage_in_ = nullptr;
  char* stage_out_ = nullptr;
  uint64_t* flags_ = nullptr;
  uint64_t* lane_sync_ = nullptr;
  uint64_t* ready_ = nullptr;
  std::vector<uint64_t> expected_;  // [lane][src]: signals consumed so far
  uint64_t lane_barriers_ = 0;
  uint64_t ops_ = 0;
  bool stream_ops_ = false;  // cuStreamWrite/WaitValue64 usable
  std::map<uint64_t, std::vector<uint64_t>> tickets_;  // op -> per-lane tickets
  cudaStream_t setup_stream_ = nullptr;
  // heap offsets of {ctrl, scratch, stage_in, stage_out} on every rank: ranks may have allocated at
  // different offsets (fragmented heaps), so they are all-gathered once instead of assumed equal
  enum { kRegions = 4 };
  uint64_t my_offs_[kRegions] = {0, 0, 0, 0};
  uint64_t region_bytes_[kRegions] = {0, 0, 0, 0};
  std::vector<uint64_t> all_offs_;  // [rank][region]
  uint64_t* offs_dev_ = nullptr;    // device staging of my_offs_ / the gathered table
  uint64_t* offs_host_ = nullptr;   // pinned
  bool offs_ready_ = false;
  Stats stats_;
};

}  // namespace ub
