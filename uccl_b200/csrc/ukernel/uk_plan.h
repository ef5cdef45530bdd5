// ukernel CCL planner: turns a collective into a per-rank DAG of tile operations
//   Send (copy into a peer's buffer + signal)   Recv (wait for a peer's signal)
//   Copy / Reduce (local)
// that a backend executes (device worker FIFOs, host threads, or a recording mock).
//
// Reference role: experimental/ukernel/src/ccl/plan.{h,cc} (`plan.h:13-25`: ring AllReduce and
// AllToAll tile DAGs lowered to TransportSend/Recv + DeviceCopy/Reduce) and selector.cc:9-52.
// On NVSwitch the planner additionally knows a full-mesh two-shot AllReduce (2 communication
// steps instead of 2(N-1)) and picks it by default; the ring is kept for parity and for testing
// the executor against a second, structurally different DAG.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace ub {

enum class UkColl : int { AllReduce = 0, AllToAll = 1, AllGather = 2, Barrier = 3, ReduceScatter = 4, Broadcast = 5 };
enum class UkAlgo : int { Auto = 0, Ring = 1, FullMesh = 2 };
enum class UkBuf : int { In = 0, Out = 1, Scratch = 2 };

struct UkRef {
  UkBuf buf = UkBuf::In;
  uint64_t off = 0;
};

struct UkPlanOp {
  enum Kind : int { Copy = 0, Reduce = 1, Send = 2, Recv = 3 };
  int kind = Copy;
  int lane = 0;
  int tile = 0;
  int step = 0;
  int peer = -1;        // Send: destination rank; Recv: source rank
  UkRef dst, src, src2; // Send: dst is in `peer`'s buffer, src in mine.  Reduce: dst = src (op) src2
  uint64_t bytes = 0;   // Send with 0 bytes = pure signal
  std::vector<int> deps;  // indices (into UkPlan::ops) this op must follow
};

struct UkPlan {
  UkColl coll = UkColl::AllReduce;
  UkAlgo algo = UkAlgo::FullMesh;
  int nranks = 1, rank = 0, nlanes = 1;
  uint64_t bytes = 0;          // AllReduce/Broadcast: message bytes; AllToAll/AllGather/ReduceScatter: bytes per peer block
  int root = 0;                // Broadcast
  uint64_t tile_bytes = 0;
  uint64_t scratch_bytes = 0;  // scratch this plan addresses
  std::vector<UkPlanOp> ops;   // emission order == per-lane program order
  std::string describe() const;
};

struct UkPlanParams {
  int nranks = 1, rank = 0, nlanes = 1;
  uint64_t tile_bytes = 1 << 20;  // multiple of 16
  uint64_t elem_size = 1;
  UkAlgo algo = UkAlgo::Auto;
};

UkAlgo uk_select_algo(UkColl coll, int nranks, uint64_t bytes);
uint64_t uk_scratch_bytes(UkAlgo algo, int nranks, int nlanes, uint64_t tile_bytes);

UkPlan uk_plan_allreduce(uint64_t bytes, const UkPlanParams& p);
UkPlan uk_plan_alltoall(uint64_t bytes_per_peer, const UkPlanParams& p);
UkPlan uk_plan_allgather(uint64_t bytes_per_rank, const UkPlanParams& p);
UkPlan uk_plan_barrier(const UkPlanParams& p);
// In holds nranks pieces of `bytes_per_rank`; Out receives the reduction of piece `rank`
UkPlan uk_plan_reduce_scatter(uint64_t bytes_per_rank, const UkPlanParams& p);
// `bytes` of the root's In end up in every rank's Out
UkPlan uk_plan_broadcast(uint64_t bytes, int root, const UkPlanParams& p);

// Structural checks: deps acyclic and backwards, every Send of rank a to rank b on lane l is matched by
// a Recv of b from a on lane l with the same ordinal and size.  `plans` holds one plan per rank.
// Returns "" if valid, otherwise a description of the first violation.
std::string uk_validate(const std::vector<UkPlan>& plans);

// User-authored plans (uccl_b200.ukernel.dsl -- the role of the reference's MSCCL++-DSL JSON plans and their
// interpreter, experimental/lite/collective/execution_kernel.hpp:898): checks one rank's plan against the buffers it
// is going to run on.  Every reference must stay inside its buffer (In / Out sizes are the same on every rank,
// Scratch is bounded by `scratch_cap`), offsets must be 16-byte aligned (the worker's copy units), Reduce sizes a
// multiple of the element size, lanes < max_lanes.  Returns "" if the plan may be executed.
std::string uk_check_bounds(const UkPlan& plan, uint64_t in_bytes, uint64_t out_bytes, uint64_t scratch_cap,
                            int max_lanes, uint64_t elem_size);

// Reference executor over plain host memory (all ranks in one address space): runs the plans of
// all ranks to completion with a round-robin scheduler that only fires an op when its deps and its
// matching Send have fired.  Detects deadlock.  Used by the unit tests as the "ring-allreduce
// simulator" (reference: experimental/ukernel/src/ccl/test/unit/test_components.cc:240-635).
struct UkSimBuffers {
  std::vector<char*> in, out, scratch;  // per rank
};
std::string uk_simulate(const std::vector<UkPlan>& plans, const UkSimBuffers& b, int dtype, int redop);

}  // namespace ub
