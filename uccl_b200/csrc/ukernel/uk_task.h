// ukernel: launch-free execution of copy / reduce / signal / wait tasks by a persistent
// worker kernel that polls CPU->device FIFOs.
//
// Role parity: the reference's experimental/ukernel device layer (persistent_kernel_ops.cu:180,262
// `singlePersistentKernel/multiPersistentKernel`, task.h, fifo/*) -- re-designed for one NVSwitch
// node: there is no transport adapter zoo because every peer buffer is load/store reachable
// through the symmetric heap, so "TransportSend" lowers to COPY-to-peer + SIGNAL and
// "TransportRecv" to WAIT, all executed by the same worker CTA.  The FIFO is a host-pinned,
// device-mapped ring (no GDRCopy needed: the device polls over C2C/PCIe, the CPU writes with
// plain stores + a release of the sequence word).
#pragma once
#include <stdint.h>

namespace ub {

enum UkOp : uint32_t {
  UK_NOP = 0,
  UK_COPY = 1,    // dst[0:bytes] = src
  UK_REDUCE = 2,  // dst = src (op) src2, element type `dtype`
  UK_SIGNAL = 3,  // release-add `sig_val` to the 64-bit counter at sig_addr (usually a peer's)
  UK_WAIT = 4,    // acquire-spin until the 64-bit counter at sig_addr >= sig_val
  UK_EXIT = 5,    // worker CTA returns
};

struct alignas(64) UkTask {
  uint32_t op;
  uint32_t dtype;  // ub::DType
  uint32_t redop;  // ub::RedOp (kSum/kProd/kMax/kMin)
  uint32_t flags;
  union {
    uint64_t dst;       // COPY / REDUCE
    uint64_t sig_addr;  // SIGNAL / WAIT: address of the 64-bit counter
  };
  uint64_t src;
  uint64_t src2;
  union {
    uint64_t bytes;    // COPY / REDUCE
    uint64_t sig_val;  // SIGNAL: increment; WAIT: threshold
  };
  uint64_t reserved;
  uint64_t seq;  // written last: index of this task + 1
};
static_assert(sizeof(UkTask) == 64, "UkTask is one cache line");

constexpr int kUkMaxLanes = 32;
constexpr int kUkRingEntries = 1024;  // per lane, power of two

// One FIFO per worker CTA ("lane").  `ring` and `done_host` are device aliases of pinned host
// memory; `done_dev` is device memory (target of cuStreamWaitValue64 on user streams).
struct UkLane {
  const UkTask* ring;
  uint64_t* done_host;
  uint64_t* done_dev;
  uint64_t start;  // first sequence number this launch consumes
};

struct UkWorkerArgs {
  UkLane lane[kUkMaxLanes];
  int nlanes;
  uint32_t* err;        // host-mapped error word (0 = ok)
  uint64_t timeout_ns;  // WAIT timeout (0 = never)
  // Semi-persistence: a lane that has seen no task for `idle_ns` votes to quit; when every lane has
  // voted the kernel exits as a whole (at task boundaries) and the host relaunches it on demand.
  // This bounds how long cudaFree / cudaDeviceSynchronize / lazy module loads can be stalled by
  // an idle worker.  0 = stay resident until UK_EXIT.
  uint64_t idle_ns;
  uint32_t* votes;      // device word, zeroed before every launch: low 16 bits = voters, bit 31 = exit
};
constexpr uint32_t kUkExitBit = 0x80000000u;

}  // namespace ub
