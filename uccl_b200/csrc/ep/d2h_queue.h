// GPU -> CPU command queue ("software IBGDA" role of the reference: ep/include/ring_buffer.cuh,
// 128-bit TransferCmd + host-mapped ring polled by CPU proxy threads, ep/src/proxy.cpp:172-1532).
//
// On one NVSwitch node the EP kernels never need it -- peers are load/store reachable -- so here the
// queue is a general device-initiated service channel: a kernel asks the CPU proxy to move data with
// the copy engines (no SM time, "put with signal"), to bump a remote counter after those copies, or
// to deliver a notification to the application.  Multi-producer (any thread of any CTA), single
// consumer per queue.
#pragma once
#include <stdint.h>

namespace ub {

enum D2HCmdType : uint32_t {
  D2H_NOP = 1,     // consumed and acknowledged (latency / throughput microbenchmarks)
  D2H_WRITE = 2,   // heap[self]+src_off -> heap[dst_rank]+dst_off, `bytes` bytes, by copy engine
  D2H_ATOMIC = 3,  // 64-bit add of `value` at heap[dst_rank]+dst_off, ordered after earlier WRITEs
  D2H_NOTIFY = 4,  // (aux, value) delivered to the application (Proxy::poll_notifications)
  D2H_QUIT = 5,
};

struct alignas(32) D2HCmd {
  uint32_t type_dst_aux;  // type | dst_rank << 8 | aux << 16
  uint32_t value;
  uint64_t src_off;
  uint64_t dst_off;
  uint32_t bytes;
  uint32_t tag;  // written last: (slot index + 1) truncated to 32 bits, never 0
};
static_assert(sizeof(D2HCmd) == 32, "D2HCmd is two 16-byte stores");

// Device-side handle (pass by value to kernels).
struct D2HQueueDev {
  D2HCmd* ring;               // host-pinned, device-mapped
  unsigned long long* head;   // device memory: next slot to claim
  const volatile uint64_t* tail;  // host-pinned: number of commands consumed (flow control)
  volatile uint64_t* ack;     // host-pinned: number of NOPs acknowledged
  uint32_t capacity;          // power of two
};

inline uint32_t d2h_pack(uint32_t type, uint32_t dst_rank, uint32_t aux) {
  return (type & 0xffu) | ((dst_rank & 0xffu) << 8) | ((aux & 0xffffu) << 16);
}

}  // namespace ub
