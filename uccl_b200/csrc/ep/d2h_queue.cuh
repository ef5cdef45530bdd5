// Device side of the GPU->CPU command queue.
#pragma once
#include "../kernels/prims.cuh"
#include "d2h_queue.h"

namespace ub {

// Claims a slot, waits for room, publishes the command.  Any thread may call it.
__device__ __forceinline__ uint64_t d2h_push(const D2HQueueDev& q, uint32_t type, uint32_t dst_rank, uint32_t aux,
                                             uint64_t src_off, uint64_t dst_off, uint32_t bytes, uint32_t value) {
  const uint64_t slot = atomicAdd(q.head, 1ull);
  // flow control: the host publishes how many commands it has consumed; this read crosses PCIe /
  // C2C, so it only happens when the ring might be full
  while (slot - *q.tail >= q.capacity) __nanosleep(200);
  D2HCmd* c = q.ring + (slot & (q.capacity - 1));
  const uint32_t head_word = (type & 0xffu) | ((dst_rank & 0xffu) << 8) | ((aux & 0xffffu) << 16);
  uint4 lo = make_uint4(head_word, value, (uint32_t)src_off, (uint32_t)(src_off >> 32));
  uint4 hi = make_uint4((uint32_t)dst_off, (uint32_t)(dst_off >> 32), bytes, (uint32_t)(slot + 1) | 0x80000000u);
  st_v4(c, lo);
  // the tag sits in the last word of the second store; release ordering keeps the first store and
  // any payload the command refers to (data written by this thread before the call) ahead of it
  fence_acq_rel_sys();
  st_v4(reinterpret_cast<char*>(c) + 16, hi);
  return slot;
}

}  // namespace ub
