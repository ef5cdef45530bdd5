// Barrier-free small-message AllGather / AllToAll / ReduceScatter ("LL exchange").
//
// Same packet protocol as ar_oneshot (allreduce_impl.cuh): every rank stores LL16 packets
// {data,flag,data,flag} straight into slot[src_rank] of the *destination's* scratch and the
// destination polls the flags while unpacking -- no entry barrier, no exit barrier, works on
// arbitrary (non-symmetric) user buffers, CUDA-graph safe (the flag is a device-side epoch).
// The parity double-buffer is safe because in these three collectives every rank receives from
// every peer: nobody can be two calls ahead of a peer that has not finished reading.
// (Broadcast / Reduce do not have that property and keep their barrier-based kernels.)
//
// The reference has no counterpart: its lite layer only has LL packets for AllReduce
// (experimental/lite/collective/allreduce_packet.cu) and no native ReduceScatter / AllToAll
// at all (nccl.cu:1952-1966, 2069-2102).
#pragma once
#include "coll_common.cuh"
#include "launch.h"

namespace ub {

struct LLCtx {
  uint32_t* misc;
  uint32_t flag;
  uint64_t parity_off;
};

__device__ __forceinline__ LLCtx ll_begin(const DevComm& c, const CollArgs& a) {
  __shared__ uint32_t s_ll_flag;
  LLCtx l;
  l.misc = reinterpret_cast<uint32_t*>(c.heap[c.rank] + a.misc_off);
  if (threadIdx.x == 0) s_ll_flag = ld_volatile(l.misc + kLLEpoch) + 1;
  __syncthreads();
  l.flag = s_ll_flag;
  l.parity_off = a.ll_off + (uint64_t)(l.flag & 1u) * (kMaxRanks * kLLSlotBytes);
  return l;
}

// last block to finish bumps the epoch (graph-replay safe: no host-side counter)
__device__ __forceinline__ void ll_end(const LLCtx& l) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    uint32_t old = atomicAdd(l.misc + kLLDone, 1u);
    if (old == gridDim.x - 1) {
      l.misc[kLLDone] = 0;
      l.misc[kLLEpoch] = l.flag;
      __threadfence();
    }
  }
}

__device__ __forceinline__ void ll_put(char* dst, const uint4& d, uint32_t flag) {
  st_v4(dst, make_uint4(d.x, flag, d.y, flag));
  st_v4(dst + 16, make_uint4(d.z, flag, d.w, flag));
}

__device__ __forceinline__ uint4 ll_get(const DevComm& c, const char* slot, uint32_t flag, int src) {
  uint4 p0, p1;
  SpinGuard g(c.timeout_ns);
  while (true) {
    p0 = ld_volatile_v4(slot);
    p1 = ld_volatile_v4(slot + 16);
    if (p0.y == flag && p0.w == flag && p1.y == flag && p1.w == flag) break;
    if (g.expired()) comm_abort(c, 11, src, (int)flag);
  }
  return make_uint4(p0.x, p0.z, p1.x, p1.z);
}

// phase 1 of AllToAll / ReduceScatter: piece p of my input -> slot[rank] of peer p
__device__ __forceinline__ void ll_scatter_pieces(const DevComm& c, const CollArgs& a, const LLCtx& l,
                                                  uint64_t units) {
  const int n = c.nranks, rank = c.rank;
  const char* in = reinterpret_cast<const char*>(a.in);
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t total = (uint64_t)(n - 1) * units;
  constexpr int B = 4;
  for (uint64_t i0 = gtid; i0 < total; i0 += gstride * B) {
    uint4 d[B];
    int pe[B];
    uint64_t uu[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const uint64_t i = i0 + (uint64_t)j * gstride;
      if (i < total) {
        const uint64_t k = i / units;
        uu[j] = i - k * units;
        int p = rank + 1 + (int)k;
        if (p >= n) p -= n;
        pe[j] = p;
        d[j] = ld_v4(in + (uint64_t)p * a.bytes + uu[j] * 16);
      }
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const uint64_t i = i0 + (uint64_t)j * gstride;
      if (i < total) ll_put(c.heap[pe[j]] + l.parity_off + (uint64_t)rank * kLLSlotBytes + uu[j] * 32, d[j], l.flag);
    }
  }
}

// MODE 0 = AllGather (a.bytes per rank; out holds n pieces)   MODE 1 = AllToAll (a.bytes per peer)
// MC: AllGather publishes with one multimem.st through the switch instead of n-1 P2P stores.
// a.bytes % 16 == 0 (host-checked).
template <int MODE, bool MC>
__global__ void __launch_bounds__(512) xchg_ll_kernel(const __grid_constant__ DevComm c,
                                                      const __grid_constant__ CollArgs a) {
  const LLCtx l = ll_begin(c, a);
  const int n = c.nranks, rank = c.rank;
  const uint64_t units = a.bytes / 16;
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
  const char* in = reinterpret_cast<const char*>(a.in);
  char* out = reinterpret_cast<char*>(a.out);
  if constexpr (MODE == 0) {
    constexpr int B = 4;
    for (uint64_t u0 = gtid; u0 < units; u0 += gstride * B) {
      uint4 d[B];
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const uint64_t u = u0 + (uint64_t)j * gstride;
        if (u < units) d[j] = ld_v4(in + u * 16);
      }
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const uint64_t u = u0 + (uint64_t)j * gstride;
        if (u >= units) continue;
        const uint64_t off = l.parity_off + (uint64_t)rank * kLLSlotBytes + u * 32;
        if constexpr (MC) {
          multimem_st_v4(c.mc + off, make_uint4(d[j].x, l.flag, d[j].y, l.flag));
          multimem_st_v4(c.mc + off + 16, make_uint4(d[j].z, l.flag, d[j].w, l.flag));
        } else {
          for (int k = 1; k < n; ++k) {
            int p = rank + k;
            if (p >= n) p -= n;
            ll_put(c.heap[p] + off, d[j], l.flag);
          }
        }
        // my own piece goes straight to its place (no-op when the call is in-place)
        char* own = out + (uint64_t)rank * a.bytes + u * 16;
        if (own != in + u * 16) st_v4(own, d[j]);
      }
    }
  } else {
    ll_scatter_pieces(c, a, l, units);
    const char* own_in = in + (uint64_t)rank * a.bytes;
    char* own_out = out + (uint64_t)rank * a.bytes;
    for (uint64_t u = gtid; u < units; u += gstride) st_v4(own_out + u * 16, ld_v4(own_in + u * 16));
  }
  // phase 2: unpack what the peers sent me
  const char* my_ll = c.heap[rank] + l.parity_off;
  const uint64_t total = (uint64_t)(n - 1) * units;
  for (uint64_t i = gtid; i < total; i += gstride) {
    const uint64_t k = i / units;
    const uint64_t u = i - k * units;
    int s = rank + 1 + (int)k;
    if (s >= n) s -= n;
    const uint4 d = ll_get(c, my_ll + (uint64_t)s * kLLSlotBytes + u * 32, l.flag, s);
    st_v4(out + (uint64_t)s * a.bytes + u * 16, d);
  }
  ll_end(l);
}

// ReduceScatter: a.bytes = bytes each rank receives (a.bytes % 16 == 0); input holds n pieces.
// Reduction in rank order => bitwise identical to the barrier-based kernels' order.
template <typename T, int OP>
__global__ void __launch_bounds__(512) rs_ll_kernel(const __grid_constant__ DevComm c,
                                                    const __grid_constant__ CollArgs a) {
  const LLCtx l = ll_begin(c, a);
  const int n = c.nranks, rank = c.rank;
  const uint64_t units = a.bytes / 16;
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
  const char* in = reinterpret_cast<const char*>(a.in);
  ll_scatter_pieces(c, a, l, units);
  const char* my_ll = c.heap[rank] + l.parity_off;
  const char* own_in = in + (uint64_t)rank * a.bytes;
  for (uint64_t u = gtid; u < units; u += gstride) {
    Vec16<T, OP> acc;
    for (int s = 0; s < n; ++s) {
      uint4 d;
      if (s == rank) d = ld_v4(own_in + u * 16);
      else d = ll_get(c, my_ll + (uint64_t)s * kLLSlotBytes + u * 32, l.flag, s);
      if (s == 0) acc.init(d);
      else acc.accum(d);
    }
    acc.epilogue(a.ep);
    st_v4(reinterpret_cast<char*>(a.out) + u * 16, acc.pack_same());
  }
  ll_end(l);
}

}  // namespace ub
