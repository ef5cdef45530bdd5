// Device code shared by the register-path (ep_kernels.cu) and the TMA-pipelined (ep_tma_kernels.cu)
// expert-parallel kernels: bf16 <-> fp32 / e4m3 helpers, the per-128-channel quantiser and the
// count-exchange prologue of a non-cached dispatch.
#pragma once
#include "../kernels/prims.cuh"
#include "ep_types.h"

namespace ub {

__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 p = __bfloat1622float2(h[i]);
    f[2 * i] = p.x;
    f[2 * i + 1] = p.y;
  }
}

__device__ __forceinline__ uint32_t pack4_e4m3(float a, float b, float c, float d) {
  __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return (uint32_t)lo | ((uint32_t)hi << 16);
}


// scale / inverse scale of one 128-channel group from its amax (bit-identical to the torch reference
// `amax.clamp(1e-4) / 448`; round_scale: power-of-two inverse scale, UE8M0 compatible)
__device__ __forceinline__ void fp8_group_scale(float amax, int round_scale, float& scale, float& scale_inv) {
  amax = fmaxf(amax, 1e-4f);
  if (round_scale) {
    const float raw = amax * (1.0f / 448.0f);
    int ex = ((__float_as_int(raw) >> 23) & 0xff) - 127;
    if ((__float_as_int(raw) & 0x7fffff) != 0) ex += 1;
    scale_inv = __int_as_float((ex + 127) << 23);
    scale = __int_as_float((127 - ex) << 23);
  } else {
    scale = 448.0f / amax;
    scale_inv = __fdiv_rn(amax, 448.0f);
  }
}

// Count exchange of a non-cached dispatch (all threads of the block call it; the block has at
// least kMaxRanks * kMaxRanks threads).  On return s_base[r] = first slot of my tokens inside rank
// r's arena, *s_abort != 0 if some arena would overflow (same decision on every rank) and
// *s_recv_total = tokens this rank receives.  Block 0 also publishes the receive counts.
// Cached dispatches only rendezvous (the peers' arenas may be overwritten from here on).
__device__ __forceinline__ void ep_dispatch_prologue(const DevComm& c, const EpDispatchArgs& a, BlockSync& s,
                                                     int (*s_M)[kMaxRanks], int* s_base, int* s_abort,
                                                     int* s_recv_total) {
  const int R = c.nranks, me = c.rank;
  const int tid = threadIdx.x;
  const int E_local = a.E / R;
  if (!a.cached) {
    // ---- phase 0: all-gather the R x R count matrix (each block keeps a private copy so that
    //      only same-index blocks of different ranks need to synchronise)
    if (tid < R * R) {
      const int dst = tid / R, j = tid % R;
      int* p = reinterpret_cast<int*>(c.heap[dst] + a.cnt_tab_off) + ((size_t)blockIdx.x * kMaxRanks + me) * kMaxRanks + j;
      *p = a.tokens_per_rank[j];
    }
    if (blockIdx.x == 0) {
      for (int i = tid; i < R * E_local; i += blockDim.x) {
        const int dst = i / E_local, e = i % E_local;
        int* p = reinterpret_cast<int*>(c.heap[dst] + a.exp_tab_off) + (size_t)me * kEpMaxLocalExperts + e;
        *p = a.tokens_per_expert[dst * E_local + e];
      }
    }
    sync_barrier(c, s);
    const int* my_tab = reinterpret_cast<const int*>(c.heap[me] + a.cnt_tab_off) + (size_t)blockIdx.x * kMaxRanks * kMaxRanks;
    if (tid < R * R) s_M[tid / R][tid % R] = my_tab[(tid / R) * kMaxRanks + (tid % R)];
    __syncthreads();
    if (tid < R) {
      int base = 0, tot = 0;
      for (int q = 0; q < R; ++q) {
        if (q < me) base += s_M[q][tid];
        tot += s_M[q][tid];
      }
      s_base[tid] = base;
      if (tot > a.arena.capacity) *s_abort = 1;  // identical decision on every rank (same matrix)
      if (tid == me) *s_recv_total = tot;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
      if (a.rank_prefix && tid < R * R) a.rank_prefix[tid] = s_M[tid / R][tid % R];
      const int* exp_tab = reinterpret_cast<const int*>(c.heap[me] + a.exp_tab_off);
      for (int e = tid; e < E_local; e += blockDim.x) {
        int tot = 0;
        for (int q = 0; q < R; ++q) tot += exp_tab[(size_t)q * kEpMaxLocalExperts + e];
        const int al = a.expert_alignment > 1 ? a.expert_alignment : 1;
        tot = (tot + al - 1) / al * al;
        a.dev_counts[1 + e] = tot;
        if (a.host_counts) a.host_counts[1 + e] = tot;
      }
      __syncthreads();
      if (tid == 0) {
        const int v = *s_abort ? -2 : *s_recv_total;
        a.dev_counts[0] = v;
        if (a.host_counts) {
          __threadfence_system();
          a.host_counts[0] = v;
          __threadfence_system();
        }
      }
    }
  } else {
    sync_barrier_relaxed(c, s);  // peers have entered this dispatch: their arena may be overwritten
    if (tid < R) s_base[tid] = 0;
    __syncthreads();
  }

}

}  // namespace ub
