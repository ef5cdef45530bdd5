// dtype / op tables (numbering matches ncclDataType_t / ncclRedOp_t so the NCCL
// shim can pass values straight through) and 16-byte vector reduce functors.
// Accumulation for 16-bit and 8-bit floats is done in fp32 across all ranks and
// rounded once (the reference's lite kernels add in the storage type pairwise:
// experimental/lite/collective/common.hpp:62-67).
#pragma once
#include "prims.cuh"

namespace ub {

template <typename T> struct AccOf { using type = T; };
template <> struct AccOf<__half> { using type = float; };
template <> struct AccOf<__nv_bfloat16> { using type = float; };
template <> struct AccOf<__nv_fp8_e4m3> { using type = float; };
template <> struct AccOf<__nv_fp8_e5m2> { using type = float; };

template <typename A, typename T> __device__ __forceinline__ A to_acc(T v) { return (A)v; }
template <> __device__ __forceinline__ float to_acc<float, __half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_acc<float, __nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_acc<float, __nv_fp8_e4m3>(__nv_fp8_e4m3 v) { return (float)v; }
template <> __device__ __forceinline__ float to_acc<float, __nv_fp8_e5m2>(__nv_fp8_e5m2 v) { return (float)v; }

template <typename T, typename A> __device__ __forceinline__ T from_acc(A v) { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __nv_fp8_e4m3 from_acc<__nv_fp8_e4m3, float>(float v) { return __nv_fp8_e4m3(v); }
template <> __device__ __forceinline__ __nv_fp8_e5m2 from_acc<__nv_fp8_e5m2, float>(float v) { return __nv_fp8_e5m2(v); }

template <int OP, typename A> __device__ __forceinline__ A red_apply(A a, A b) {
  if constexpr (OP == kSum || OP == kAvg) return a + b;
  else if constexpr (OP == kProd) return a * b;
  else if constexpr (OP == kMax) return a > b ? a : b;
  else return a < b ? a : b;
}

template <typename A> struct IsFloatAcc { static constexpr bool value = false; };
template <> struct IsFloatAcc<float> { static constexpr bool value = true; };
template <> struct IsFloatAcc<double> { static constexpr bool value = true; };

// A 16-byte word seen as N elements of T with accumulators of AccOf<T>.
template <typename T, int OP>
struct Vec16 {
  using A = typename AccOf<T>::type;
  static constexpr int N = 16 / (int)sizeof(T);
  A a[N];

  __device__ __forceinline__ void init(const uint4& v) {
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = to_acc<A, T>(e[i]);
  }
  __device__ __forceinline__ void accum(const uint4& v) {
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = red_apply<OP, A>(a[i], to_acc<A, T>(e[i]));
  }
  __device__ __forceinline__ void epilogue(const Epilogue& ep) {
    if constexpr (IsFloatAcc<A>::value) {
      if (ep.scale != 1.0f) {
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = (A)(a[i] * (A)ep.scale);
      }
    } else {
      if (ep.idiv != 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = (A)(a[i] / (A)ep.idiv);
      }
    }
  }
  // pack to N elements of TO (N * sizeof(TO) bytes; up to 32)
  template <typename TO>
  struct Out {
    TO e[N];
  };
  template <typename TO>
  __device__ __forceinline__ void pack(Out<TO>& o) const {
#pragma unroll
    for (int i = 0; i < N; ++i) o.e[i] = from_acc<TO, A>(a[i]);
  }
  __device__ __forceinline__ uint4 pack_same() const {
    uint4 v;
    T* e = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = from_acc<T, A>(a[i]);
    return v;
  }
};

// Store N elements of TO produced from one 16-byte input word, at element index `elem`.
template <typename T, int OP, typename TO, bool MC = false>
__device__ __forceinline__ void store_out(TO* base, size_t elem, const Vec16<T, OP>& v) {
  constexpr int N = Vec16<T, OP>::N;
  constexpr int BYTES = N * (int)sizeof(TO);
  alignas(16) typename Vec16<T, OP>::template Out<TO> o;
  v.pack(o);
  char* dst = reinterpret_cast<char*>(base + elem);
  if constexpr (BYTES == 16) {
    if constexpr (MC) multimem_st_v4(dst, *reinterpret_cast<uint4*>(&o));
    else st_v4(dst, *reinterpret_cast<uint4*>(&o));
  } else if constexpr (BYTES == 32) {
    if constexpr (MC) {
      multimem_st_v4(dst, reinterpret_cast<uint4*>(&o)[0]);
      multimem_st_v4(dst + 16, reinterpret_cast<uint4*>(&o)[1]);
    } else {
      st_v4(dst, reinterpret_cast<uint4*>(&o)[0]);
      st_v4(dst + 16, reinterpret_cast<uint4*>(&o)[1]);
    }
  } else if constexpr (BYTES == 8) {
    if constexpr (MC) multimem_st_v2(dst, *reinterpret_cast<uint2*>(&o));
    else st_v2(dst, *reinterpret_cast<uint2*>(&o));
  } else {
    static_assert(BYTES == 64 || BYTES == 4 || BYTES == 2, "unsupported output width");
    if constexpr (BYTES == 64) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (MC) multimem_st_v4(dst + 16 * k, reinterpret_cast<uint4*>(&o)[k]);
        else st_v4(dst + 16 * k, reinterpret_cast<uint4*>(&o)[k]);
      }
    } else if constexpr (BYTES == 4) {
      *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<uint32_t*>(&o);
    } else {
      *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<uint16_t*>(&o);
    }
  }
}

// NVLS in-switch reduction of one 16-byte word (sum only for f32; sum/min/max for 16-bit).
template <typename T, int OP> struct MmLdRed { static constexpr bool ok = false; __device__ static uint4 ld(const void*) { return uint4(); } };
template <> struct MmLdRed<float, kSum> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_add_f32(p); } };
template <> struct MmLdRed<float, kAvg> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_add_f32(p); } };
template <> struct MmLdRed<__nv_bfloat16, kSum> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_add_bf16(p); } };
template <> struct MmLdRed<__nv_bfloat16, kAvg> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_add_bf16(p); } };
template <> struct MmLdRed<__nv_bfloat16, kMax> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_max_bf16(p); } };
template <> struct MmLdRed<__nv_bfloat16, kMin> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_min_bf16(p); } };
template <> struct MmLdRed<__half, kSum> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_add_f16(p); } };
template <> struct MmLdRed<__half, kAvg> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_add_f16(p); } };
template <> struct MmLdRed<__half, kMax> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_max_f16(p); } };
template <> struct MmLdRed<__half, kMin> { static constexpr bool ok = true; __device__ static __forceinline__ uint4 ld(const void* p) { return mm_ldred_min_f16(p); } };

}  // namespace ub
