// LogFMT-10 "simulated cast" arithmetic shared by the CUDA kernel (ep_ll_kernels.cu) and a host entry point that the
// CPU tests compare against the PyTorch definition (uccl_b200.ep.utils.logfmt10_simulate).  Reference behaviour:
// ep/src/internode_ll.cu:934-995.  Plain C++ so that g++ and nvcc compile the same lines.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define UB_LF_HD __host__ __device__ __forceinline__
#else
#define UB_LF_HD inline
#endif

namespace ub {

struct LogFmtParams {
  float lmin, step, step_inv, rounding;
  bool use;  // false: the group passes through unchanged
};

// amax: largest magnitude of the 128-channel group; lmax / lmin: log2 of the largest / smallest NON-ZERO magnitude
// (lmin = +inf when the group is all zero)
UB_LF_HD LogFmtParams logfmt10_params(float amax, float lmax, float lmin) {
  LogFmtParams p;
  p.lmin = fmaxf(lmin, lmax - 32.f);  // range clipped to 2^-32 of the maximum
  p.use = amax <= 1.f && p.lmin < lmax;
  p.step = (lmax - p.lmin) / 510.f;   // 2^9 - 2 intervals
  p.step_inv = 1.f / p.step;
  p.rounding = 2.f - log2f((1.f + exp2f(p.step)) * 0.5f) * p.step_inv;
  return p;
}

UB_LF_HD uint16_t f32_to_bf16_rn_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
  return (uint16_t)(u >> 16);
}

UB_LF_HD float bf16_bits_to_f32(uint16_t b) {
  const uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// la = log2(|value|) (-inf for 0); returns the bf16 bits of the value snapped to the group's grid, sign kept
UB_LF_HD uint16_t logfmt10_quantize_bits(uint16_t bits, float la, const LogFmtParams& p) {
  const float enc = floorf((la - p.lmin) * p.step_inv + p.rounding);
  const float dec = exp2f((enc - 1.f) * p.step + p.lmin);
  return (uint16_t)((bits & 0x8000u) | f32_to_bf16_rn_bits(dec));
}

// host reference over [rows, H] bf16 (H % 128 == 0), in place
inline void logfmt10_host(uint16_t* x, size_t rows, size_t H) {
  for (size_t r = 0; r < rows; ++r)
    for (size_t g = 0; g < H / 128; ++g) {
      uint16_t* v = x + r * H + g * 128;
      float la[128], amax = 0.f, lmax = -INFINITY, lmin = INFINITY;
      for (int i = 0; i < 128; ++i) {
        const float a = fabsf(bf16_bits_to_f32(v[i]));
        la[i] = log2f(a);
        amax = fmaxf(amax, a);
        lmax = fmaxf(lmax, la[i]);
        if (a != 0.f) lmin = fminf(lmin, la[i]);
      }
      const LogFmtParams p = logfmt10_params(amax, lmax, lmin);
      if (!p.use) continue;
      for (int i = 0; i < 128; ++i) v[i] = logfmt10_quantize_bits(v[i], la[i], p);
    }
}

}  // namespace ub
