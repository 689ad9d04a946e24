// common.cuh — shared host/device helpers for libtmx (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/tmx.h"

// ---------------------------------------------------------------- context
typedef CUresult (*tmx_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct tmx_ctx {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  int max_smem_optin;
  tmx_encode_tiled_fn encode_tiled;
  unsigned long long launches;
};

// ---------------------------------------------------------------- errors
int tmx_fail(int code, const char* fmt, ...);  // records thread-local message, returns code
int tmx_cuda_fail(cudaError_t e, const char* what);

#define TMX_REQUIRE(cond, code, ...)               \
  do {                                             \
    if (!(cond)) return tmx_fail(code, __VA_ARGS__); \
  } while (0)

#define TMX_CUDA(expr)                                \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return tmx_cuda_fail(_e, #expr); \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define TMX_LAUNCHED(h, name)                                    \
  do {                                                           \
    (h)->launches++;                                             \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return tmx_cuda_fail(_e, name);       \
  } while (0)

static inline int tmx_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
// debugging / A-B switches read from the environment (unset or "0" = off)
static inline bool tmx_env_flag(const char* name) {
  const char* v = getenv(name);
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ int tmx_reflect(int i, int n) {
  // tf.pad(mode='REFLECT') by one pixel: -1 -> 1, n -> n-2 (networks.py:55)
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

// fp32 -> (hi, lo) bf16 pair with x ~= hi + lo; round-to-nearest-even both times (cvt.rn.bf16.f32: one instruction per
// conversion - the integer emulation this replaces made the split-heavy kernels issue-bound: grad_prepare spent 354
// instructions per 8-channel unit at 61 % issue utilisation and 0.59 of the HBM peak).
__device__ __forceinline__ uint32_t tmx_f32_to_bf16_rn(float x) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ void tmx_split_bf16(float x, uint32_t& hi, uint32_t& lo) {
  hi = tmx_f32_to_bf16_rn(x);
  float r = x - __uint_as_float(hi << 16);  // exact in fp32
  lo = tmx_f32_to_bf16_rn(r);
}
// two values at once: packed words (a in the low half, b in the high half) of the hi and lo planes
__device__ __forceinline__ void tmx_split_bf16x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);                     // cvt.rn.bf16x2.f32
  hi2 = *reinterpret_cast<uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi2 << 16);
  const float rb = b - __uint_as_float(hi2 & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo2 = *reinterpret_cast<uint32_t*>(&l);
}

#endif  // __CUDACC__
