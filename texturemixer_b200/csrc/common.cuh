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

// Programmatic dependent launch (PDL), OPT-IN with TMX_PDL=1: a kernel launched with tmx_launch_pdl may start while its
// predecessor on the stream is still running - its CTAs take the SMs the predecessor's CTAs leave, set up their
// barriers / TMEM / index arithmetic, and block in tmx_pdl_wait() until the predecessor grid has completed and its
// writes are visible.  Every kernel launched that way calls tmx_pdl_trigger() first (lets ITS successor be scheduled
// early) and tmx_pdl_wait() before its first global-memory access; both are no-ops in an ordinary launch.
// Measured on the train step (conv_tc / conv_lin / conv_wgrad / wgrad_reduce / grad_prepare, ~1 700 of the 2 700
// launches of a step, inside the CUDA graph; all 211 GPU tests pass with it on): 73.9 ms with PDL vs 72.9 ms without
// (profiles/r02_pdl_ab.txt) - the step runs under the board's power cap (SM clocks ~1 780 of 1 965 MHz), so closing
// the ~1 us gaps between kernels buys no time.  Hence off by default.
__device__ __forceinline__ void tmx_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tmx_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// cudaLaunchKernelEx with an optional cluster width and, with TMX_PDL=1, the programmatic-stream-serialization
// attribute.  Works under stream capture: the edge to the previous kernel node becomes a programmatic dependency.
template <typename... KArgs, typename... Args>
static inline cudaError_t tmx_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         int cluster_x, Args... args) {
  static const bool pdl = tmx_env_flag("TMX_PDL");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

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
