// optim.cu — the optimizer step of the reference's tfutil.Optimizer (tfutil.py:246-399) and the EMA of
// Network.setup_as_moving_average_of (tfutil.py:611-621) as fused multi-tensor kernels over the flat fp32
// variable buffer of a network (HBM-bound: 16 B read + 8 B written per parameter for Adam).
//
//   grads arrive summed over ranks (one NCCL all-reduce on the flat buffer, tfutil.py:326-333);
//   g <- g * grad_scale (1/num_gpus, tfutil.py:340-344);
//   if any g is non-finite the whole update is skipped (tfutil.py:347-355);
//   TF1 Adam (tf.train.AdamOptimizer): m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;
//     lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  w <- w - lr_t m / (sqrt(v) + eps)   (epsilon OUTSIDE the sqrt,
//     bias correction folded into lr_t); b1^t, b2^t live on the device and advance only on applied steps.
#include "common.cuh"

__global__ void __launch_bounds__(256) nonfinite_kernel(const float* __restrict__ g, long long n, int* flag) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = __ldg(g + i);
    bad |= !(fabsf(v) <= 3.402823466e38f);   // inf or nan
  }
  if (bad) *flag = 1;
}

// float mark in a gradient bucket's tail: every rank marks its LOCAL non-finite gradients before the SUM all-reduce,
// so that afterwards mark > 0 on all ranks iff any rank overflowed (the skip decision must be global)
__global__ void __launch_bounds__(256) nonfinite_mark_kernel(const float* __restrict__ g, long long n, float* mark) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = __ldg(g + i);
    bad |= !(fabsf(v) <= 3.402823466e38f);
  }
  if (bad) *mark = 1.f;
}

__device__ __forceinline__ bool adam_skipped(const int* flag, const float* mark) {
  // (a NaN mark - a rank whose mark slot itself was poisoned - also skips)
  return (flag != nullptr && *flag != 0) || (mark != nullptr && !(*mark == 0.f));
}

// 128-bit variant over the 256-B aligned flat buffers (n4 = n / 4 float4 elements): the step is pure HBM traffic
__global__ void __launch_bounds__(256) adam_kernel_v4(float4* __restrict__ w, const float4* __restrict__ g,
                                                      float4* __restrict__ m, float4* __restrict__ v, long long n4,
                                                      float lr, float b1, float b2, float eps, float grad_scale,
                                                      const float* __restrict__ powers, const int* __restrict__ flag,
                                                      const float* __restrict__ mark) {
  if (adam_skipped(flag, mark)) return;
  const float p1 = powers[0], p2 = powers[1];
  const float lr_t = lr * sqrtf(1.f - p2) / (1.f - p1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g4 = __ldg(g + i);
    float4 w4 = w[i], m4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = v[i];
    if (b1 != 0.f) m4 = m[i];                       // beta1 = 0 (config.py:84-87): m == g, no need to read it
    const float gi[4] = {g4.x * grad_scale, g4.y * grad_scale, g4.z * grad_scale, g4.w * grad_scale};
    float wi[4] = {w4.x, w4.y, w4.z, w4.w}, mi[4] = {m4.x, m4.y, m4.z, m4.w}, vi[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mi[j] = b1 * mi[j] + (1.f - b1) * gi[j];
      vi[j] = b2 * vi[j] + (1.f - b2) * gi[j] * gi[j];
      wi[j] = wi[j] - lr_t * mi[j] / (sqrtf(vi[j]) + eps);
    }
    m[i] = make_float4(mi[0], mi[1], mi[2], mi[3]);
    v[i] = make_float4(vi[0], vi[1], vi[2], vi[3]);
    w[i] = make_float4(wi[0], wi[1], wi[2], wi[3]);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                   float b1, float b2, float eps, float grad_scale,
                                                   const float* __restrict__ powers, const int* __restrict__ flag) {
  if (flag != nullptr && *flag != 0) return;   // skipped step: nothing moves (tfutil.py:355 tf.cond -> no_op)
  const float p1 = powers[0], p2 = powers[1];
  const float lr_t = lr * sqrtf(1.f - p2) / (1.f - p1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = __ldg(g + i) * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    w[i] = w[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void adam_powers_kernel(float* powers, float b1, float b2, const int* flag, const float* mark = nullptr) {
  if (adam_skipped(flag, mark)) return;
  powers[0] *= b1;
  powers[1] *= b2;
}

__global__ void __launch_bounds__(256) ema_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n,
                                                  float beta) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float s = __ldg(src + i);
    dst[i] = s + (dst[i] - s) * beta;   // tfutil.lerp(src, cur, beta), tfutil.py:41-43, 617
  }
}

static int grid_for(tmx_handle_t h, long long n) {
  long long blocks = (n + 255) / 256;
  long long cap = (long long)h->sm_count * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

extern "C" int tmx_nonfinite_check(tmx_handle_t h, const float* g, int64_t n, int* flag, tmx_stream_t s) {
  TMX_REQUIRE(h && g && flag && n >= 0, TMX_ERR_ARG, "tmx_nonfinite_check: bad argument");
  if (n == 0) return TMX_OK;
  nonfinite_kernel<<<grid_for(h, n), 256, 0, (cudaStream_t)s>>>(g, n, flag);
  TMX_LAUNCHED(h, "nonfinite_kernel");
  return TMX_OK;
}

extern "C" int tmx_adam_step(tmx_handle_t h, float* w, const float* g, float* m, float* v, int64_t n, float lr,
                             float beta1, float beta2, float eps, float grad_scale, float* powers, const int* skip_flag,
                             tmx_stream_t s) {
  TMX_REQUIRE(h && w && g && m && v && powers && n >= 0, TMX_ERR_ARG, "tmx_adam_step: bad argument");
  if (n == 0) return TMX_OK;
  adam_kernel<<<grid_for(h, n), 256, 0, (cudaStream_t)s>>>(w, g, m, v, n, lr, beta1, beta2, eps, grad_scale, powers,
                                                          skip_flag);
  TMX_LAUNCHED(h, "adam_kernel");
  adam_powers_kernel<<<1, 1, 0, (cudaStream_t)s>>>(powers, beta1, beta2, skip_flag);
  TMX_LAUNCHED(h, "adam_powers_kernel");
  return TMX_OK;
}

extern "C" int tmx_nonfinite_mark(tmx_handle_t h, const float* g, int64_t n, float* mark, tmx_stream_t s) {
  TMX_REQUIRE(h && g && mark && n >= 0, TMX_ERR_ARG, "tmx_nonfinite_mark: bad argument");
  if (n == 0) return TMX_OK;
  nonfinite_mark_kernel<<<grid_for(h, n), 256, 0, (cudaStream_t)s>>>(g, n, mark);
  TMX_LAUNCHED(h, "nonfinite_mark_kernel");
  return TMX_OK;
}

extern "C" int tmx_adam_update(tmx_handle_t h, float* w, const float* g, float* m, float* v, int64_t n, float lr,
                               float beta1, float beta2, float eps, float grad_scale, const float* powers,
                               const int* skip_flag, const float* skip_mark, tmx_stream_t s) {
  TMX_REQUIRE(h && w && g && m && v && powers && n >= 0, TMX_ERR_ARG, "tmx_adam_update: bad argument");
  if (n == 0) return TMX_OK;
  const bool aligned = n % 4 == 0 && ((uintptr_t)w | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0;
  if (aligned) {
    adam_kernel_v4<<<grid_for(h, n / 4), 256, 0, (cudaStream_t)s>>>(
        reinterpret_cast<float4*>(w), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), n / 4, lr, beta1, beta2, eps, grad_scale, powers, skip_flag, skip_mark);
    TMX_LAUNCHED(h, "adam_kernel_v4");
  } else {
    TMX_REQUIRE(skip_mark == nullptr, TMX_ERR_ARG, "tmx_adam_update: skip_mark needs 16-B aligned buffers, n %% 4 == 0");
    adam_kernel<<<grid_for(h, n), 256, 0, (cudaStream_t)s>>>(w, g, m, v, n, lr, beta1, beta2, eps, grad_scale, powers,
                                                            skip_flag);
    TMX_LAUNCHED(h, "adam_kernel");
  }
  return TMX_OK;
}

extern "C" int tmx_adam_advance(tmx_handle_t h, float* powers, float beta1, float beta2, const int* skip_flag,
                                const float* skip_mark, tmx_stream_t s) {
  TMX_REQUIRE(h && powers, TMX_ERR_ARG, "tmx_adam_advance: bad argument");
  adam_powers_kernel<<<1, 1, 0, (cudaStream_t)s>>>(powers, beta1, beta2, skip_flag, skip_mark);
  TMX_LAUNCHED(h, "adam_powers_kernel");
  return TMX_OK;
}

extern "C" int tmx_ema_update(tmx_handle_t h, const float* src, float* dst, int64_t n, float beta, tmx_stream_t s) {
  TMX_REQUIRE(h && src && dst && n >= 0, TMX_ERR_ARG, "tmx_ema_update: bad argument");
  if (n == 0) return TMX_OK;
  ema_kernel<<<grid_for(h, n), 256, 0, (cudaStream_t)s>>>(src, dst, n, beta);
  TMX_LAUNCHED(h, "ema_kernel");
  return TMX_OK;
}
