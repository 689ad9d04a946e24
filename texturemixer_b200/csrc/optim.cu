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

__global__ void adam_powers_kernel(float* powers, float b1, float b2, const int* flag) {
  if (flag != nullptr && *flag != 0) return;
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

extern "C" int tmx_ema_update(tmx_handle_t h, const float* src, float* dst, int64_t n, float beta, tmx_stream_t s) {
  TMX_REQUIRE(h && src && dst && n >= 0, TMX_ERR_ARG, "tmx_ema_update: bad argument");
  if (n == 0) return TMX_OK;
  ema_kernel<<<grid_for(h, n), 256, 0, (cudaStream_t)s>>>(src, dst, n, beta);
  TMX_LAUNCHED(h, "ema_kernel");
  return TMX_OK;
}
