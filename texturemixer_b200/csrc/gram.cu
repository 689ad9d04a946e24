// gram.cu — the VGG-19 Gram-matrix loss of the reference's E/G objective:
//   custom_vgg19.py:31-40   input scaling + BGR mean subtraction               tmx_vgg_preprocess(_bwd)
//   loss.py:29-35           gram_matrix: F^T F / h / w per sample              tmx_gram_fwd
//   loss.py:68-75           multi_layer_diff: mean |G - T| per sample          tmx_gram_l1 (value + its gradient dL/dG)
//   tf.gradients of it      dL/dF = (S + S^T) F / h / w                        tmx_gram_bwd
// The VGG conv stack itself runs on the tensor-core conv kernel (zero-padded, ReLU = leaky slope 0).  The Gram
// products are fp32 CUDA-core GEMMs over the NCHW feature maps [N][C][P = h*w] (rows contiguous along the pixel
// axis = both operands K-major): 0.57 GFLOP per 128x128 image, < 1 % of a train step, and their dynamic range (sums
// of 16 384 products of activations of a few hundred) is what the L1 against the target matrix differences away.
#include "common.cuh"

// ---------------------------------------------------------------- preprocessing
__global__ void __launch_bounds__(256) vgg_pre_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                      long long npix_total, long long hw) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix_total) return;
  const long long n = t / hw, p = t - n * hw;
  const float* s = img + n * 3 * hw + p;
  // custom_vgg19.py:32-40: rgb_scaled = (rgb + 1) / 2 * 255; bgr = [b - 103.939, g - 116.779, r - 123.68]
  const float r = (__ldg(s) + 1.0f) / 2.0f * 255.0f;
  const float g = (__ldg(s + hw) + 1.0f) / 2.0f * 255.0f;
  const float b = (__ldg(s + 2 * hw) + 1.0f) / 2.0f * 255.0f;
  float4* o = reinterpret_cast<float4*>(out + t * 16);       // 16-channel pixel: BGR + 13 zero channels (UMMA K chunk)
  o[0] = make_float4(b - 103.939f, g - 116.779f, r - 123.68f, 0.f);
  o[1] = o[2] = o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) vgg_pre_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dimg,
                                                          long long npix_total, long long hw) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix_total) return;
  const long long n = t / hw, p = t - n * hw;
  const float4 d = __ldg(reinterpret_cast<const float4*>(dout + t * 16));
  float* o = dimg + n * 3 * hw + p;
  o[0] = d.z * 127.5f;          // d/d rgb of (rgb + 1) / 2 * 255
  o[hw] = d.y * 127.5f;
  o[2 * hw] = d.x * 127.5f;
}

extern "C" int tmx_vgg_preprocess(tmx_handle_t h, const float* img_nchw, float* out_nhwc16, int N, int H, int W,
                                  tmx_stream_t s) {
  TMX_REQUIRE(h && img_nchw && out_nhwc16 && N > 0 && H > 0 && W > 0, TMX_ERR_ARG, "tmx_vgg_preprocess: bad argument");
  const long long hw = (long long)H * W, tot = hw * N;
  vgg_pre_kernel<<<tmx_ceil_div(tot, 256), 256, 0, (cudaStream_t)s>>>(img_nchw, out_nhwc16, tot, hw);
  TMX_LAUNCHED(h, "vgg_pre_kernel");
  return TMX_OK;
}

extern "C" int tmx_vgg_preprocess_bwd(tmx_handle_t h, const float* dout_nhwc16, float* dimg_nchw, int N, int H, int W,
                                      tmx_stream_t s) {
  TMX_REQUIRE(h && dout_nhwc16 && dimg_nchw && N > 0 && H > 0 && W > 0, TMX_ERR_ARG,
              "tmx_vgg_preprocess_bwd: bad argument");
  const long long hw = (long long)H * W, tot = hw * N;
  vgg_pre_bwd_kernel<<<tmx_ceil_div(tot, 256), 256, 0, (cudaStream_t)s>>>(dout_nhwc16, dimg_nchw, tot, hw);
  TMX_LAUNCHED(h, "vgg_pre_bwd_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- Gram matrix
// G[n][i][j] = sum_p F[n][i][p] F[n][j][p] / h / w.  One CTA = one T x T tile of one sample, 256 threads, each a
// (T/16) x (T/16) micro-tile; both operand tiles are rows of F (contiguous along p), staged transposed in shared
// memory 16 pixels at a time.
template <int T>
__global__ void __launch_bounds__(256) gram_fwd_kernel(const float* __restrict__ F, float* __restrict__ G, int C, int P,
                                                       float fh, float fw) {
  constexpr int R = T / 16;
  __shared__ float As[16][T + 4], Bs[16][T + 4];
  const int n = blockIdx.z, i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const float* Fn = F + (long long)n * C * P;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[R][R];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) acc[a][b] = 0.f;
  for (int p0 = 0; p0 < P; p0 += 16) {
    // T rows x 16 pixels per operand = T*4 float4: threads 0..T*4-1 load one float4 each (T <= 64 -> <= 256)
    for (int e = threadIdx.x; e < T * 4; e += 256) {
      const int row = e >> 2, q = (e & 3) * 4;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (p0 + q < P) {
        if (i0 + row < C) va = __ldg(reinterpret_cast<const float4*>(Fn + (long long)(i0 + row) * P + p0 + q));
        if (j0 + row < C) vb = __ldg(reinterpret_cast<const float4*>(Fn + (long long)(j0 + row) * P + p0 + q));
      }
      As[q][row] = va.x; As[q + 1][row] = va.y; As[q + 2][row] = va.z; As[q + 3][row] = va.w;
      Bs[q][row] = vb.x; Bs[q + 1][row] = vb.y; Bs[q + 2][row] = vb.z; Bs[q + 3][row] = vb.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[R], b[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        a[r] = As[k][ty * R + r];
        b[r] = Bs[k][tx * R + r];
      }
#pragma unroll
      for (int x = 0; x < R; ++x)
#pragma unroll
        for (int y = 0; y < R; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  float* Gn = G + (long long)n * C * C;
#pragma unroll
  for (int x = 0; x < R; ++x)
#pragma unroll
    for (int y = 0; y < R; ++y) {
      const int i = i0 + ty * R + x, j = j0 + tx * R + y;
      if (i < C && j < C) Gn[(long long)i * C + j] = acc[x][y] / fh / fw;      // loss.py:34: / h / w
    }
}

extern "C" int tmx_gram_fwd(tmx_handle_t h, const float* F, float* G, int N, int C, int H, int W, tmx_stream_t s) {
  TMX_REQUIRE(h && F && G && N > 0 && C > 0 && H > 0 && W > 0, TMX_ERR_ARG, "tmx_gram_fwd: bad argument");
  const int P = H * W;
  TMX_REQUIRE(P % 4 == 0, TMX_ERR_SHAPE, "tmx_gram_fwd: h*w = %d must be a multiple of 4", P);
  if (C <= 128) {
    dim3 grid(tmx_ceil_div(C, 32), tmx_ceil_div(C, 32), N);
    gram_fwd_kernel<32><<<grid, 256, 0, (cudaStream_t)s>>>(F, G, C, P, (float)H, (float)W);
  } else {
    dim3 grid(tmx_ceil_div(C, 64), tmx_ceil_div(C, 64), N);
    gram_fwd_kernel<64><<<grid, 256, 0, (cudaStream_t)s>>>(F, G, C, P, (float)H, (float)W);
  }
  TMX_LAUNCHED(h, "gram_fwd_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- L1 against the target + its gradient
// per sample n: sums[n] += w * val_scale * sum_{i,j} |G[n] - T[n']|,  S[n] (+)= w * coef * sign(G[n] - T[n']),
// n' = N-1-n when reverse_t (tf.reverse(real_gram, axis=[0]), loss.py:252).  tf.abs'(0) = 0.
// w = 1, *wdev or 1 - *wdev (wmode 0 / 1 / 2): the batch-mean mixing weight of loss.py:253-254 lives on the device.
__global__ void __launch_bounds__(256) gram_l1_kernel(const float* __restrict__ G, const float* __restrict__ T,
                                                      float* __restrict__ S, float* __restrict__ sums, int N,
                                                      long long cc, int reverse_t, float coef, float val_scale,
                                                      int accumulate, const float* __restrict__ wdev, int wmode) {
  __shared__ float red[256];
  const float wt = wmode == 0 ? 1.f : (wmode == 1 ? __ldg(wdev) : 1.f - __ldg(wdev));
  coef *= wt;
  val_scale *= wt;
  const int n = blockIdx.y;
  const float* g = G + (long long)n * cc;
  const float* t = T + (long long)(reverse_t ? N - 1 - n : n) * cc;
  float* sp = S + (long long)n * cc;
  float acc = 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cc; e += (long long)gridDim.x * blockDim.x) {
    const float d = __ldg(g + e) - __ldg(t + e);
    acc += fabsf(d);
    const float sg = d > 0.f ? coef : (d < 0.f ? -coef : 0.f);
    sp[e] = accumulate ? sp[e] + sg : sg;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(sums + n, red[0] * val_scale);
}

extern "C" int tmx_gram_l1(tmx_handle_t h, const float* G, const float* T, float* S, float* sums, int N, int C,
                           int reverse_t, float coef, float val_scale, int accumulate, const float* wdev, int wmode,
                           tmx_stream_t s) {
  TMX_REQUIRE(h && G && T && S && sums && N > 0 && C > 0 && wmode >= 0 && wmode <= 2 && (wmode == 0 || wdev),
              TMX_ERR_ARG, "tmx_gram_l1: bad argument");
  const long long cc = (long long)C * C;
  int bx = (int)((cc + 255) / 256);
  if (bx > 64) bx = 64;
  gram_l1_kernel<<<dim3(bx, N), 256, 0, (cudaStream_t)s>>>(G, T, S, sums, N, cc, reverse_t, coef, val_scale, accumulate,
                                                           wdev, wmode);
  TMX_LAUNCHED(h, "gram_l1_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- gradient w.r.t. the features
// dF[n][i][p] = sum_j (S[n][i][j] + S[n][j][i]) F[n][j][p] / h / w     (G = F F^T / h / w)
__global__ void __launch_bounds__(256) gram_bwd_kernel(const float* __restrict__ S, const float* __restrict__ F,
                                                       float* __restrict__ dF, int C, int P, float fh, float fw) {
  constexpr int T = 64, R = 4;
  __shared__ float As[16][T + 4], Bs[16][T + 4];       // As[k = j][i], Bs[k = j][p]
  const int n = blockIdx.z, i0 = blockIdx.y * T, p0 = blockIdx.x * T;
  const float* Sn = S + (long long)n * C * C;
  const float* Fn = F + (long long)n * C * P;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[R][R];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int b = 0; b < R; ++b) acc[a][b] = 0.f;
  for (int j0 = 0; j0 < C; j0 += 16) {
    // A: 16 (j) x 64 (i) symmetric part of S; B: 16 (j) x 64 (p) of F
    for (int e = threadIdx.x; e < 16 * T; e += 256) {
      const int k = e / T, i = e - k * T;
      float v = 0.f;
      if (j0 + k < C && i0 + i < C)
        v = __ldg(Sn + (long long)(i0 + i) * C + j0 + k) + __ldg(Sn + (long long)(j0 + k) * C + i0 + i);
      As[k][i] = v;
    }
    for (int e = threadIdx.x; e < 16 * (T / 4); e += 256) {
      const int k = e / (T / 4), q = (e - k * (T / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + k < C && p0 + q < P) v = __ldg(reinterpret_cast<const float4*>(Fn + (long long)(j0 + k) * P + p0 + q));
      Bs[k][q] = v.x; Bs[k][q + 1] = v.y; Bs[k][q + 2] = v.z; Bs[k][q + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[R], b[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        a[r] = As[k][ty * R + r];
        b[r] = Bs[k][tx * R + r];
      }
#pragma unroll
      for (int x = 0; x < R; ++x)
#pragma unroll
        for (int y = 0; y < R; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  float* dn = dF + (long long)n * C * P;
#pragma unroll
  for (int x = 0; x < R; ++x) {
    const int i = i0 + ty * R + x;
    if (i >= C) continue;
#pragma unroll
    for (int y = 0; y < R; ++y) {
      const int p = p0 + tx * R + y;
      if (p < P) dn[(long long)i * P + p] = acc[x][y] / fh / fw;
    }
  }
}

extern "C" int tmx_gram_bwd(tmx_handle_t h, const float* S, const float* F, float* dF, int N, int C, int H, int W,
                            tmx_stream_t s) {
  TMX_REQUIRE(h && S && F && dF && N > 0 && C > 0 && H > 0 && W > 0, TMX_ERR_ARG, "tmx_gram_bwd: bad argument");
  const int P = H * W;
  TMX_REQUIRE(P % 4 == 0, TMX_ERR_SHAPE, "tmx_gram_bwd: h*w = %d must be a multiple of 4", P);
  dim3 grid(tmx_ceil_div(P, 64), tmx_ceil_div(C, 64), N);
  gram_bwd_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(S, F, dF, C, P, (float)H, (float)W);
  TMX_LAUNCHED(h, "gram_bwd_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- per-sample weight planes of the gradient GEMM
// w[n][i][j] = (S[n][i][j] + S[n][j][i]) * scale -> bf16 hi / lo (the B operand of the TMX_CONV_W_PER_SAMPLE 1x1 conv
// dF[n][p][i] = sum_j F[n][p][j] w[n][i][j]).  32x32 tiles through shared memory: both reads are row-contiguous.
__global__ void __launch_bounds__(256) gram_sym_split_kernel(const float* __restrict__ S, uint16_t* __restrict__ hi,
                                                             uint16_t* __restrict__ lo, int C, float scale) {
  __shared__ float t[32][33];
  const float* Sn = S + (long long)blockIdx.z * C * C;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + tx;
    t[r][tx] = (j < C && i < C) ? __ldg(Sn + (long long)j * C + i) : 0.f;       // S[j][i]
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    if (i < C && j < C) {
      const float v = (__ldg(Sn + (long long)i * C + j) + t[tx][r]) * scale;
      uint32_t a, b;
      tmx_split_bf16(v, a, b);
      const long long o = ((long long)blockIdx.z * C + i) * C + j;
      hi[o] = (uint16_t)a;
      lo[o] = (uint16_t)b;
    }
  }
}

extern "C" int tmx_gram_sym_split(tmx_handle_t h, const float* S, uint16_t* w_hi, uint16_t* w_lo, int N, int C,
                                  float scale, tmx_stream_t s) {
  TMX_REQUIRE(h && S && w_hi && w_lo && N > 0 && C > 0, TMX_ERR_ARG, "tmx_gram_sym_split: bad argument");
  dim3 grid(tmx_ceil_div(C, 32), tmx_ceil_div(C, 32), N);
  gram_sym_split_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(S, w_hi, w_lo, C, scale);
  TMX_LAUNCHED(h, "gram_sym_split_kernel");
  return TMX_OK;
}
