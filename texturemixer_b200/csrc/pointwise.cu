// pointwise.cu — HBM-bound kernels of the TextureMixer hot path: 1x1 RGB heads,
// 2x2 average pool, NCHW<->NHWC moves, split-bf16 halo pack/unpack, weight
// preparation and the latent tile gather/blend (K6).  All are written for
// coalesced 128-bit accesses; none has data reuse worth a TMA pipeline.
#include "common.cuh"

// ---------------------------------------------------------------- FromRGB
// networks.py:226-228: act(bias + conv1x1(x)), x NCHW [N][Cin][H][W] -> y NHWC.
// One thread = one pixel x 4 output channels; 4 (Cout=16) consecutive threads
// write one pixel's contiguous 64 B.
__global__ void __launch_bounds__(256) fromrgb_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float wscale,
                                                      float* __restrict__ y, long long npix, int HW, int Cin,
                                                      int Cout, int lrelu, float alpha) {
  const int groups = Cout >> 2;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = npix * groups;
  if (t >= total) return;
  int g = (int)(t % groups);
  long long p = t / groups;
  long long n = p / HW;
  int hw = (int)(p % HW);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* xp = x + n * (long long)Cin * HW + hw;
  for (int c = 0; c < Cin; ++c) {
    float xv = __ldg(xp + (long long)c * HW);
    float4 wv = __ldg(reinterpret_cast<const float4*>(w + c * Cout + g * 4));
    acc[0] = fmaf(xv, wv.x, acc[0]);
    acc[1] = fmaf(xv, wv.y, acc[1]);
    acc[2] = fmaf(xv, wv.z, acc[2]);
    acc[3] = fmaf(xv, wv.w, acc[3]);
  }
  float4 o;
  float* op = &o.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = acc[i] * wscale + (bias ? __ldg(bias + g * 4 + i) : 0.f);
    if (lrelu) v = fmaxf(v * alpha, v);
    op[i] = v;
  }
  *reinterpret_cast<float4*>(y + p * Cout + g * 4) = o;
}

extern "C" int tmx_fromrgb_fwd(tmx_handle_t h, const float* x, const float* w, const float* bias, float wscale,
                               float* y, int N, int Cin, int H, int W, int Cout, int lrelu, float alpha,
                               tmx_stream_t s) {
  TMX_REQUIRE(h && x && w && y, TMX_ERR_ARG, "tmx_fromrgb_fwd: NULL argument");
  TMX_REQUIRE(N > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0 && Cout % 4 == 0, TMX_ERR_SHAPE,
              "tmx_fromrgb_fwd: bad shape N=%d Cin=%d H=%d W=%d Cout=%d (Cout %% 4 == 0)", N, Cin, H, W, Cout);
  long long npix = (long long)N * H * W;
  long long total = npix * (Cout / 4);
  fromrgb_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(x, w, bias, wscale, y, npix, H * W, Cin, Cout,
                                                                       lrelu, alpha);
  TMX_LAUNCHED(h, "fromrgb_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- ToRGB
// networks.py:454-457 (+ tanh :483): x NHWC [N][H][W][Cin] -> y NCHW [N][Cout<=4][H][W].
// One thread = one pixel: reads Cin*4 contiguous bytes, writes Cout coalesced planes.
template <int COUT>
__global__ void __launch_bounds__(256) torgb_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ bias, float wscale,
                                                    float* __restrict__ y, long long npix, int HW, int Cin,
                                                    int apply_tanh) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
  const float4* xp = reinterpret_cast<const float4*>(x + p * Cin);
  for (int c4 = 0; c4 < (Cin >> 2); ++c4) {
    float4 xv = __ldg(xp + c4);
    const float* xs = &xv.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[o] = fmaf(xs[i], __ldg(w + (c4 * 4 + i) * COUT + o), acc[o]);
    }
  }
  long long n = p / HW;
  int hw = (int)(p % HW);
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float v = acc[o] * wscale + (bias ? __ldg(bias + o) : 0.f);
    if (apply_tanh) v = tanhf(v);
    y[(n * COUT + o) * HW + hw] = v;
  }
}

extern "C" int tmx_torgb_fwd(tmx_handle_t h, const float* x, const float* w, const float* bias, float wscale, float* y,
                             int N, int H, int W, int Cin, int Cout, int apply_tanh, tmx_stream_t s) {
  TMX_REQUIRE(h && x && w && y, TMX_ERR_ARG, "tmx_torgb_fwd: NULL argument");
  TMX_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cin % 4 == 0 && Cout >= 1 && Cout <= 4, TMX_ERR_SHAPE,
              "tmx_torgb_fwd: bad shape N=%d H=%d W=%d Cin=%d Cout=%d (Cin %% 4 == 0, 1 <= Cout <= 4)", N, H, W, Cin,
              Cout);
  long long npix = (long long)N * H * W;
  dim3 grid(tmx_ceil_div(npix, 256));
  cudaStream_t st = (cudaStream_t)s;
  switch (Cout) {
    case 1: torgb_kernel<1><<<grid, 256, 0, st>>>(x, w, bias, wscale, y, npix, H * W, Cin, apply_tanh); break;
    case 2: torgb_kernel<2><<<grid, 256, 0, st>>>(x, w, bias, wscale, y, npix, H * W, Cin, apply_tanh); break;
    case 3: torgb_kernel<3><<<grid, 256, 0, st>>>(x, w, bias, wscale, y, npix, H * W, Cin, apply_tanh); break;
    default: torgb_kernel<4><<<grid, 256, 0, st>>>(x, w, bias, wscale, y, npix, H * W, Cin, apply_tanh); break;
  }
  TMX_LAUNCHED(h, "torgb_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- avg pool 2x2
// networks.py:131-136.  NHWC; one thread = one output pixel x 4 channels.
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                       long long total, int Ho, int Wo, int C4) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int c4 = (int)(t % C4);
  long long p = t / C4;
  int xo = (int)(p % Wo);
  long long q = p / Wo;
  int yo = (int)(q % Ho);
  long long n = q / Ho;
  int W = Wo * 2;
  const float4* base = reinterpret_cast<const float4*>(x) + ((n * (Ho * 2) + yo * 2) * W + xo * 2) * C4 + c4;
  float4 a = __ldg(base), b = __ldg(base + C4), c = __ldg(base + (long long)W * C4),
         d = __ldg(base + (long long)W * C4 + C4);
  float4 o;
  // TF/Eigen avg-pool sums the window then divides by its size
  o.x = (a.x + b.x + c.x + d.x) * 0.25f;
  o.y = (a.y + b.y + c.y + d.y) * 0.25f;
  o.z = (a.z + b.z + c.z + d.z) * 0.25f;
  o.w = (a.w + b.w + c.w + d.w) * 0.25f;
  reinterpret_cast<float4*>(y)[t] = o;
}

extern "C" int tmx_avgpool2_fwd(tmx_handle_t h, const float* x, float* y, int N, int H, int W, int C, tmx_stream_t s) {
  TMX_REQUIRE(h && x && y, TMX_ERR_ARG, "tmx_avgpool2_fwd: NULL argument");
  TMX_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0, TMX_ERR_SHAPE,
              "tmx_avgpool2_fwd: bad shape N=%d H=%d W=%d C=%d (even H, W; C %% 4 == 0)", N, H, W, C);
  long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
  avgpool2_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(x, y, total, H / 2, W / 2, C / 4);
  TMX_LAUNCHED(h, "avgpool2_kernel");
  return TMX_OK;
}

// downscale2d (networks.py:131-136) whose consumer is a tensor-core conv: the pooled map as fp32 NHWC (optional) AND as
// split bf16 planes with the consumer's halo in ONE pass (was: avgpool2 -> fp32, then split_halo_pack re-reading it).
// One thread = one PADDED output pixel x 8 channels; halo threads average the 2x2 window of their mirrored / clamped
// source pixel again (a ring's worth of extra reads).
__global__ void __launch_bounds__(256) avgpool2_pack_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                            long long total, int Ho, int Wo, int C8, int halo_kind) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c8 = (int)(t % C8);
  long long p = t / C8;
  const int Wp = Wo + 2, Hp = Ho + 2;
  const int xp = (int)(p % Wp);
  long long q = p / Wp;
  const int yp = (int)(q % Hp);
  const long long n = q / Hp;
  const bool ring = yp == 0 || yp == Hp - 1 || xp == 0 || xp == Wp - 1;
  const int ys = halo_kind == 1 ? min(max(yp - 1, 0), Ho - 1) : tmx_reflect(yp - 1, Ho);
  const int xs = halo_kind == 1 ? min(max(xp - 1, 0), Wo - 1) : tmx_reflect(xp - 1, Wo);
  const int W = Wo * 2;
  const float4* base = reinterpret_cast<const float4*>(x) + ((n * (Ho * 2) + ys * 2) * W + xs * 2) * (C8 * 2) + c8 * 2;
  float v[8];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float4 a = __ldg(base + half), b = __ldg(base + C8 * 2 + half),
                 c = __ldg(base + (long long)W * C8 * 2 + half), d = __ldg(base + (long long)W * C8 * 2 + C8 * 2 + half);
    // TF/Eigen avg-pool sums the window then divides by its size (same order as avgpool2_kernel)
    v[4 * half] = (a.x + b.x + c.x + d.x) * 0.25f;
    v[4 * half + 1] = (a.y + b.y + c.y + d.y) * 0.25f;
    v[4 * half + 2] = (a.z + b.z + c.z + d.z) * 0.25f;
    v[4 * half + 3] = (a.w + b.w + c.w + d.w) * 0.25f;
  }
  if (!ring && y != nullptr) {
    float4* o = reinterpret_cast<float4*>(y) + ((n * Ho + ys) * Wo + xs) * (C8 * 2) + c8 * 2;
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (halo_kind == 2 && ring) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tmx_split_bf16x2(v[2 * i], v[2 * i + 1], ph[i], pl[i]);
  reinterpret_cast<uint4*>(hi)[t] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  reinterpret_cast<uint4*>(lo)[t] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

extern "C" int tmx_avgpool2_pack(tmx_handle_t h, const float* x, float* y, uint16_t* hi, uint16_t* lo, int N, int H, int W,
                                 int C, int halo_kind, tmx_stream_t s) {
  TMX_REQUIRE(h && x && hi && lo, TMX_ERR_ARG, "tmx_avgpool2_pack: NULL argument");
  TMX_REQUIRE(halo_kind >= 0 && halo_kind <= 2, TMX_ERR_ARG, "tmx_avgpool2_pack: halo kind %d not in 0..2", halo_kind);
  TMX_REQUIRE(N > 0 && H >= 4 && W >= 4 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, TMX_ERR_SHAPE,
              "tmx_avgpool2_pack: bad shape N=%d H=%d W=%d C=%d (even H, W >= 4: the pooled map needs 2 x 2 pixels for its "
              "halo; C %% 8 == 0)", N, H, W, C);
  const long long total = (long long)N * (H / 2 + 2) * (W / 2 + 2) * (C / 8);
  avgpool2_pack_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(x, y, hi, lo, total, H / 2, W / 2, C / 8,
                                                                              halo_kind);
  TMX_LAUNCHED(h, "avgpool2_pack_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- NCHW <-> NHWC
// 32x32 smem tile transpose per image between [C][HW] and [HW][C_total] (+c_off).
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C,
                                                           int HW, int c_off, int C_total, int bcast) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* xn = x + (long long)n * C * (bcast ? 1 : HW);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + ty + i * 8, p = p0 + tx;
    if (c < C && p < HW) tile[ty + i * 8][tx] = bcast ? __ldg(xn + c) : __ldg(xn + (long long)c * HW + p);
  }
  __syncthreads();
  float* yn = y + (long long)n * HW * C_total + c_off;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int p = p0 + ty + i * 8, c = c0 + tx;
    if (c < C && p < HW) yn[(long long)p * C_total + c] = tile[tx][ty + i * 8];
  }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C,
                                                           int HW, int c_off, int C_total) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xn = x + (long long)n * HW * C_total + c_off;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int p = p0 + ty + i * 8, c = c0 + tx;
    if (c < C && p < HW) tile[ty + i * 8][tx] = __ldg(xn + (long long)p * C_total + c);
  }
  __syncthreads();
  float* yn = y + (long long)n * C * HW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + ty + i * 8, p = p0 + tx;
    if (c < C && p < HW) yn[(long long)c * HW + p] = tile[tx][ty + i * 8];
  }
}

extern "C" int tmx_nchw_to_nhwc(tmx_handle_t h, const float* x, float* y, int N, int C, int H, int W, int c_off,
                                int C_total, int bcast_hw, tmx_stream_t s) {
  TMX_REQUIRE(h && x && y, TMX_ERR_ARG, "tmx_nchw_to_nhwc: NULL argument");
  TMX_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && c_off >= 0 && c_off + C <= C_total && N <= 65535, TMX_ERR_SHAPE,
              "tmx_nchw_to_nhwc: bad shape N=%d C=%d H=%d W=%d c_off=%d C_total=%d", N, C, H, W, c_off, C_total);
  dim3 grid(tmx_ceil_div(H * W, 32), tmx_ceil_div(C, 32), N);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, y, C, H * W, c_off, C_total, bcast_hw);
  TMX_LAUNCHED(h, "nchw_to_nhwc_kernel");
  return TMX_OK;
}

extern "C" int tmx_nhwc_to_nchw(tmx_handle_t h, const float* x, float* y, int N, int C, int H, int W, int c_off,
                                int C_total, tmx_stream_t s) {
  TMX_REQUIRE(h && x && y, TMX_ERR_ARG, "tmx_nhwc_to_nchw: NULL argument");
  TMX_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && c_off >= 0 && c_off + C <= C_total && N <= 65535, TMX_ERR_SHAPE,
              "tmx_nhwc_to_nchw: bad shape N=%d C=%d H=%d W=%d c_off=%d C_total=%d", N, C, H, W, c_off, C_total);
  dim3 grid(tmx_ceil_div(H * W, 32), tmx_ceil_div(C, 32), N);
  nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, y, C, H * W, c_off, C_total);
  TMX_LAUNCHED(h, "nhwc_to_nchw_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- split-bf16 halo pack / unpack
// One thread = one padded pixel x 8 channels (2 float4 in, 16 B hi + 16 B lo out).
__global__ void __launch_bounds__(256) split_halo_pack_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi,
                                                              uint16_t* __restrict__ lo, long long total, int H, int W,
                                                              int C8, int halo_kind) {
  // halo_kind 0: REFLECT (networks.py:55), 1: REPLICATE (REFLECT seen through upscale2d), 2: ZERO (the SAME
  // padding of the fused_scale convs, networks.py:101,148)
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int c8 = (int)(t % C8);
  long long p = t / C8;
  int Wp = W + 2, Hp = H + 2;
  int xp = (int)(p % Wp);
  long long q = p / Wp;
  int yp = (int)(q % Hp);
  long long n = q / Hp;
  const bool replicate = halo_kind == 1;
  int ys = replicate ? min(max(yp - 1, 0), H - 1) : tmx_reflect(yp - 1, H);
  int xs = replicate ? min(max(xp - 1, 0), W - 1) : tmx_reflect(xp - 1, W);
  const float4* src = reinterpret_cast<const float4*>(x) + ((n * H + ys) * W + xs) * (C8 * 2) + c8 * 2;
  float4 a = __ldg(src), b = __ldg(src + 1);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (halo_kind == 2 && (yp == 0 || yp == Hp - 1 || xp == 0 || xp == Wp - 1)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    tmx_split_bf16x2(v[2 * i], v[2 * i + 1], ph[i], pl[i]);
  }
  reinterpret_cast<uint4*>(hi)[t] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  reinterpret_cast<uint4*>(lo)[t] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

__global__ void __launch_bounds__(256) split_halo_unpack_kernel(const uint16_t* __restrict__ hi,
                                                                const uint16_t* __restrict__ lo, float* __restrict__ y,
                                                                long long total, int H, int W, int C8) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int c8 = (int)(t % C8);
  long long p = t / C8;
  int xo = (int)(p % W);
  long long q = p / W;
  int yo = (int)(q % H);
  long long n = q / H;
  long long src = ((n * (H + 2) + yo + 1) * (W + 2) + xo + 1) * C8 + c8;
  uint4 a = __ldg(reinterpret_cast<const uint4*>(hi) + src);
  uint4 b = __ldg(reinterpret_cast<const uint4*>(lo) + src);
  const uint32_t ah[4] = {a.x, a.y, a.z, a.w}, al[4] = {b.x, b.y, b.z, b.w};
  float v[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(ah[i] << 16) + __uint_as_float(al[i] << 16);
    v[2 * i + 1] = __uint_as_float(ah[i] & 0xffff0000u) + __uint_as_float(al[i] & 0xffff0000u);
  }
  float4* dst = reinterpret_cast<float4*>(y) + t * 2;
  dst[0] = make_float4(v[0], v[1], v[2], v[3]);
  dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

extern "C" int tmx_split_halo_pack(tmx_handle_t h, const float* x, uint16_t* hi, uint16_t* lo, int N, int H, int W,
                                   int C, int replicate, tmx_stream_t s) {
  TMX_REQUIRE(h && x && hi && lo, TMX_ERR_ARG, "tmx_split_halo_pack: NULL argument");
  TMX_REQUIRE(replicate >= 0 && replicate <= 2, TMX_ERR_ARG, "tmx_split_halo_pack: halo kind %d not in 0..2", replicate);
  TMX_REQUIRE(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0, TMX_ERR_SHAPE,
              "tmx_split_halo_pack: bad shape N=%d H=%d W=%d C=%d (H, W >= 2; C %% 8 == 0)", N, H, W, C);
  long long total = (long long)N * (H + 2) * (W + 2) * (C / 8);
  split_halo_pack_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(x, hi, lo, total, H, W, C / 8,
                                                                                replicate);
  TMX_LAUNCHED(h, "split_halo_pack_kernel");
  return TMX_OK;
}

extern "C" int tmx_split_halo_unpack(tmx_handle_t h, const uint16_t* hi, const uint16_t* lo, float* y, int N, int H,
                                     int W, int C, tmx_stream_t s) {
  TMX_REQUIRE(h && y && hi && lo, TMX_ERR_ARG, "tmx_split_halo_unpack: NULL argument");
  TMX_REQUIRE(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0, TMX_ERR_SHAPE,
              "tmx_split_halo_unpack: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  long long total = (long long)N * H * W * (C / 8);
  split_halo_unpack_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(hi, lo, y, total, H, W, C / 8);
  TMX_LAUNCHED(h, "split_halo_unpack_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- weight preparation (tensor-core path)
// w HWIO [taps][Cin][Cout] fp32 -> hi/lo bf16 [Cout][taps*Cin], scaled by wscale.
// 32x32 tile transpose: coalesced reads along Cout, coalesced writes along K.
// Cin_pad > Cin: every tap's channel run is zero-extended to Cin_pad (activation padded to a multiple of 16).
__global__ void __launch_bounds__(256) weights_prepare_padded_kernel(const float* __restrict__ w, float wscale,
                                                                     uint16_t* __restrict__ hi,
                                                                     uint16_t* __restrict__ lo, int taps, int Cin,
                                                                     int Cin_pad, int Cout) {
  const long long Kp = (long long)taps * Cin_pad;
  const long long total = (long long)Cout * Kp;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int kk = (int)(t % Kp);
  const int o = (int)(t / Kp);
  const int c = kk % Cin_pad, tap = kk / Cin_pad;
  const float v = c < Cin ? __ldg(w + ((long long)tap * Cin + c) * Cout + o) * wscale : 0.f;
  uint32_t a, b;
  tmx_split_bf16(v, a, b);
  hi[t] = (uint16_t)a;
  lo[t] = (uint16_t)b;
}

__global__ void __launch_bounds__(256) weights_prepare_kernel(const float* __restrict__ w, float wscale,
                                                              uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                              int K, int Cout) {
  __shared__ float tile[32][33];
  int k0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + ty + i * 8, o = o0 + tx;
    if (k < K && o < Cout) tile[ty + i * 8][tx] = __ldg(w + (long long)k * Cout + o);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int o = o0 + ty + i * 8, k = k0 + tx;
    if (k < K && o < Cout) {
      uint32_t a, b;
      tmx_split_bf16(tile[tx][ty + i * 8] * wscale, a, b);
      hi[(long long)o * K + k] = (uint16_t)a;
      lo[(long long)o * K + k] = (uint16_t)b;
    }
  }
}

// Sub-pixel weights of conv3x3(upscale2d(x)) (see tmx.h): out [4*Cout][9*Cin], one thread per element.
__global__ void __launch_bounds__(256) weights_prepare_phase_kernel(const float* __restrict__ w, float wscale,
                                                                    uint16_t* __restrict__ hi,
                                                                    uint16_t* __restrict__ lo, int Cin, int Cout) {
  const long long K = 9ll * Cin;
  const long long total = 4ll * Cout * K;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int kk = (int)(t % K);
  const int row = (int)(t / K);
  const int c = kk % Cin, tap = kk / Cin;
  const int U = tap / 3, V = tap % 3;
  const int o = row % Cout, ph = row / Cout;
  const int a = ph >> 1, b = ph & 1;
  // upsampled taps u that land on low-res offset U-1 for output parity a: a=0: {0},{1,2},{}; a=1: {},{0,1},{2}
  const int u0 = a == 0 ? (U == 0 ? 0 : (U == 1 ? 1 : 3)) : (U == 0 ? 3 : (U == 1 ? 0 : 2));
  const int u1 = a == 0 ? (U == 0 ? 0 : (U == 1 ? 2 : 2)) : (U == 0 ? 2 : (U == 1 ? 1 : 2));
  const int v0 = b == 0 ? (V == 0 ? 0 : (V == 1 ? 1 : 3)) : (V == 0 ? 3 : (V == 1 ? 0 : 2));
  const int v1 = b == 0 ? (V == 0 ? 0 : (V == 1 ? 2 : 2)) : (V == 0 ? 2 : (V == 1 ? 1 : 2));
  float acc = 0.f;
  for (int u = u0; u <= u1; ++u)
    for (int v = v0; v <= v1; ++v) acc += __ldg(w + ((long long)(u * 3 + v) * Cin + c) * Cout + o);
  uint32_t x, y;
  tmx_split_bf16(acc * wscale, x, y);
  hi[t] = (uint16_t)x;
  lo[t] = (uint16_t)y;
}

extern "C" int tmx_conv_weights_prepare(tmx_handle_t h, const float* w, float wscale, int k, int Cin, int Cin_pad,
                                        int Cout, int up2_phase, uint16_t* w_hi, uint16_t* w_lo, tmx_stream_t s) {
  TMX_REQUIRE(h && w && w_hi && w_lo, TMX_ERR_ARG, "tmx_conv_weights_prepare: NULL argument");
  TMX_REQUIRE((k == 1 || k == 3) && Cin > 0 && Cout > 0 && Cin_pad >= Cin, TMX_ERR_SHAPE,
              "tmx_conv_weights_prepare: bad shape k=%d Cin=%d Cin_pad=%d Cout=%d", k, Cin, Cin_pad, Cout);
  if (Cin_pad > Cin) {
    TMX_REQUIRE(!up2_phase, TMX_ERR_UNSUPPORTED, "tmx_conv_weights_prepare: Cin padding with up2_phase");
    long long total = (long long)Cout * k * k * Cin_pad;
    weights_prepare_padded_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(w, wscale, w_hi, w_lo, k * k,
                                                                                        Cin, Cin_pad, Cout);
    TMX_LAUNCHED(h, "weights_prepare_padded_kernel");
    return TMX_OK;
  }
  if (up2_phase) {
    TMX_REQUIRE(k == 3, TMX_ERR_SHAPE, "tmx_conv_weights_prepare: up2_phase needs k == 3");
    long long total = 36ll * Cin * Cout;
    weights_prepare_phase_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(w, wscale, w_hi, w_lo, Cin,
                                                                                       Cout);
    TMX_LAUNCHED(h, "weights_prepare_phase_kernel");
    return TMX_OK;
  }
  int K = k * k * Cin;
  dim3 grid(tmx_ceil_div(K, 32), tmx_ceil_div(Cout, 32));
  weights_prepare_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(w, wscale, w_hi, w_lo, K, Cout);
  TMX_LAUNCHED(h, "weights_prepare_kernel");
  return TMX_OK;
}

__global__ void __launch_bounds__(256) weights_prepare_xmerge_kernel(const float* __restrict__ w, float wscale,
                                                                     uint16_t* __restrict__ hi,
                                                                     uint16_t* __restrict__ lo, int Cin, int Cout) {
  const int total = Cout * 192;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int kk = t % 192, o = t / 192;
  const int u = kk / 64, r = kk % 64, v = r / 16, c = r % 16;
  // channels >= Cin are the zero padding of a narrower input (VGG-19's conv1_1: 3 -> 16)
  const float val = (v < 3 && c < Cin) ? __ldg(w + ((long long)((u * 3 + v) * Cin + c)) * Cout + o) * wscale : 0.f;
  uint32_t a, b;
  tmx_split_bf16(val, a, b);
  hi[t] = (uint16_t)a;
  lo[t] = (uint16_t)b;
}

extern "C" int tmx_conv_weights_prepare_xmerge(tmx_handle_t h, const float* w, float wscale, int Cin, int Cout,
                                               uint16_t* w_hi, uint16_t* w_lo, tmx_stream_t s) {
  TMX_REQUIRE(h && w && w_hi && w_lo && Cout > 0 && Cin > 0 && Cin <= 16, TMX_ERR_ARG,
              "tmx_conv_weights_prepare_xmerge: bad argument");
  weights_prepare_xmerge_kernel<<<tmx_ceil_div(Cout * 192, 256), 256, 0, (cudaStream_t)s>>>(w, wscale, w_hi, w_lo, Cin,
                                                                                            Cout);
  TMX_LAUNCHED(h, "weights_prepare_xmerge_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- latent tile gather / blend (K6)
// Block = (32-wide strip of canvas columns, one canvas row i, one sample n);
// 8 warps sweep the channels, lane = column.  Values are written straight to
// the NCHW canvas (coalesced along j) and/or staged in smem and written to the
// NHWC trunk input (coalesced along c).
struct BlendParams {
  tmx_blend_desc_t d;
  tmx_blend_io_t io;
};

template <int MODE, bool F32>
__global__ void __launch_bounds__(256) latent_blend_kernel(const BlendParams P) {
  extern __shared__ float tile[];  // [Cchunk][33]
  const tmx_blend_desc_t& d = P.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.z, i = blockIdx.y, j0 = blockIdx.x * 32, j = j0 + lane;
  const bool jv = j < d.W;
  const bool pin_row = d.pin_rows != 0 && ((d.pin_rows >> (i / d.h)) & 1ull);
  int sy[4], sx[4];
  double wgt[4];
  float wgtf[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sy[k] = sx[k] = 0;
    wgt[k] = 0.0;
    wgtf[k] = 0.f;
    if (k < d.K && jv) {
      bool pinned = pin_row && ((d.pin_cols >> (j / d.w)) & 1ull);
      int yy = i, xx = j;
      if (!pinned) {
        if (P.io.idx_h[k]) yy = __ldg(P.io.idx_h[k] + (long long)n * d.H + i);
        if (P.io.idx_w[k]) xx = __ldg(P.io.idx_w[k] + (long long)n * d.W + j);
      }
      sy[k] = yy % d.h;
      sx[k] = xx % d.w;
      if (MODE == TMX_BLEND_MATTE) {
        double rh = __ldg(P.io.ramp_h[k] + i), rw = __ldg(P.io.ramp_w[k] + j);
        if (F32) wgtf[k] = __fmul_rn((float)rh, (float)rw);
        else wgt[k] = __dmul_rn(rh, rw);
      }
    }
  }
  const int CCH = 128;
  for (int cbase = 0; cbase < d.C; cbase += CCH) {
    const int cn = min(CCH, d.C - cbase);
    for (int cc = warp; cc < cn; cc += 8) {
      const int c = cbase + cc;
      float v = 0.f;
      if (jv) {
        if (MODE == TMX_BLEND_MATTE) {
          float sv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < d.K) {
              const int nn = ((d.src_reverse >> k) & 1u) ? d.N - 1 - n : n;
              const float* s = P.io.src[k] + ((long long)nn * d.C + c) * (d.src_bcast ? 1 : d.h * d.w);
              sv[k] = d.src_bcast ? __ldg(s) : __ldg(s + sy[k] * d.w + sx[k]);
            }
          }
          if (F32) {
            float acc = __fmul_rn(sv[0], wgtf[0]);
#pragma unroll
            for (int k = 1; k < 4; ++k)
              if (k < d.K) acc = __fadd_rn(acc, __fmul_rn(sv[k], wgtf[k]));
            v = acc;
          } else {
            double acc = __dmul_rn((double)sv[0], wgt[0]);
#pragma unroll
            for (int k = 1; k < 4; ++k)
              if (k < d.K) acc = __dadd_rn(acc, __dmul_rn((double)sv[k], wgt[k]));
            v = (float)acc;
          }
        } else {
          float sv[2] = {0.f, 0.f};
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k < d.K) {
              const int nn = ((d.src_reverse >> k) & 1u) ? d.N - 1 - n : n;
              const float* s = P.io.src[k] + ((long long)nn * d.C + c) * (d.src_bcast ? 1 : d.h * d.w);
              sv[k] = d.src_bcast ? __ldg(s) : __ldg(s + sy[k] * d.w + sx[k]);
            }
          }
          if (MODE == TMX_BLEND_LERP) {
            float t = __ldg(P.io.t + n);
            v = __fadd_rn(sv[0], __fmul_rn(__fsub_rn(sv[1], sv[0]), t));  // tfutil.py:41-43, unfused
          } else {
            v = sv[0];
          }
        }
        if (P.io.out_nchw) P.io.out_nchw[(((long long)n * d.C + c) * d.H + i) * d.W + j] = v;
      }
      if (P.io.out_nhwc) tile[cc * 33 + lane] = v;
    }
    if (P.io.out_nhwc) {
      __syncthreads();
      const int jn = min(32, d.W - j0);
      float* o = P.io.out_nhwc + (((long long)n * d.H + i) * d.W + j0) * d.C_total + d.c_off + cbase;
      for (int e = threadIdx.x; e < jn * cn; e += 256) {
        int jj = e / cn, cc = e % cn;
        o[(long long)jj * d.C_total + cc] = tile[cc * 33 + jj];
      }
      __syncthreads();
    }
  }
}

// TILED form for NCHW outputs whose source tiles fit in shared memory (h*w floats per channel: the 32 x 32 latent
// tiles of the path).  The row-per-block kernel above re-reads every source line once per canvas position it feeds
// (K = 4 sources: four 128-B lines from L2 per 128-B line written) and spends ~90 instructions per output on 64-bit
// address arithmetic.  Here a block owns (sample n, `cg` consecutive channels, 32 canvas columns): it stages the
// K * cg source tiles once (contiguous in NCHW) plus per-row tables (source row offset, row ramp, re-pin flag), then
// walks ALL canvas rows out of shared memory - every source byte crosses L2 -> SM once per column tile, and the
// index work of a canvas position is shared by its cg channels.  Column indices / column ramps are per-lane
// constants of the block.  Same arithmetic, same order as the kernel above (bit-identical results).
template <int MODE, bool F32>
__global__ void __launch_bounds__(256) latent_blend_tiled_kernel(const BlendParams P, int cg) {
  extern __shared__ __align__(16) uint8_t blend_smem[];
  const tmx_blend_desc_t& d = P.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.z, c0 = blockIdx.y * cg, j0 = blockIdx.x * 32, j = j0 + lane;
  const int cn = min(cg, d.C - c0);
  const int hw = d.h * d.w;
  const bool jv = j < d.W;
  float* src_s = reinterpret_cast<float*>(blend_smem);                                   // [K][cg][h*w]
  double* rh_s = reinterpret_cast<double*>(blend_smem + (size_t)d.K * cg * hw * 4);      // [K][H] row ramps
  int* syw_s = reinterpret_cast<int*>(rh_s + (size_t)d.K * d.H);                         // [K][H] (source row) * w
  int* piw_s = syw_s + (size_t)d.K * d.H;                                                // [H] (i % h) * w, < 0: row not re-pinned
  for (int k = 0; k < d.K; ++k) {
    const int nn = ((d.src_reverse >> k) & 1u) ? d.N - 1 - n : n;
    const float4* g = reinterpret_cast<const float4*>(P.io.src[k] + ((long long)nn * d.C + c0) * hw);
    float4* t = reinterpret_cast<float4*>(src_s + (size_t)k * cg * hw);
    for (int e = threadIdx.x; e < cn * hw / 4; e += 256) t[e] = __ldg(g + e);
  }
  for (int e = threadIdx.x; e < d.K * d.H; e += 256) {
    const int k = e / d.H, i = e - k * d.H;
    const int yy = P.io.idx_h[k] ? __ldg(P.io.idx_h[k] + (long long)n * d.H + i) : i;
    syw_s[e] = (yy % d.h) * d.w;
    rh_s[e] = MODE == TMX_BLEND_MATTE ? __ldg(P.io.ramp_h[k] + i) : 0.0;
  }
  for (int i = threadIdx.x; i < d.H; i += 256)
    piw_s[i] = (d.pin_rows != 0 && ((d.pin_rows >> (i / d.h)) & 1ull)) ? (i % d.h) * d.w : -1;
  // per-lane constants of this column tile
  int sxf[4], sxp = 0;
  double rwd[4];
  bool pcol = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sxf[k] = 0;
    rwd[k] = 0.0;
  }
  if (jv) {
    pcol = d.pin_rows != 0 && ((d.pin_cols >> (j / d.w)) & 1ull);
    sxp = j % d.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < d.K) {
        const int xx = P.io.idx_w[k] ? __ldg(P.io.idx_w[k] + (long long)n * d.W + j) : j;
        sxf[k] = xx % d.w;
        if (MODE == TMX_BLEND_MATTE) rwd[k] = __ldg(P.io.ramp_w[k] + j);
      }
    }
  }
  float tl = 0.f;
  if (MODE == TMX_BLEND_LERP) tl = __ldg(P.io.t + n);
  __syncthreads();
  if (!jv) return;
  float* out0 = P.io.out_nchw + ((long long)n * d.C + c0) * d.H * d.W + j;
  const int plane = d.H * d.W;
  for (int i = warp; i < d.H; i += 8) {
    const int piw = piw_s[i];
    const bool pinned = pcol && piw >= 0;
    int off[4];
    double wgt[4];
    float wgtf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      off[k] = 0;
      wgt[k] = 0.0;
      wgtf[k] = 0.f;
      if (k < d.K) {
        off[k] = pinned ? piw + sxp : syw_s[k * d.H + i] + sxf[k];
        if (MODE == TMX_BLEND_MATTE) {
          const double rh = rh_s[k * d.H + i];
          if (F32) wgtf[k] = __fmul_rn((float)rh, (float)rwd[k]);
          else wgt[k] = __dmul_rn(rh, rwd[k]);
        }
      }
    }
    float* o = out0 + i * d.W;
    for (int cc = 0; cc < cn; ++cc) {
      const float* sc = src_s + cc * hw;
      float sv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < d.K) sv[k] = sc[k * cg * hw + off[k]];
      float v;
      if (MODE == TMX_BLEND_MATTE) {
        if (F32) {
          float acc = __fmul_rn(sv[0], wgtf[0]);
#pragma unroll
          for (int k = 1; k < 4; ++k)
            if (k < d.K) acc = __fadd_rn(acc, __fmul_rn(sv[k], wgtf[k]));
          v = acc;
        } else {
          double acc = __dmul_rn((double)sv[0], wgt[0]);
#pragma unroll
          for (int k = 1; k < 4; ++k)
            if (k < d.K) acc = __dadd_rn(acc, __dmul_rn((double)sv[k], wgt[k]));
          v = (float)acc;
        }
      } else if (MODE == TMX_BLEND_LERP) {
        v = __fadd_rn(sv[0], __fmul_rn(__fsub_rn(sv[1], sv[0]), tl));  // tfutil.py:41-43, unfused
      } else {
        v = sv[0];
      }
      o[(long long)cc * plane] = v;
    }
  }
}

template <int MODE, bool F32>
int launch_blend_tiled(tmx_handle_t h, const BlendParams& P, int cg, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = latent_blend_tiled_kernel<MODE, F32>;
  static thread_local int configured_device = -1;
  if (configured_device != h->device) {
    TMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured_device = h->device;
  }
  kern<<<grid, 256, smem, st>>>(P, cg);
  TMX_LAUNCHED(h, "latent_blend_tiled_kernel");
  return TMX_OK;
}

extern "C" int tmx_latent_blend(tmx_handle_t h, const tmx_blend_desc_t* d, const tmx_blend_io_t* io, tmx_stream_t s) {
  TMX_REQUIRE(h && d && io, TMX_ERR_ARG, "tmx_latent_blend: NULL argument");
  TMX_REQUIRE(d->N > 0 && d->C > 0 && d->h > 0 && d->w > 0 && d->H > 0 && d->W > 0 && d->N <= 65535 && d->H <= 65535,
              TMX_ERR_SHAPE, "tmx_latent_blend: bad shape N=%d C=%d h=%d w=%d H=%d W=%d", d->N, d->C, d->h, d->w, d->H,
              d->W);
  TMX_REQUIRE(d->K >= 1 && d->K <= 4, TMX_ERR_SHAPE, "tmx_latent_blend: K=%d not in 1..4", d->K);
  TMX_REQUIRE(d->mode == TMX_BLEND_COPY || d->mode == TMX_BLEND_MATTE || d->mode == TMX_BLEND_LERP, TMX_ERR_ARG,
              "tmx_latent_blend: bad mode %d", d->mode);
  TMX_REQUIRE(d->mode != TMX_BLEND_COPY || d->K == 1, TMX_ERR_SHAPE, "tmx_latent_blend: COPY needs K == 1");
  TMX_REQUIRE(d->mode != TMX_BLEND_LERP || (d->K == 2 && io->t), TMX_ERR_SHAPE,
              "tmx_latent_blend: LERP needs K == 2 and t");
  TMX_REQUIRE((d->pin_rows == 0 && d->pin_cols == 0) ||
                  ((d->H + d->h - 1) / d->h <= 64 && (d->W + d->w - 1) / d->w <= 64),
              TMX_ERR_SHAPE, "tmx_latent_blend: re-pinned tiles need at most 64 tiles per side");
  TMX_REQUIRE(io->out_nchw || io->out_nhwc, TMX_ERR_ARG, "tmx_latent_blend: no output");
  TMX_REQUIRE(!io->out_nhwc || (d->c_off >= 0 && d->c_off + d->C <= d->C_total), TMX_ERR_SHAPE,
              "tmx_latent_blend: bad NHWC slice c_off=%d C=%d C_total=%d", d->c_off, d->C, d->C_total);
  for (int k = 0; k < d->K; ++k) {
    TMX_REQUIRE(io->src[k] != nullptr, TMX_ERR_ARG, "tmx_latent_blend: src[%d] is NULL", k);
    if (d->mode == TMX_BLEND_MATTE)
      TMX_REQUIRE(io->ramp_h[k] && io->ramp_w[k], TMX_ERR_ARG, "tmx_latent_blend: ramp[%d] is NULL", k);
  }
  BlendParams P;
  P.d = *d;
  P.io = *io;
  cudaStream_t st = (cudaStream_t)s;
  // tiled form: NCHW output only, real (not broadcast) source tiles; K * cg tiles + the row tables in <= 96 KB
  {
    const long long hw = (long long)d->h * d->w;
    bool aligned = hw % 4 == 0 && (long long)d->H * d->W < (1ll << 31) / 8;
    for (int k = 0; k < d->K; ++k) aligned = aligned && ((uintptr_t)io->src[k] & 15) == 0;
    const long long tables = (long long)d->K * d->H * 12 + (long long)d->H * 4;
    int cg = 0;
    if (!io->out_nhwc && !d->src_bcast && aligned && !tmx_env_flag("TMX_BLEND_ROWS"))
      for (int c = 4; c >= 1; c /= 2)
        if ((long long)d->K * c * hw * 4 + tables <= 80 * 1024 && c <= d->C) {
          cg = c;
          break;
        }
    if (cg > 0 && tmx_ceil_div(d->C, cg) <= 65535) {
      dim3 grid(tmx_ceil_div(d->W, 32), tmx_ceil_div(d->C, cg), d->N);
      const size_t smem = (size_t)d->K * cg * hw * sizeof(float) + (size_t)tables;
      if (d->mode == TMX_BLEND_COPY) return launch_blend_tiled<TMX_BLEND_COPY, false>(h, P, cg, grid, smem, st);
      if (d->mode == TMX_BLEND_LERP) return launch_blend_tiled<TMX_BLEND_LERP, false>(h, P, cg, grid, smem, st);
      if (d->math_f32) return launch_blend_tiled<TMX_BLEND_MATTE, true>(h, P, cg, grid, smem, st);
      return launch_blend_tiled<TMX_BLEND_MATTE, false>(h, P, cg, grid, smem, st);
    }
  }
  dim3 grid(tmx_ceil_div(d->W, 32), d->H, d->N);
  size_t smem = io->out_nhwc ? (size_t)128 * 33 * sizeof(float) : 0;
  if (d->mode == TMX_BLEND_COPY) latent_blend_kernel<TMX_BLEND_COPY, false><<<grid, 256, smem, st>>>(P);
  else if (d->mode == TMX_BLEND_LERP) latent_blend_kernel<TMX_BLEND_LERP, false><<<grid, 256, smem, st>>>(P);
  else if (d->math_f32) latent_blend_kernel<TMX_BLEND_MATTE, true><<<grid, 256, smem, st>>>(P);
  else latent_blend_kernel<TMX_BLEND_MATTE, false><<<grid, 256, smem, st>>>(P);
  TMX_LAUNCHED(h, "latent_blend_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- minibatch stddev (D_patch)
// networks.py:177-189.  x NHWC [N][H][W][C] -> y NHWC [N][H][W][C_total]: channels [0,C) copy x, channel C
// holds the group statistic s[n % M] (M = N / G groups; group m = samples {m, m+M, ...}), channels
// (C, C_total) are zero padding so that the consumer conv sees a multiple of 16 input channels.
// One block per group m: s[m] = mean_{h,w,c} sqrt( mean_g (x - mean_g x)^2 + 1e-8 ).
__global__ void __launch_bounds__(256) mbstd_stat_kernel(const float* __restrict__ x, float* __restrict__ stat, int G,
                                                         int M, long long per_sample) {
  const int m = blockIdx.x;
  float acc = 0.f;
  for (long long e = threadIdx.x; e < per_sample; e += blockDim.x) {
    float mean = 0.f;
    for (int g = 0; g < G; ++g) mean += __ldg(x + ((long long)(g * M + m)) * per_sample + e);
    mean /= (float)G;
    float var = 0.f;
    for (int g = 0; g < G; ++g) {
      const float d = __ldg(x + ((long long)(g * M + m)) * per_sample + e) - mean;
      var += d * d;
    }
    acc += sqrtf(var / (float)G + 1e-8f);
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) stat[m] = red[0] / (float)per_sample;
}

__global__ void __launch_bounds__(256) mbstd_concat_kernel(const float* __restrict__ x, const float* __restrict__ stat,
                                                           float* __restrict__ y, long long total, int C, int C_total,
                                                           int M, long long pix_per_sample) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C_total);
  const long long pix = t / C_total;
  float v = 0.f;
  if (c < C) v = __ldg(x + pix * C + c);
  else if (c == C) v = __ldg(stat + (int)((pix / pix_per_sample) % M));
  y[t] = v;
}

extern "C" int tmx_mbstd_fwd(tmx_handle_t h, const float* x, float* y, float* stat, int N, int H, int W, int C,
                             int C_total, int group_size, tmx_stream_t s) {
  TMX_REQUIRE(h && x && y && stat, TMX_ERR_ARG, "tmx_mbstd_fwd: NULL argument");
  TMX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C_total > C && group_size >= 1, TMX_ERR_SHAPE,
              "tmx_mbstd_fwd: bad shape N=%d H=%d W=%d C=%d C_total=%d group=%d", N, H, W, C, C_total, group_size);
  const int G = group_size < N ? group_size : N;
  TMX_REQUIRE(N % G == 0, TMX_ERR_SHAPE, "tmx_mbstd_fwd: batch %d is not a multiple of the group size %d", N, G);
  const int M = N / G;
  const long long per_sample = (long long)H * W * C;
  mbstd_stat_kernel<<<M, 256, 0, (cudaStream_t)s>>>(x, stat, G, M, per_sample);
  TMX_LAUNCHED(h, "mbstd_stat_kernel");
  const long long total = (long long)N * H * W * C_total;
  mbstd_concat_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(x, stat, y, total, C, C_total, M,
                                                                             (long long)H * W);
  TMX_LAUNCHED(h, "mbstd_concat_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- dense (D_patch head)
// networks.py:38-43 + 61-67 + 72-75: y[n][o] = act(wscale * sum_k x[n][k] w[k][o] + b[o]).
// Deterministic split-K: block (o-tile of 64, k-slice) accumulates a [NT samples x 64] partial tile into the
// workspace, a second kernel sums the slices in order and applies bias / leaky ReLU.
constexpr int kDenseNT = 32;   // samples per block
constexpr int kDenseKS = 256;  // k-slice length (32 x 257 floats of shared memory)
__global__ void __launch_bounds__(256) dense_partial_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            float* __restrict__ part, int N, int K, int Cout) {
  __shared__ float xs[kDenseNT][kDenseKS + 1];
  const int o = blockIdx.x * 64 + (threadIdx.x & 63);
  const int ty = threadIdx.x >> 6;                       // 0..3 -> samples ty*8 .. ty*8+7
  const int k0 = blockIdx.y * kDenseKS;
  const int n0 = blockIdx.z * kDenseNT;
  const int klen = min(kDenseKS, K - k0);
  for (int e = threadIdx.x; e < kDenseNT * kDenseKS; e += 256) {
    const int n = e / kDenseKS, k = e % kDenseKS;
    xs[n][k] = (n0 + n < N && k < klen) ? __ldg(x + (long long)(n0 + n) * K + k0 + k) : 0.f;
  }
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (o < Cout) {
    for (int k = 0; k < klen; ++k) {
      const float wv = __ldg(w + (long long)(k0 + k) * Cout + o);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(xs[ty * 8 + i][k], wv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + ty * 8 + i;
      if (n < N) part[((long long)blockIdx.y * N + n) * Cout + o] = acc[i];
    }
  }
}

// Fast path (Cout % 128 == 0, K % 128 == 0: the 8192 -> 512 head): the weight matrix is streamed ONCE with 128-bit
// loads.  Block = (128 outputs, 128-long k-slice, 32 samples); lane tx owns 4 consecutive outputs, warp ty owns
// samples ty*4 .. ty*4+3; x of the slice sits in shared memory (broadcast reads).  Same partial layout as above.
constexpr int kDenseKS4 = 64;    // short slices + 16 weight rows in flight per thread: the k loop is a chain of DRAM round trips
__global__ void __launch_bounds__(256) dense_partial4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             float* __restrict__ part, int N, int K, int Cout) {
  __shared__ float xs[kDenseNT][kDenseKS4 + 4];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = blockIdx.x * 128 + tx * 4;
  const int k0 = blockIdx.y * kDenseKS4;
  const int n0 = blockIdx.z * kDenseNT;
  for (int e = threadIdx.x; e < kDenseNT * (kDenseKS4 / 4); e += 256) {
    const int n = e / (kDenseKS4 / 4), k4 = (e % (kDenseKS4 / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + n < N) v = __ldg(reinterpret_cast<const float4*>(x + (long long)(n0 + n) * K + k0 + k4));
    *reinterpret_cast<float4*>(&xs[n][k4]) = v;
  }
  __syncthreads();
  float4 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* wp = reinterpret_cast<const float4*>(w + (long long)k0 * Cout + o);
  const long long wstride = Cout / 4;
#pragma unroll 16
  for (int k = 0; k < kDenseKS4; ++k) {
    const float4 wv = __ldg(wp + k * wstride);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float xv = xs[ty * 4 + i][k];
      acc[i].x = fmaf(xv, wv.x, acc[i].x);
      acc[i].y = fmaf(xv, wv.y, acc[i].y);
      acc[i].z = fmaf(xv, wv.z, acc[i].z);
      acc[i].w = fmaf(xv, wv.w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n < N) *reinterpret_cast<float4*>(part + ((long long)blockIdx.y * N + n) * Cout + o) = acc[i];
  }
}

// Streaming path of the 8192 -> 512 head (Cout % 512 == 0, K % 64 == 0).  The kernel above keeps only
// (rows in flight) x 512 B of DISTINCT weight bytes in flight per block - its eight warps all fetch the same rows - which
// is ~1 MB over the whole GPU: 27 us for 16.8 MB (0.07 of the HBM peak, ncu: DRAM 8 % busy).  Here a thread owns TWO
// outputs of 16 samples: 256 threads cover a whole 2-KB weight row (the block's two halves take 16 samples each), and
// the k loop is software-pipelined in batches of 16 rows (32 KB of distinct weights in flight per block).
// x of the slice sits in shared memory transposed ([k][sample]: one broadcast LDS.128 feeds 4 samples).
constexpr int kDenseKSS = 64;      // == kDenseKS4: same partial layout / workspace size
static_assert(kDenseKSS == kDenseKS4, "the streaming and the 128-bit dense kernels share one workspace layout");
__global__ void __launch_bounds__(512, 1) dense_partial_stream_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ w,
                                                                     float* __restrict__ part, int N, int K, int Cout) {
  // 512 threads: thread (half, tt) owns outputs 2*tt, 2*tt+1 of samples half*16 .. half*16+15 (32 accumulators).  With
  // all 32 samples per thread (256 threads, 64 accumulators, 255 registers) the eight resident warps of an SM issued
  // 1.2 instructions per cycle: the kernel was FFMA-issue bound at 22 us, not memory bound (ncu).
  __shared__ __align__(16) float xs[kDenseKSS][kDenseNT + 4];
  const int t = threadIdx.x, half = t >> 8, tt = t & 255;
  const int o = blockIdx.x * 512 + tt * 2;
  const int k0 = blockIdx.y * kDenseKSS;
  const int n0 = blockIdx.z * kDenseNT;
  const float2* wp = reinterpret_cast<const float2*>(w + (long long)k0 * Cout + o);
  const long long ws2 = Cout / 2;
  float2 wa[16], wb[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) wa[j] = __ldg(wp + j * ws2);          // first batch in flight while x is staged
  {
    const int e = t;                                                   // 32 samples x 16 float4 == 512 threads
    const int n = e / (kDenseKSS / 4), k4 = (e % (kDenseKSS / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + n < N) v = __ldg(reinterpret_cast<const float4*>(x + (long long)(n0 + n) * K + k0 + k4));
    xs[k4][n] = v.x;
    xs[k4 + 1][n] = v.y;
    xs[k4 + 2][n] = v.z;
    xs[k4 + 3][n] = v.w;
  }
  __syncthreads();
  constexpr int NH = kDenseNT / 2;
  float2 acc[NH];
#pragma unroll
  for (int n = 0; n < NH; ++n) acc[n] = make_float2(0.f, 0.f);
  auto fma_batch = [&](const float2 (&wv)[16], int kb) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int n4 = 0; n4 < NH / 4; ++n4) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[kb + j][half * NH + n4 * 4]);
        acc[4 * n4].x = fmaf(xv.x, wv[j].x, acc[4 * n4].x);
        acc[4 * n4].y = fmaf(xv.x, wv[j].y, acc[4 * n4].y);
        acc[4 * n4 + 1].x = fmaf(xv.y, wv[j].x, acc[4 * n4 + 1].x);
        acc[4 * n4 + 1].y = fmaf(xv.y, wv[j].y, acc[4 * n4 + 1].y);
        acc[4 * n4 + 2].x = fmaf(xv.z, wv[j].x, acc[4 * n4 + 2].x);
        acc[4 * n4 + 2].y = fmaf(xv.z, wv[j].y, acc[4 * n4 + 2].y);
        acc[4 * n4 + 3].x = fmaf(xv.w, wv[j].x, acc[4 * n4 + 3].x);
        acc[4 * n4 + 3].y = fmaf(xv.w, wv[j].y, acc[4 * n4 + 3].y);
      }
    }
  };
#pragma unroll
  for (int j = 0; j < 16; ++j) wb[j] = __ldg(wp + (16 + j) * ws2);
  fma_batch(wa, 0);
#pragma unroll
  for (int j = 0; j < 16; ++j) wa[j] = __ldg(wp + (32 + j) * ws2);
  fma_batch(wb, 16);
#pragma unroll
  for (int j = 0; j < 16; ++j) wb[j] = __ldg(wp + (48 + j) * ws2);
  fma_batch(wa, 32);
  fma_batch(wb, 48);
#pragma unroll
  for (int n = 0; n < NH; ++n) {
    const int nn = n0 + half * NH + n;
    if (nn < N) *reinterpret_cast<float2*>(part + ((long long)blockIdx.y * N + nn) * Cout + o) = acc[n];
  }
}

// y = act(wscale * sum_s part[s] + bias) with the slices spread over four thread groups (the one-thread-per-output
// kernel below walks 128 slices one after the other: 11 us of chained L2 round trips).
__global__ void __launch_bounds__(256) dense_finish4_kernel(const float* __restrict__ part, const float* __restrict__ bias,
                                                            float* __restrict__ y, int N, int Cout, int slices,
                                                            float wscale, int lrelu, float alpha) {
  __shared__ float4 red[3][64];
  const int q = threadIdx.x >> 6, e = threadIdx.x & 63;
  const long long total4 = (long long)N * Cout / 4;
  const long long t4 = (long long)blockIdx.x * 64 + e;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t4 < total4) {
    const float4* p = reinterpret_cast<const float4*>(part) + t4;
#pragma unroll 8
    for (int sl = q; sl < slices; sl += 4) {
      const float4 v = __ldg(p + (long long)sl * total4);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
  }
  if (q > 0) red[q - 1][e] = acc;
  __syncthreads();
  if (q == 0 && t4 < total4) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      acc.x += red[r][e].x;
      acc.y += red[r][e].y;
      acc.z += red[r][e].z;
      acc.w += red[r][e].w;
    }
    const int c = (int)((t4 * 4) % Cout);
    const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = make_float4(fmaf(acc.x, wscale, b.x), fmaf(acc.y, wscale, b.y), fmaf(acc.z, wscale, b.z),
                           fmaf(acc.w, wscale, b.w));
    if (lrelu) {
      v.x = fmaxf(v.x * alpha, v.x);
      v.y = fmaxf(v.y * alpha, v.y);
      v.z = fmaxf(v.z * alpha, v.z);
      v.w = fmaxf(v.w * alpha, v.w);
    }
    reinterpret_cast<float4*>(y)[t4] = v;
  }
}

static bool dense_fast(int K, int Cout) { return Cout % 128 == 0 && K % kDenseKS4 == 0; }

__global__ void __launch_bounds__(256) dense_finish_kernel(const float* __restrict__ part, const float* __restrict__ bias,
                                                           float* __restrict__ y, int N, int Cout, int slices,
                                                           float wscale, int lrelu, float alpha) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * Cout) return;
  float acc = 0.f;
  for (int s = 0; s < slices; ++s) acc += __ldg(part + (long long)s * N * Cout + t);
  float v = fmaf(acc, wscale, bias ? __ldg(bias + (int)(t % Cout)) : 0.f);
  if (lrelu) v = fmaxf(v * alpha, v);
  y[t] = v;
}

extern "C" int tmx_dense_workspace_bytes(int N, int K, int Cout, size_t* bytes) {
  TMX_REQUIRE(bytes && N > 0 && K > 0 && Cout > 0, TMX_ERR_ARG, "tmx_dense_workspace_bytes: bad argument");
  *bytes = (size_t)tmx_ceil_div(K, dense_fast(K, Cout) ? kDenseKS4 : kDenseKS) * N * Cout * sizeof(float);
  return TMX_OK;
}

extern "C" int tmx_dense_fwd(tmx_handle_t h, const float* x, const float* w, const float* bias, float wscale, float* y,
                             float* workspace, int N, int K, int Cout, int lrelu, float alpha, tmx_stream_t s) {
  TMX_REQUIRE(h && x && w && y && workspace, TMX_ERR_ARG, "tmx_dense_fwd: NULL argument");
  TMX_REQUIRE(N > 0 && K > 0 && Cout > 0, TMX_ERR_SHAPE, "tmx_dense_fwd: bad shape N=%d K=%d Cout=%d", N, K, Cout);
  const bool fast = dense_fast(K, Cout) && (((uintptr_t)x | (uintptr_t)w | (uintptr_t)workspace) & 15) == 0;
  const int slices = tmx_ceil_div(K, fast ? kDenseKS4 : kDenseKS);
  dim3 grid(tmx_ceil_div(Cout, fast ? 128 : 64), slices, tmx_ceil_div(N, kDenseNT));
  TMX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, TMX_ERR_SHAPE, "tmx_dense_fwd: problem too large");
  const bool stream = fast && Cout % 512 == 0 && !tmx_env_flag("TMX_DENSE_NO_STREAM");
  if (stream) {
    dim3 gs(Cout / 512, slices, tmx_ceil_div(N, kDenseNT));
    dense_partial_stream_kernel<<<gs, 512, 0, (cudaStream_t)s>>>(x, w, workspace, N, K, Cout);
  } else if (fast) {
    dense_partial4_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, w, workspace, N, K, Cout);
  } else {
    dense_partial_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, w, workspace, N, K, Cout);
  }
  TMX_LAUNCHED(h, "dense_partial_kernel");
  if (fast && (((uintptr_t)y | (uintptr_t)bias) & 15) == 0)
    dense_finish4_kernel<<<tmx_ceil_div((long long)N * Cout / 4, 64), 256, 0, (cudaStream_t)s>>>(
        workspace, bias, y, N, Cout, slices, wscale, lrelu, alpha);
  else
    dense_finish_kernel<<<tmx_ceil_div((long long)N * Cout, 256), 256, 0, (cudaStream_t)s>>>(
        workspace, bias, y, N, Cout, slices, wscale, lrelu, alpha);
  TMX_LAUNCHED(h, "dense_finish_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- Network.run output conversion (tfutil.py:649-659)
// y = saturate_cast(round(avg_pool_s(x * mul + add))) per image plane; the ops keep the reference's order and
// rounding: separate multiply and add (no fma contraction), window sum / count, tf.round = round-half-to-even,
// saturate_cast clamps to the integer range (NaN -> 0).  One thread per OUTPUT element; rows = N * C planes.
template <int OUT_U8>
__global__ void __launch_bounds__(256) convert_output_kernel(const float* __restrict__ x, void* __restrict__ y,
                                                             long long total, int H, int W, int shrink, float mul,
                                                             float add, int do_round) {
  const int Ho = H / shrink, Wo = W / shrink;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int xo = (int)(i % Wo);
    const int yo = (int)((i / Wo) % Ho);
    const long long plane = i / ((long long)Wo * Ho);
    const float* p = x + (plane * H + (long long)yo * shrink) * W + (long long)xo * shrink;
    float v;
    if (shrink == 1) {
      v = __fadd_rn(__fmul_rn(__ldg(p), mul), add);
    } else {
      float acc = 0.f;
      for (int a = 0; a < shrink; ++a)
        for (int b = 0; b < shrink; ++b) acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(__ldg(p + (long long)a * W + b), mul), add));
      v = __fdiv_rn(acc, (float)(shrink * shrink));
    }
    if (OUT_U8) {
      v = rintf(v);
      v = v != v ? 0.f : fminf(fmaxf(v, 0.f), 255.f);
      ((uint8_t*)y)[i] = (uint8_t)v;
    } else {
      ((float*)y)[i] = do_round ? rintf(v) : v;
    }
  }
}

extern "C" int tmx_convert_output(tmx_handle_t h, const float* x, void* y, int64_t planes, int H, int W, float mul,
                                  float add, int shrink, int out_kind, tmx_stream_t s) {
  TMX_REQUIRE(h && x && y, TMX_ERR_ARG, "tmx_convert_output: NULL argument");
  TMX_REQUIRE(planes > 0 && H > 0 && W > 0 && shrink >= 1 && H % shrink == 0 && W % shrink == 0, TMX_ERR_SHAPE,
              "tmx_convert_output: bad shape planes=%lld H=%d W=%d shrink=%d", (long long)planes, H, W, shrink);
  TMX_REQUIRE(out_kind >= 0 && out_kind <= 2, TMX_ERR_ARG, "tmx_convert_output: out_kind %d", out_kind);
  const long long total = planes * (H / shrink) * (W / shrink);
  const int grid = (int)(total / 256 + 1 < (long long)h->sm_count * 16 ? total / 256 + 1 : (long long)h->sm_count * 16);
  if (out_kind == 1)
    convert_output_kernel<1><<<grid, 256, 0, (cudaStream_t)s>>>(x, y, total, H, W, shrink, mul, add, 1);
  else
    convert_output_kernel<0><<<grid, 256, 0, (cudaStream_t)s>>>(x, y, total, H, W, shrink, mul, add, out_kind == 2);
  TMX_LAUNCHED(h, "convert_output_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- tanh over a flat fp32 tensor (networks.py:482-483)
// the image head's tanh when it cannot ride in the ToRGB epilogue (lod != 0: tanh follows the fade / upscale)
__global__ void __launch_bounds__(256) tanh_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = tanhf(__ldg(in + i));
}

extern "C" int tmx_tanh_f32(tmx_handle_t h, const float* in, float* out, int64_t n, tmx_stream_t s) {
  TMX_REQUIRE(h && in && out && n > 0, TMX_ERR_ARG, "tmx_tanh_f32: bad argument");
  const int grid = (int)(n / 256 + 1 < (long long)h->sm_count * 16 ? n / 256 + 1 : (long long)h->sm_count * 16);
  tanh_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(in, out, n);
  TMX_LAUNCHED(h, "tanh_kernel");
  return TMX_OK;
}

// adjoint of the trailing tanh: dx = dy * (1 - y^2), y = tanh(x)
__global__ void __launch_bounds__(256) tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                       float* __restrict__ dx, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float t = __ldg(y + i);
    dx[i] = __ldg(dy + i) * (1.f - t * t);
  }
}

extern "C" int tmx_tanh_bwd(tmx_handle_t h, const float* dy, const float* y, float* dx, int64_t n, tmx_stream_t s) {
  TMX_REQUIRE(h && dy && y && dx && n > 0, TMX_ERR_ARG, "tmx_tanh_bwd: bad argument");
  const int grid = (int)(n / 256 + 1 < (long long)h->sm_count * 16 ? n / 256 + 1 : (long long)h->sm_count * 16);
  tanh_bwd_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(dy, y, dx, n);
  TMX_LAUNCHED(h, "tanh_bwd_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- pixel_norm (networks.py:170-172)
// y[p][c] = x[p][c] * rsqrt(mean_c x[p][c]^2 + eps) on NHWC fp32; one warp per pixel, channels strided over lanes.
__global__ void __launch_bounds__(256) pixel_norm_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         long long npix, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < npix; p += nwarp) {
    const float* xp = x + p * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = __ldg(xp + c);
      acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float r = rsqrtf(acc / (float)C + eps);
    for (int c = lane; c < C; c += 32) y[p * C + c] = __ldg(xp + c) * r;
  }
}

// adjoint: r = rsqrt(mean_c x^2 + eps), dx_c = r * dy_c - r^3 / C * x_c * sum_c' dy_c' x_c'
__global__ void __launch_bounds__(256) pixel_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             float* __restrict__ dx, long long npix, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp; p < npix; p += nwarp) {
    const float* xp = x + p * C;
    const float* gp = dy + p * C;
    float sq = 0.f, dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = __ldg(xp + c);
      sq = fmaf(v, v, sq);
      dot = fmaf(v, __ldg(gp + c), dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    const float r = rsqrtf(sq / (float)C + eps);
    const float k = r * r * r * dot / (float)C;
    for (int c = lane; c < C; c += 32) dx[p * C + c] = r * __ldg(gp + c) - k * __ldg(xp + c);
  }
}

extern "C" int tmx_pixel_norm_bwd(tmx_handle_t h, const float* x, const float* dy, float* dx, int64_t npix, int C,
                                  float eps, tmx_stream_t s) {
  TMX_REQUIRE(h && x && dy && dx && npix > 0 && C > 0, TMX_ERR_ARG, "tmx_pixel_norm_bwd: bad argument");
  const long long blocks = (npix + 7) / 8;
  const int grid = (int)(blocks < (long long)h->sm_count * 16 ? blocks : (long long)h->sm_count * 16);
  pixel_norm_bwd_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, dy, dx, npix, C, eps);
  TMX_LAUNCHED(h, "pixel_norm_bwd_kernel");
  return TMX_OK;
}

extern "C" int tmx_pixel_norm(tmx_handle_t h, const float* x, float* y, int64_t npix, int C, float eps, tmx_stream_t s) {
  TMX_REQUIRE(h && x && y && npix > 0 && C > 0, TMX_ERR_ARG, "tmx_pixel_norm: bad argument");
  const long long blocks = (npix + 7) / 8;
  const int grid = (int)(blocks < (long long)h->sm_count * 16 ? blocks : (long long)h->sm_count * 16);
  pixel_norm_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, y, npix, C, eps);
  TMX_LAUNCHED(h, "pixel_norm_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- app-level matte compositing
// out[n][c][p] = sum_k src_k[n][c][p or 0] * w[k][p]  (util_scripts.py:1262,1268: np.sum(latents * weights, axis=0);
// :1337,1342: left * matt + right * (1 - matt)).  float64 products and sums rounded once (numpy promotes the float32
// latents to the float64 weights), or float32 throughout when the matte itself is float32 (math_f32).
__global__ void __launch_bounds__(256) weighted_sum_kernel(const float* const* __restrict__ srcs,
                                                           const int* __restrict__ bcast,
                                                           const double* __restrict__ w, float* __restrict__ out, int K,
                                                           long long planes, int HW, int math_f32) {
  const long long total = planes * HW;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int p = (int)(i % HW);
    const long long plane = i / HW;
    if (math_f32) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const float v = __ldg(srcs[k] + (bcast[k] ? plane : i));
        const float t = __fmul_rn(v, (float)__ldg(w + (long long)k * HW + p));
        acc = k == 0 ? t : __fadd_rn(acc, t);
      }
      out[i] = acc;
    } else {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) {
        const double v = (double)__ldg(srcs[k] + (bcast[k] ? plane : i));
        const double t = __dmul_rn(v, __ldg(w + (long long)k * HW + p));
        acc = k == 0 ? t : __dadd_rn(acc, t);
      }
      out[i] = (float)acc;
    }
  }
}

extern "C" int tmx_weighted_sum(tmx_handle_t h, const float* const* srcs, const int* bcast, const double* weights,
                                float* out, int K, int N, int C, int H, int W, int math_f32, tmx_stream_t s) {
  TMX_REQUIRE(h && srcs && bcast && weights && out, TMX_ERR_ARG, "tmx_weighted_sum: NULL argument");
  TMX_REQUIRE(K >= 1 && N > 0 && C > 0 && H > 0 && W > 0, TMX_ERR_SHAPE, "tmx_weighted_sum: bad shape K=%d N=%d C=%d %dx%d",
              K, N, C, H, W);
  const long long total = (long long)N * C * H * W;
  const int grid = (int)(total / 256 + 1 < (long long)h->sm_count * 16 ? total / 256 + 1 : (long long)h->sm_count * 16);
  weighted_sum_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(srcs, bcast, weights, out, K, (long long)N * C, H * W, math_f32);
  TMX_LAUNCHED(h, "weighted_sum_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- bias + leaky ReLU on an NHWC fp32 map
// act(apply_bias(.)) behind conv2d_downscale2d (networks.py:61-75, 142-148): with fused_scale the bias and the
// activation follow the (linear) conv + 2x2 average instead of preceding the pool.
__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                                       float* __restrict__ y, long long total, int C, int lrelu,
                                                       float alpha) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float v = __ldg(x + i) + (bias ? __ldg(bias + (int)(i % C)) : 0.f);
    y[i] = lrelu ? fmaxf(v * alpha, v) : v;
  }
}

extern "C" int tmx_bias_act(tmx_handle_t h, const float* x, const float* bias, float* y, int64_t npix, int C, int lrelu,
                            float alpha, tmx_stream_t s) {
  TMX_REQUIRE(h && x && y && npix > 0 && C > 0, TMX_ERR_ARG, "tmx_bias_act: bad argument");
  const long long total = npix * C;
  const int grid = (int)(total / 256 + 1 < (long long)h->sm_count * 16 ? total / 256 + 1 : (long long)h->sm_count * 16);
  bias_act_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(x, bias, y, total, C, lrelu, alpha);
  TMX_LAUNCHED(h, "bias_act_kernel");
  return TMX_OK;
}


// ---------------------------------------------------------------- window copy / embed with device-side offsets
// The crop-aware train step (loss.crop_window / mid_window / tail_window, loss.py:78-90 random_crop) slices fixed-SIZE
// windows at offsets that change every step.  Reading the offsets from device memory keeps every launch of the step
// identical from one step to the next, so the whole step can be replayed as CUDA graphs (train.GraphedStep).
// Tensors are [A][H][W][B] fp32: NCHW -> A = N*C, B = 1; NHWC -> A = N, B = C.
//   embed == 0: dst[A][wh][ww][B] = src[A][oy + y][ox + x][B]
//   embed == 1: dst[A][H][W][B]   = src[A][y - oy][x - ox][B] inside the window, 0 elsewhere (adjoint of the slice)
// One thread = 4 consecutive floats of a DESTINATION row (rows are W*B or ww*B floats long, a multiple of 4: the host
// checks), stored as one float4; the source side is read with scalar loads (an NCHW window starts at any x) or one
// float4 (NHWC: pixels are >= 16 B).
__global__ void __launch_bounds__(256) window_copy_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                          long long quads, int H, int W, int B, int wh, int ww, int oy_h,
                                                          int ox_h, const int32_t* __restrict__ off_dev, int embed,
                                                          int vec) {
  const int oy = off_dev ? __ldg(off_dev) : oy_h, ox = off_dev ? __ldg(off_dev + 1) : ox_h;
  const int rows_h = embed ? H : wh;                       // destination rows per slab a
  const int q_per_row = (embed ? W : ww) * B / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
    const long long row = q / q_per_row;
    const int e0 = (int)(q - row * q_per_row) * 4;         // first float of the quad inside its row
    const int a = (int)(row / rows_h), y = (int)(row - (long long)a * rows_h);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!embed) {
      const float* sp = src + (((long long)a * H + oy + y) * W + ox) * B + e0;
      if (vec) v = __ldg(reinterpret_cast<const float4*>(sp));
      else v = make_float4(__ldg(sp), __ldg(sp + 1), __ldg(sp + 2), __ldg(sp + 3));
    } else if (y >= oy && y < oy + wh) {
      const float* sp = src + (((long long)a * wh + (y - oy)) * ww - ox) * B + e0;   // sp[i] valid inside the window
      const int lo = ox * B, hi = (ox + ww) * B;
      if (vec) {
        if (e0 >= lo && e0 < hi) v = __ldg(reinterpret_cast<const float4*>(sp));
      } else {
        v.x = (e0 >= lo && e0 < hi) ? __ldg(sp) : 0.f;
        v.y = (e0 + 1 >= lo && e0 + 1 < hi) ? __ldg(sp + 1) : 0.f;
        v.z = (e0 + 2 >= lo && e0 + 2 < hi) ? __ldg(sp + 2) : 0.f;
        v.w = (e0 + 3 >= lo && e0 + 3 < hi) ? __ldg(sp + 3) : 0.f;
      }
    }
    reinterpret_cast<float4*>(dst)[q] = v;
  }
}

extern "C" int tmx_window_copy(tmx_handle_t h, const float* src, float* dst, int64_t A, int H, int W, int B, int wh,
                               int ww, int oy, int ox, const int32_t* off_dev, int embed, tmx_stream_t s) {
  TMX_REQUIRE(h && src && dst, TMX_ERR_ARG, "tmx_window_copy: NULL argument");
  TMX_REQUIRE(A > 0 && H > 0 && W > 0 && B > 0 && wh > 0 && ww > 0 && wh <= H && ww <= W, TMX_ERR_SHAPE,
              "tmx_window_copy: bad shape A=%lld H=%d W=%d B=%d window %dx%d", (long long)A, H, W, B, wh, ww);
  TMX_REQUIRE(off_dev != nullptr || (oy >= 0 && ox >= 0 && oy + wh <= H && ox + ww <= W), TMX_ERR_SHAPE,
              "tmx_window_copy: window (%d,%d)+%dx%d outside %dx%d", oy, ox, wh, ww, H, W);
  const long long row_len = (long long)(embed ? W : ww) * B;
  TMX_REQUIRE(row_len % 4 == 0, TMX_ERR_SHAPE, "tmx_window_copy: destination rows of %lld floats (need a multiple of 4)",
              row_len);
  const long long quads = A * (embed ? H : wh) * row_len / 4;
  const int vec = (B % 4 == 0) ? 1 : 0;          // NHWC pixels are 16-B aligned at any offset; NCHW windows are not
  long long blocks = (quads + 255) / 256;
  const long long cap = (long long)h->sm_count * 32;
  if (blocks > cap) blocks = cap;
  window_copy_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(src, dst, quads, H, W, B, wh, ww, oy, ox, off_dev,
                                                                   embed, vec);
  TMX_LAUNCHED(h, "window_copy_kernel");
  return TMX_OK;
}
