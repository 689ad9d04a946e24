// conv_ffma.cu — CUDA-core fp32 implicit-GEMM convolution (k = 1 or 3, stride 1,
// REFLECT pad) over NHWC activations, with the TextureMixer epilogue fused:
//   y = [residual +] lrelu(wscale * acc + bias)            (networks.py:48-75, :437)
// and optional nearest-neighbour x2 upsampling of the input folded into the
// gather (networks.py:80-88 never materialised).
//
// Role: exact-fp32 path for the layers whose contraction is too thin for the
// tensor-core kernel (Cin*k*k < 576) or that are HBM-bound anyway, and the
// device-side cross-check of conv_tc.cu.  GEMM view: M = N*H*W pixels,
// N = Cout, K = k*k*Cin ordered (tap, channel) - exactly the HWIO weight
// layout flattened, so B is read row-major with Cout contiguous.
//
// Tile: BM x BN outputs per 256-thread CTA (BM*BN = 4096), 4x4 per thread,
// BK = 16 channels of one tap per step, register-prefetch + double-buffered smem.
#include "common.cuh"

struct ConvFfmaParams {
  const float* x;
  const float* w;
  const float* bias;
  const float* residual;
  float* y;
  int N, H, W, Hin, Win, Cin, Cout, k;
  long long M;
  int up2, lrelu, has_res;
  float wscale, alpha;
};

template <int BM, int BN>
__global__ void __launch_bounds__(256) conv_ffma_kernel(const ConvFfmaParams p) {
  constexpr int BK = 16;
  constexpr int AS = BM + 4;      // padded row stride of the transposed A tile
  constexpr int NA = BM / 64;     // float4 A loads per thread per step
  constexpr int TXN = BN / 4;     // threads along N
  __shared__ __align__(16) float As[2][BK][AS];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int pad = p.k >> 1;
  const int cchunks = p.Cin / BK;
  const int KT = p.k * p.k * cchunks;

  // A-load role: pixel (t/4 + i*64), channel quarter t%4
  const int aq = t & 3;
  int an[NA], ay[NA], ax[NA];
  bool av[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    long long m = m0 + (t >> 2) + i * 64;
    av[i] = m < p.M;
    long long mm = av[i] ? m : 0;
    ax[i] = (int)(mm % p.W);
    long long q = mm / p.W;
    ay[i] = (int)(q % p.H);
    an[i] = (int)(q / p.H);
  }
  // B-load role: row t / TXN (k within the step), float4 column t % TXN
  const int bk = t / TXN, bq = t % TXN;
  const bool bv = bk < BK && (n0 + bq * 4) < p.Cout;

  float4 ra[NA];
  float4 rb;

  auto load_regs = [&](int kt) {
    const int tap = kt / cchunks;
    const int c0 = (kt - tap * cchunks) * BK;
    const int u = tap / p.k, v = tap - u * p.k;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (av[i]) {
        int sy = tmx_reflect(ay[i] + u - pad, p.H);
        int sx = tmx_reflect(ax[i] + v - pad, p.W);
        if (p.up2) {
          sy >>= 1;
          sx >>= 1;
        }
        const float* src = p.x + (((long long)an[i] * p.Hin + sy) * p.Win + sx) * p.Cin + c0 + aq * 4;
        ra[i] = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (bv) {
      const float* src = p.w + ((long long)(tap * p.Cin + c0 + bk)) * p.Cout + n0 + bq * 4;
      rb = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int pl = (t >> 2) + i * 64;
      As[buf][aq * 4 + 0][pl] = ra[i].x;
      As[buf][aq * 4 + 1][pl] = ra[i].y;
      As[buf][aq * 4 + 2][pl] = ra[i].z;
      As[buf][aq * 4 + 3][pl] = ra[i].w;
    }
    if (bk < BK) *reinterpret_cast<float4*>(&Bs[buf][bk][bq * 4]) = rb;
  };

  const int ty = t / TXN, tx = t % TXN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_regs(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_regs(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av4[4] = {a.x, a.y, a.z, a.w};
      const float bv4[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av4[i], bv4[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_smem(buf ^ 1);
    __syncthreads();
  }

  const int col = n0 + tx * 4;
  if (col >= p.Cout) return;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  const float b4[4] = {bias4.x, bias4.y, bias4.z, bias4.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = fmaf(acc[i][j], p.wscale, b4[j]);
      if (p.lrelu) v = fmaxf(v * p.alpha, v);
      o[j] = v;
    }
    if (p.has_res) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + m * p.Cout + col));
      o[0] += r.x;
      o[1] += r.y;
      o[2] += r.z;
      o[3] += r.w;
    }
    *reinterpret_cast<float4*>(p.y + m * p.Cout + col) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int tmx_conv2d_fwd_ffma(tmx_handle_t h, const tmx_conv_desc_t* d, const tmx_conv_io_t* io, cudaStream_t st) {
  TMX_REQUIRE(io->x_f32 && io->w && io->y_f32, TMX_ERR_ARG,
              "tmx_conv2d_fwd[FFMA]: x_f32, w and y_f32 are required");
  TMX_REQUIRE(d->Cin % 16 == 0 && d->Cout % 4 == 0, TMX_ERR_SHAPE,
              "tmx_conv2d_fwd[FFMA]: Cin=%d must be a multiple of 16 and Cout=%d of 4", d->Cin, d->Cout);
  const bool up2 = (d->flags & TMX_CONV_UP2_IN) != 0;
  TMX_REQUIRE(!up2 || (d->H % 2 == 0 && d->W % 2 == 0), TMX_ERR_SHAPE,
              "tmx_conv2d_fwd[FFMA]: UP2_IN needs even H, W (got %d x %d)", d->H, d->W);
  TMX_REQUIRE(!(d->flags & TMX_CONV_RESIDUAL) || io->residual, TMX_ERR_ARG,
              "tmx_conv2d_fwd[FFMA]: RESIDUAL flag without residual pointer");
  TMX_REQUIRE(!(d->flags & (TMX_CONV_UP2_OUT | TMX_CONV_HALO_REPLICATE | TMX_CONV_HALO_ZERO | TMX_CONV_TORGB | TMX_CONV_W_PER_SAMPLE)) && !io->y_hi && !io->y_lo,
              TMX_ERR_UNSUPPORTED,
              "tmx_conv2d_fwd[FFMA]: split-plane / ToRGB outputs are produced by the TC path (or tmx_split_halo_pack)");
  ConvFfmaParams p;
  p.x = io->x_f32;
  p.w = io->w;
  p.bias = io->bias;
  p.residual = io->residual;
  p.y = io->y_f32;
  p.N = d->N;
  p.H = d->H;
  p.W = d->W;
  p.Hin = up2 ? d->H / 2 : d->H;
  p.Win = up2 ? d->W / 2 : d->W;
  p.Cin = d->Cin;
  p.Cout = d->Cout;
  p.k = d->k;
  p.M = (long long)d->N * d->H * d->W;
  p.up2 = up2;
  p.lrelu = (d->flags & TMX_CONV_LRELU) != 0;
  p.has_res = (d->flags & TMX_CONV_RESIDUAL) != 0;
  p.wscale = d->wscale;
  p.alpha = d->lrelu_alpha;
  if (d->Cout <= 16) {
    dim3 grid(tmx_ceil_div(p.M, 256), tmx_ceil_div(d->Cout, 16));
    conv_ffma_kernel<256, 16><<<grid, 256, 0, st>>>(p);
  } else if (d->Cout <= 32) {
    dim3 grid(tmx_ceil_div(p.M, 128), tmx_ceil_div(d->Cout, 32));
    conv_ffma_kernel<128, 32><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(tmx_ceil_div(p.M, 64), tmx_ceil_div(d->Cout, 64));
    conv_ffma_kernel<64, 64><<<grid, 256, 0, st>>>(p);
  }
  TMX_LAUNCHED(h, "conv_ffma_kernel");
  return TMX_OK;
}
