// backward.cu — HBM-bound glue of the backward pass (K11 in SURVEY §2.2):
//
//  tmx_grad_prepare   turns "gradient w.r.t. a layer output y" into the operand of the next data-/weight-
//                     gradient GEMMs in ONE pass: adjoint of the padding (fold the ring of the LIN-mode
//                     dgrad output back onto the interior: REFLECT of networks.py:55, or REPLICATE for the
//                     sub-pixel upsample form), optional second addend (residual branch, networks.py:437),
//                     adjoint of downscale2d (networks.py:131-136: x0.25 broadcast), leaky-ReLU mask from the
//                     saved forward output (networks.py:72-75), bias gradient (per-channel sum, apply_bias
//                     networks.py:61-67), re-split into bf16 hi/lo planes on the zero-ringed grid
//                     [N][H+4][W+4][C] the dgrad / wgrad kernels read (optionally phase-packed at half
//                     resolution for a sub-pixel upsample conv), and/or plain fp32 NHWC.
//  tmx_conv_weights_transpose  prepared forward weight planes [Nrows][taps*K] -> data-gradient planes
//                     [K][taps*Nrows] with the taps flipped (no re-rounding: bit-identical hi/lo values).
#include "common.cuh"

struct GradPrepParams {
  tmx_grad_desc_t d;
  tmx_grad_io_t io;
  int C8;          // channel groups of 8
  int ppb;         // pixel lanes per block
  int iters;       // pixels per lane
};

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// One thread = 8 channels of one position of the OUTPUT grid (zero ring included, so the ring gets zeroed).
__global__ void __launch_bounds__(256) grad_prepare_kernel(const GradPrepParams P) {
  extern __shared__ float red[];   // [ppb][C8*8] bias-gradient partials
  const tmx_grad_desc_t& d = P.d;
  const int cg = threadIdx.x % P.C8;
  const int lane = threadIdx.x / P.C8;
  const int H = d.H, W = d.W, C = d.C;
  const int Hq = H + 4, Wq = W + 4;
  const long long total = (long long)d.N * Hq * Wq;
  float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int it = 0; it < P.iters; ++it) {
    const long long pos = ((long long)blockIdx.x * P.iters + it) * P.ppb + lane;
    if (pos >= total) break;
    const int c = (int)(pos % Wq) - 2;
    const long long q = pos / Wq;
    const int r = (int)(q % Hq) - 2;
    const int n = (int)(q / Hq);
    const bool interior = r >= 0 && r < H && c >= 0 && c < W;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (interior) {
      if (d.src_kind == 0) {
        // rows / cols of the grid whose values land on (r, c): itself plus the folded ring
        int rs[2], cs[2], nr = 0, nc = 0;
        rs[nr++] = r + 2;
        cs[nc++] = c + 2;
        if (d.fold == 0) {         // REFLECT adjoint: padded row -1 -> row 1, padded row H -> row H-2
          if (r == 1) rs[nr++] = 1;
          if (r == H - 2) rs[nr++] = H + 2;
          if (c == 1) cs[nc++] = 1;
          if (c == W - 2) cs[nc++] = W + 2;
        } else if (d.fold == 1) {  // REPLICATE adjoint: padded row -1 -> row 0, padded row H -> row H-1
          if (r == 0) rs[nr++] = 1;
          if (r == H - 1) rs[nr++] = H + 2;
          if (c == 0) cs[nc++] = 1;
          if (c == W - 1) cs[nc++] = W + 2;
        }
        // H == 2 (or W == 2) with REFLECT: row 1 receives padded -1 AND row 0 receives padded H; both handled above
        if (d.fold == 0 && H == 2 && nr == 1) {   // r == 0 == H-2 was caught; r == 1 caught too
        }
        for (int a = 0; a < nr; ++a)
          for (int b = 0; b < nc; ++b) {
            float t[8];
            ld8(P.io.g + (((long long)n * Hq + rs[a]) * Wq + cs[b]) * C + cg * 8, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += t[j];
          }
      } else if (d.src_kind == 1) {
        ld8(P.io.g + (((long long)n * H + r) * W + c) * C + cg * 8, v);
      } else {                     // avg-pool adjoint: every pixel of a 2x2 window receives a quarter
        ld8(P.io.g + (((long long)n * (H / 2) + (r >> 1)) * (W / 2) + (c >> 1)) * C + cg * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
      }
      if (P.io.add != nullptr) {
        float t[8];
        ld8(P.io.add + (((long long)n * H + r) * W + c) * C + cg * 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
      }
      if (d.mask_kind == 1) {
        float y[8];
        ld8(reinterpret_cast<const float*>(P.io.y_mask) + (((long long)n * H + r) * W + c) * C + cg * 8, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= (y[j] > 0.f) ? 1.f : d.alpha;
      } else if (d.mask_kind == 2) {
        const uint4 yb = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(P.io.y_mask) +
                                                             (((long long)n * (H + 2) + r + 1) * (W + 2) + c + 1) * C +
                                                             cg * 8));
        const uint32_t w4[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t hb = (w4[j >> 1] >> ((j & 1) * 16)) & 0xffffu;     // bf16 bits of hi(y)
          const bool pos_y = hb != 0u && (hb & 0x8000u) == 0u;
          v[j] *= pos_y ? 1.f : d.alpha;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) bsum[j] += v[j];
      if (P.io.dz_f32 != nullptr) {
        float4* o = reinterpret_cast<float4*>(P.io.dz_f32 + (((long long)n * H + r) * W + c) * C + cg * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    if (P.io.dz_hi != nullptr) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t h0, l0, h1, l1;
        tmx_split_bf16(v[2 * j], h0, l0);
        tmx_split_bf16(v[2 * j + 1], h1, l1);
        ph[j] = h0 | (h1 << 16);
        pl[j] = l0 | (l1 << 16);
      }
      long long o;
      if (!d.phase_pack) {
        o = pos * C + cg * 8;
      } else {
        // half-resolution grid [N][H/2+4][W/2+4][4C]; position (r,c) -> low-res (r>>1, c>>1), phase (r&1, c&1).
        // Ring positions of the full-res grid have no image there; the low-res ring is zeroed by the host.
        if (!interior) continue;
        const int Hl = H / 2 + 4, Wl = W / 2 + 4;
        const int phs = ((r & 1) << 1) | (c & 1);
        o = ((((long long)n * Hl + (r >> 1) + 2) * Wl + (c >> 1) + 2) * 4 + phs) * C + cg * 8;
      }
      *reinterpret_cast<uint4*>(P.io.dz_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(P.io.dz_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
  if (P.io.dbias != nullptr) {
    float* mine = red + (lane * P.C8 + cg) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = bsum[j];
    __syncthreads();
    if (lane == 0) {
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
        for (int l = 0; l < P.ppb; ++l) s += red[(l * P.C8 + cg) * 8 + j];
        atomicAdd(P.io.dbias + cg * 8 + j, s * d.dbias_scale);
      }
    }
  }
}

extern "C" int tmx_grad_prepare(tmx_handle_t h, const tmx_grad_desc_t* d, const tmx_grad_io_t* io, tmx_stream_t s) {
  TMX_REQUIRE(h && d && io && io->g, TMX_ERR_ARG, "tmx_grad_prepare: NULL argument");
  TMX_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->C % 8 == 0 && d->C <= 2048, TMX_ERR_SHAPE,
              "tmx_grad_prepare: bad shape N=%d H=%d W=%d C=%d (C %% 8 == 0, C <= 2048)", d->N, d->H, d->W, d->C);
  TMX_REQUIRE(d->src_kind >= 0 && d->src_kind <= 2 && d->fold >= 0 && d->fold <= 2 && d->mask_kind >= 0 &&
                  d->mask_kind <= 2, TMX_ERR_ARG, "tmx_grad_prepare: bad enum");
  TMX_REQUIRE(d->mask_kind == 0 || io->y_mask, TMX_ERR_ARG, "tmx_grad_prepare: mask requested without y_mask");
  TMX_REQUIRE(d->src_kind != 2 || (d->H % 2 == 0 && d->W % 2 == 0), TMX_ERR_SHAPE,
              "tmx_grad_prepare: pool adjoint needs even H, W");
  TMX_REQUIRE(d->src_kind != 0 || d->fold == 2 || (d->H >= 2 && d->W >= 2), TMX_ERR_SHAPE,
              "tmx_grad_prepare: padding adjoint needs H, W >= 2");
  TMX_REQUIRE(!d->phase_pack || (d->H % 2 == 0 && d->W % 2 == 0 && io->dz_hi), TMX_ERR_SHAPE,
              "tmx_grad_prepare: phase_pack needs even H, W and plane outputs");
  TMX_REQUIRE((io->dz_hi == nullptr) == (io->dz_lo == nullptr), TMX_ERR_ARG, "tmx_grad_prepare: dz_hi/dz_lo go together");
  TMX_REQUIRE(io->dz_hi || io->dz_f32 || io->dbias, TMX_ERR_ARG, "tmx_grad_prepare: no output");
  GradPrepParams P;
  P.d = *d;
  P.io = *io;
  P.C8 = d->C / 8;
  P.ppb = 256 / P.C8 > 0 ? 256 / P.C8 : 1;
  P.iters = 16;
  const int threads = P.ppb * P.C8;
  const long long total = (long long)d->N * (d->H + 4) * (d->W + 4);
  const long long per_block = (long long)P.ppb * P.iters;
  const size_t smem = io->dbias ? (size_t)threads * 8 * sizeof(float) : 0;
  TMX_REQUIRE(threads <= 256 && smem <= 48 * 1024, TMX_ERR_SHAPE, "tmx_grad_prepare: C=%d not supported", d->C);
  grad_prepare_kernel<<<tmx_ceil_div(total, per_block), threads, smem, (cudaStream_t)s>>>(P);
  TMX_LAUNCHED(h, "grad_prepare_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- weight planes for the data gradient
__global__ void __launch_bounds__(256) weights_transpose_kernel(const uint16_t* __restrict__ in_hi,
                                                                const uint16_t* __restrict__ in_lo,
                                                                uint16_t* __restrict__ out_hi,
                                                                uint16_t* __restrict__ out_lo, int rows, int taps,
                                                                int K) {
  // out[k][(taps-1-t)*rows + n] = in[n][t*K + k]; 32x32 smem tile per tap
  __shared__ uint16_t th[32][34], tl[32][34];
  const int t = blockIdx.z;
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + i * 8, k = k0 + tx;
    if (n < rows && k < K) {
      const long long src = (long long)n * taps * K + (long long)t * K + k;
      th[ty + i * 8][tx] = in_hi[src];
      tl[ty + i * 8][tx] = in_lo[src];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty + i * 8, n = n0 + tx;
    if (n < rows && k < K) {
      const long long dst = (long long)k * taps * rows + (long long)(taps - 1 - t) * rows + n;
      out_hi[dst] = th[tx][ty + i * 8];
      out_lo[dst] = tl[tx][ty + i * 8];
    }
  }
}

extern "C" int tmx_conv_weights_transpose(tmx_handle_t h, const uint16_t* w_hi, const uint16_t* w_lo, int rows, int taps,
                                          int K, uint16_t* wt_hi, uint16_t* wt_lo, tmx_stream_t s) {
  TMX_REQUIRE(h && w_hi && w_lo && wt_hi && wt_lo, TMX_ERR_ARG, "tmx_conv_weights_transpose: NULL argument");
  TMX_REQUIRE(rows > 0 && K > 0 && (taps == 1 || taps == 9), TMX_ERR_SHAPE,
              "tmx_conv_weights_transpose: bad shape rows=%d taps=%d K=%d", rows, taps, K);
  dim3 grid(tmx_ceil_div(K, 32), tmx_ceil_div(rows, 32), taps);
  weights_transpose_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(w_hi, w_lo, wt_hi, wt_lo, rows, taps, K);
  TMX_LAUNCHED(h, "weights_transpose_kernel");
  return TMX_OK;
}

int tmx_conv2d_dgrad_tc(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                        const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                        cudaStream_t st);

extern "C" int tmx_conv2d_dgrad(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                                const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                                tmx_stream_t s) {
  TMX_REQUIRE(h != nullptr, TMX_ERR_ARG, "tmx_conv2d_dgrad: NULL handle");
  const void* ptrs[] = {dz_hi, dz_lo, wt_hi, wt_lo, g_f32};
  for (const void* q : ptrs)
    TMX_REQUIRE(((uintptr_t)q & 15) == 0, TMX_ERR_ARG, "tmx_conv2d_dgrad: buffers must be 16-byte aligned (%p)", q);
  return tmx_conv2d_dgrad_tc(h, N, H, W, Cin, Cout, k, dz_hi, dz_lo, wt_hi, wt_lo, g_f32, (cudaStream_t)s);
}
