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
  int rpb;         // grid rows (n, r) per block
  int c8_shift;    // log2(C8) when C8 is a power of two (every layer of the path but the 513 -> 576 padded one), else -1
};

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// One thread = 8 channels (fixed for the thread: blockDim is a multiple of C8, so its bias-gradient partial stays in
// registers) of the units t, t + grid size, ... of the OUTPUT grid [N][H+4][W+4] (zero ring included, so the ring
// gets zeroed); a unit = (grid row, column, channel group), flattened over the whole grid so that narrow-channel
// maps (C = 16: two groups per pixel) keep every lane busy and every CTA gets the same share.  Index arithmetic is 32-bit: the per-element 64-bit
// divisions of the first version cost as much issue time as the memory traffic took.
__global__ void __launch_bounds__(256, 4) grad_prepare_kernel(const GradPrepParams P) {
  extern __shared__ float red[];   // [blockDim / C8][C] bias-gradient partials
  tmx_pdl_trigger();
  const tmx_grad_desc_t& d = P.d;
  const int C8 = P.C8;
  const int cg = threadIdx.x % C8;
  const int H = d.H, W = d.W, C = d.C;
  const int Hq = H + 4, Wq = W + 4;
  const int rows_total = d.N * Hq;
  const int co = cg * 8;
  const int units_per_row = Wq * C8;
  const int units = rows_total * units_per_row;            // < 2^31 (host check)
  const int ustride = gridDim.x * blockDim.x;              // a multiple of C8: u % C8 == cg for every unit of a thread
  const float* __restrict__ gsrc = P.io.g;
  const float* __restrict__ addp = P.io.add;
  float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  // index walk without divisions: u advances by the constant ustride = d_row rows + d_rem units, a row by row/Hq
  const int u0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int d_row = ustride / units_per_row, d_rem = ustride - d_row * units_per_row;
  int row = u0 / units_per_row, rem = u0 - row * units_per_row;
  const int d_n = d_row / Hq, d_rr = d_row - d_n * Hq;
  int n = row / Hq, rr = row - n * Hq;
  tmx_pdl_wait();      // (the index set-up above overlaps the previous kernel's tail)
  for (int u = u0; u < units; u += ustride) {
    const int cq = P.c8_shift >= 0 ? (rem >> P.c8_shift) : rem / C8;
    const int r = rr - 2;
    const int c = cq - 2;
    const bool interior = r >= 0 && r < H && c >= 0 && c < W;
    const int row_now = row, n_now = n;
    // advance to the next unit of this thread
    rem += d_rem;
    row += d_row;
    rr += d_rr;
    n += d_n;
    if (rem >= units_per_row) {
      rem -= units_per_row;
      ++row;
      ++rr;
    }
    if (rr >= Hq) {
      rr -= Hq;
      ++n;
    }
    if (rr >= Hq) {          // (d_rr + carry can pass Hq twice only when d_rr == Hq - 1 and both carries hit)
      rr -= Hq;
      ++n;
    }
    {
    const int row = row_now, n = n_now;       // (shadow the walkers: the body below works on the CURRENT unit)
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (interior) {
      const long long in_pix = ((long long)n * H + r) * W + c;                 // NHWC pixel index
      if (d.src_kind == 0) {
        // rows / cols of the dgrad grid whose values land on (r, c): itself plus the folded ring
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
        for (int a = 0; a < nr; ++a)
          for (int b = 0; b < nc; ++b) {
            float t[8];
            ld8(gsrc + (((long long)n * Hq + rs[a]) * Wq + cs[b]) * C + co, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += t[j];
          }
      } else if (d.src_kind == 1) {
        ld8(gsrc + in_pix * C + co, v);
      } else {                     // avg-pool adjoint: every pixel of a 2x2 window receives a quarter
        ld8(gsrc + (((long long)n * (H / 2) + (r >> 1)) * (W / 2) + (c >> 1)) * C + co, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
      }
      if (addp != nullptr) {
        float t[8];
        ld8(addp + in_pix * C + co, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
      }
      if (d.mask_kind == 1) {
        float y[8];
        ld8(reinterpret_cast<const float*>(P.io.y_mask) + in_pix * C + co, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= (y[j] > 0.f) ? 1.f : d.alpha;
      } else if (d.mask_kind == 2) {
        const uint4 yb = __ldg(reinterpret_cast<const uint4*>(
            reinterpret_cast<const uint16_t*>(P.io.y_mask) + (((long long)n * (H + 2) + r + 1) * (W + 2) + c + 1) * C + co));
        const uint32_t w4[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // y > 0  <=>  the bf16 word, moved to the top half, is a positive int32
          v[2 * j] *= ((int)(w4[j] << 16) > 0) ? 1.f : d.alpha;
          v[2 * j + 1] *= ((int)(w4[j] & 0xffff0000u) > 0) ? 1.f : d.alpha;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) bsum[j] += v[j];
      if (P.io.dz_f32 != nullptr) {
        float4* o = reinterpret_cast<float4*>(P.io.dz_f32 + in_pix * C + co);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    if (P.io.dz_hi != nullptr) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tmx_split_bf16x2(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
      }
      long long o;
      if (!d.phase_pack) {
        o = ((long long)row * Wq + cq) * C + co;
      } else {
        // half-resolution grid [N][H/2+4][W/2+4][4C]; position (r,c) -> low-res (r>>1, c>>1), phase (r&1, c&1).
        // Ring positions of the full-res grid have no image there; the low-res ring is zeroed by the host.
        if (!interior) continue;
        const int Hl = H / 2 + 4, Wl = W / 2 + 4;
        const int phs = ((r & 1) << 1) | (c & 1);
        o = ((((long long)n * Hl + (r >> 1) + 2) * Wl + (c >> 1) + 2) * 4 + phs) * C + co;
      }
      *reinterpret_cast<uint4*>(P.io.dz_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(P.io.dz_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
    }
  }
  if (P.io.dbias != nullptr) {
    const int lanes = blockDim.x / C8;
    float* mine = red + (threadIdx.x / C8) * C + co;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = bsum[j];
    __syncthreads();
    // all threads reduce: thread t sums channel t of the block's lanes
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * C + ch];
      atomicAdd(P.io.dbias + ch, s * d.dbias_scale);
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
  const long long rows = (long long)d->N * (d->H + 4);
  TMX_REQUIRE(rows < (1ll << 31), TMX_ERR_SHAPE, "tmx_grad_prepare: too many grid rows");
  GradPrepParams P;
  P.d = *d;
  P.io = *io;
  P.C8 = d->C / 8;
  P.ppb = 256 / P.C8 > 0 ? 256 / P.C8 : 1;      // lanes: threads = lanes * C8 (a multiple of C8, <= 256)
  const int threads = P.ppb * P.C8;
  const long long units = rows * (long long)(d->W + 4) * P.C8;
  TMX_REQUIRE(units < (1ll << 31) - (1 << 20), TMX_ERR_SHAPE, "tmx_grad_prepare: grid too large");
  // persistent-style grid: 4 resident CTAs per SM (64 registers per thread), fewer when the map is small.  (Two units
  // in flight per thread at 3 CTAs per SM was tried: 95 -> 113 us for the 64x64 c256 case - fewer threads, not more bytes)
  long long blocks = (units + threads - 1) / threads;
  if (blocks > 4LL * h->sm_count) blocks = 4LL * h->sm_count;
  P.rpb = 0;
  P.c8_shift = -1;
  for (int b = 0; b < 12; ++b)
    if ((1 << b) == P.C8) P.c8_shift = b;
  const size_t smem = io->dbias ? (size_t)threads * 8 * sizeof(float) : 0;
  TMX_REQUIRE(threads <= 256 && smem <= 48 * 1024, TMX_ERR_SHAPE, "tmx_grad_prepare: C=%d not supported", d->C);
  TMX_CUDA(tmx_launch_pdl(grad_prepare_kernel, dim3((unsigned)blocks), dim3(threads), smem, (cudaStream_t)s, 1, P));
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

// ---------------------------------------------------------------- sub-pixel weight gradient -> 3x3 gradient
// dwp: gradient of the phase weights as HWIO-like [9][Cin][4*Cout] (tap (U,V), column (a*2+b)*Cout + co).
// Tap u of the upsampled kernel lands on low-res offset U(a,u): a=0: 0,1,1; a=1: 1,1,2 (same for columns), so
// dw[u][v][ci][co] += sum_{a,b} dwp[U(a,u)*3 + V(b,v)][ci][(a*2+b)*Cout + co].
__global__ void __launch_bounds__(256) wgrad_unphase_kernel(const float* __restrict__ dwp, float* __restrict__ dw,
                                                            int Cin, int Cout) {
  const long long total = 9ll * Cin * Cout;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int co = (int)(t % Cout);
  const long long q = t / Cout;
  const int ci = (int)(q % Cin);
  const int tap = (int)(q / Cin);
  const int u = tap / 3, v = tap % 3;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int U = a == 0 ? (u == 0 ? 0 : 1) : (u == 2 ? 2 : 1);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int V = b == 0 ? (v == 0 ? 0 : 1) : (v == 2 ? 2 : 1);
      acc += __ldg(dwp + ((long long)(U * 3 + V) * Cin + ci) * (4 * Cout) + (a * 2 + b) * Cout + co);
    }
  }
  dw[t] += acc;
}

extern "C" int tmx_conv_wgrad_unphase(tmx_handle_t h, const float* dwp, float* dw, int Cin, int Cout, tmx_stream_t s) {
  TMX_REQUIRE(h && dwp && dw && Cin > 0 && Cout > 0, TMX_ERR_ARG, "tmx_conv_wgrad_unphase: bad argument");
  const long long total = 9ll * Cin * Cout;
  wgrad_unphase_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(dwp, dw, Cin, Cout);
  TMX_LAUNCHED(h, "wgrad_unphase_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- RGB heads (1x1 convs with 3 image channels)
// ToRGB backward (networks.py:454-457 + tanh :483).  Forward: img[n][o][p] = f(ws * sum_c y[p][c] w[c][o] + b[o]).
// Given dimg (NCHW) and the forward image (for tanh' = 1 - img^2):  dpre = dimg * (1 - img^2);
//   dy[p][c] = ws * sum_o dpre[o] w[c][o];  dw[c][o] += ws * sum_p y[p][c] dpre[o];  db[o] += sum_p dpre[o].
// One thread per pixel; block partial sums -> atomics (3*(Cin+1) values per block).
template <int CIN>
__global__ void __launch_bounds__(256) torgb_bwd_kernel(const float* __restrict__ dimg, const float* __restrict__ img,
                                                        const float* __restrict__ y, const float* __restrict__ w,
                                                        float wscale, float* __restrict__ dy, float* __restrict__ dw,
                                                        float* __restrict__ db, long long npix, int HW, int Cimg,
                                                        int use_tanh) {
  __shared__ float red[4][CIN + 1][8];   // [o][c or bias][warp]
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dpre[4] = {0.f, 0.f, 0.f, 0.f};
  float yv[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) yv[c] = 0.f;
  if (p < npix) {
    const long long n = p / HW;
    const int hw = (int)(p % HW);
    for (int o = 0; o < Cimg; ++o) {
      const long long idx = (n * Cimg + o) * HW + hw;
      float g = __ldg(dimg + idx);
      if (use_tanh) {
        const float t = __ldg(img + idx);
        g *= (1.f - t * t);
      }
      dpre[o] = g;
    }
    const float4* yp = reinterpret_cast<const float4*>(y + p * CIN);
#pragma unroll
    for (int c4 = 0; c4 < CIN / 4; ++c4) {
      const float4 t = __ldg(yp + c4);
      yv[4 * c4] = t.x; yv[4 * c4 + 1] = t.y; yv[4 * c4 + 2] = t.z; yv[4 * c4 + 3] = t.w;
    }
    float4* dyp = reinterpret_cast<float4*>(dy + p * CIN);
#pragma unroll
    for (int c4 = 0; c4 < CIN / 4; ++c4) {
      float o4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float acc = 0.f;
        for (int o = 0; o < Cimg; ++o) acc = fmaf(dpre[o], __ldg(w + (4 * c4 + i) * Cimg + o), acc);
        o4[i] = acc * wscale;
      }
      dyp[c4] = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
  }
  // block reduction of y[c]*dpre[o] and dpre[o]
  for (int o = 0; o < Cimg; ++o) {
#pragma unroll
    for (int c = 0; c <= CIN; ++c) {
      float vsum = (c < CIN ? yv[c < CIN ? c : 0] : 1.f) * dpre[o];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, sft);
      if (lane == 0) red[o][c][warp] = vsum;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < Cimg * (CIN + 1); e += blockDim.x) {
    const int o = e / (CIN + 1), c = e % (CIN + 1);
    float sacc = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sacc += red[o][c][wv];
    if (c < CIN) atomicAdd(dw + c * Cimg + o, sacc * wscale);
    else if (db != nullptr) atomicAdd(db + o, sacc);
  }
}

extern "C" int tmx_torgb_bwd(tmx_handle_t h, const float* dimg, const float* img, const float* y, const float* w,
                             float wscale, float* dy, float* dw, float* db, int N, int H, int W, int Cin, int Cimg,
                             int use_tanh, tmx_stream_t s) {
  TMX_REQUIRE(h && dimg && y && w && dy && dw && (!use_tanh || img), TMX_ERR_ARG, "tmx_torgb_bwd: NULL argument");
  TMX_REQUIRE(N > 0 && H > 0 && W > 0 && Cimg >= 1 && Cimg <= 4 && (Cin == 16 || Cin == 32 || Cin == 64), TMX_ERR_SHAPE,
              "tmx_torgb_bwd: bad shape (Cin in {16,32,64}, Cimg <= 4), got Cin=%d Cimg=%d", Cin, Cimg);
  const long long npix = (long long)N * H * W;
  dim3 grid(tmx_ceil_div(npix, 256));
  cudaStream_t st = (cudaStream_t)s;
  if (Cin == 16) torgb_bwd_kernel<16><<<grid, 256, 0, st>>>(dimg, img, y, w, wscale, dy, dw, db, npix, H * W, Cimg, use_tanh);
  else if (Cin == 32) torgb_bwd_kernel<32><<<grid, 256, 0, st>>>(dimg, img, y, w, wscale, dy, dw, db, npix, H * W, Cimg, use_tanh);
  else torgb_bwd_kernel<64><<<grid, 256, 0, st>>>(dimg, img, y, w, wscale, dy, dw, db, npix, H * W, Cimg, use_tanh);
  TMX_LAUNCHED(h, "torgb_bwd_kernel");
  return TMX_OK;
}

// FromRGB backward (networks.py:226-228): forward y[p][o] = lrelu(ws * sum_c img[c][p] w[c][o] + b[o]).
// dz (already masked, NHWC fp32 [npix][Cout]) ->  dw[c][o] += ws * sum_p img[c][p] dz[p][o];  optionally
// dimg[n][c][p] = ws * sum_o dz[p][o] w[c][o] (needed when the image itself is a function of trained weights).
template <int COUT>
__global__ void __launch_bounds__(256) fromrgb_bwd_kernel(const float* __restrict__ img, const float* __restrict__ dz,
                                                          const float* __restrict__ w, float wscale,
                                                          float* __restrict__ dw, float* __restrict__ dimg,
                                                          long long npix, int HW, int Cimg, int Ctot, int c0) {
  // COUT channels [c0, c0 + COUT) of a head with Ctot output channels (Ctot > 64: one launch per 64-channel chunk,
  // the chunks after the first ADD their share to dimg - same stream, so the launches are ordered)
  __shared__ float red[4][COUT][8];
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float iv[4] = {0.f, 0.f, 0.f, 0.f};
  float zv[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) zv[o] = 0.f;
  if (p < npix) {
    const long long n = p / HW;
    const int hw = (int)(p % HW);
    for (int c = 0; c < Cimg; ++c) iv[c] = __ldg(img + (n * Cimg + c) * HW + hw);
    const float4* zp = reinterpret_cast<const float4*>(dz + p * Ctot + c0);
#pragma unroll
    for (int o4 = 0; o4 < COUT / 4; ++o4) {
      const float4 t = __ldg(zp + o4);
      zv[4 * o4] = t.x; zv[4 * o4 + 1] = t.y; zv[4 * o4 + 2] = t.z; zv[4 * o4 + 3] = t.w;
    }
    if (dimg != nullptr) {
      for (int c = 0; c < Cimg; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc = fmaf(zv[o], __ldg(w + c * Ctot + c0 + o), acc);
        float* dst = dimg + (n * Cimg + c) * HW + hw;
        *dst = c0 == 0 ? acc * wscale : fmaf(acc, wscale, *dst);
      }
    }
  }
  for (int c = 0; c < Cimg; ++c) {
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      float vsum = iv[c] * zv[o];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, sft);
      if (lane == 0) red[c][o][warp] = vsum;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < Cimg * COUT; e += blockDim.x) {
    const int c = e / COUT, o = e % COUT;
    float sacc = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) sacc += red[c][o][wv];
    atomicAdd(dw + c * Ctot + c0 + o, sacc * wscale);
  }
}

extern "C" int tmx_fromrgb_bwd(tmx_handle_t h, const float* img, const float* dz, const float* w, float wscale,
                               float* dw, float* dimg, int N, int Cimg, int H, int W, int Cout, tmx_stream_t s) {
  TMX_REQUIRE(h && img && dz && w && dw, TMX_ERR_ARG, "tmx_fromrgb_bwd: NULL argument");
  TMX_REQUIRE(N > 0 && H > 0 && W > 0 && Cimg >= 1 && Cimg <= 4 && (Cout == 16 || Cout == 32 || Cout % 64 == 0),
              TMX_ERR_SHAPE, "tmx_fromrgb_bwd: bad shape (Cout 16, 32 or a multiple of 64, Cimg <= 4), got Cout=%d Cimg=%d",
              Cout, Cimg);
  const long long npix = (long long)N * H * W;
  dim3 grid(tmx_ceil_div(npix, 256));
  cudaStream_t st = (cudaStream_t)s;
  if (Cout == 16) {
    fromrgb_bwd_kernel<16><<<grid, 256, 0, st>>>(img, dz, w, wscale, dw, dimg, npix, H * W, Cimg, 16, 0);
    TMX_LAUNCHED(h, "fromrgb_bwd_kernel");
  } else if (Cout == 32) {
    fromrgb_bwd_kernel<32><<<grid, 256, 0, st>>>(img, dz, w, wscale, dw, dimg, npix, H * W, Cimg, 32, 0);
    TMX_LAUNCHED(h, "fromrgb_bwd_kernel");
  } else {
    for (int c0 = 0; c0 < Cout; c0 += 64) {   // the lower-resolution heads of progressive growing: 3 -> 128..512
      fromrgb_bwd_kernel<64><<<grid, 256, 0, st>>>(img, dz, w, wscale, dw, dimg, npix, H * W, Cimg, Cout, c0);
      TMX_LAUNCHED(h, "fromrgb_bwd_kernel");
    }
  }
  return TMX_OK;
}

// ---------------------------------------------------------------- dense / minibatch-stddev input gradients (D head)
// dx[n][k] = wscale * sum_o dz[n][o] * w[k][o],  dz = dy * lrelu'(y) (networks.py:38-43, 72-75).
// Block = 64 input features x 32 samples; the contraction over o runs in chunks of 32 staged in shared memory
// (w rows read coalesced along o).  Thread (kk, ng) owns feature kk and samples ng*8 .. ng*8+7.
__global__ void __launch_bounds__(256) dense_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                              const float* __restrict__ w, float wscale,
                                                              float* __restrict__ dx, int N, int K, int Cout, int lrelu,
                                                              float alpha) {
  __shared__ float ws[64][33];
  __shared__ float zs[32][33];
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 32;
  const int kk = threadIdx.x & 63, ng = threadIdx.x >> 6;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int o0 = 0; o0 < Cout; o0 += 32) {
    for (int e = threadIdx.x; e < 64 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      ws[r][c] = (k0 + r < K && o0 + c < Cout) ? __ldg(w + (long long)(k0 + r) * Cout + o0 + c) : 0.f;
    }
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      float g = 0.f;
      if (n0 + r < N && o0 + c < Cout) {
        g = __ldg(dy + (long long)(n0 + r) * Cout + o0 + c);
        if (lrelu) g *= (__ldg(y + (long long)(n0 + r) * Cout + o0 + c) > 0.f) ? 1.f : alpha;
      }
      zs[r][c] = g;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const float wv = ws[kk][c];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(zs[ng * 8 + i][c], wv, acc[i]);
    }
    __syncthreads();
  }
  if (k0 + kk < K) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + ng * 8 + i;
      if (n < N) dx[(long long)n * K + k0 + kk] = acc[i] * wscale;
    }
  }
}

// Fast path (Cout % 64 == 0, K % 32 == 0): 32 features per block (twice the blocks of the kernel above: the 8192-feature
// head fills the SMs), 64-output chunks, 128-bit loads of the weight rows; thread (kk, ng) owns feature kk and samples
// ng*4 .. ng*4+3.
__global__ void __launch_bounds__(256) dense_bwd_input4_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                               const float* __restrict__ w, float wscale,
                                                               float* __restrict__ dx, int N, int K, int Cout,
                                                               int lrelu, float alpha) {
  __shared__ float ws[32][65];
  __shared__ float zs[32][65];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int kk = threadIdx.x & 31, ng = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int o0 = 0; o0 < Cout; o0 += 64) {
    for (int e = threadIdx.x; e < 32 * 16; e += 256) {
      const int r = e >> 4, c4 = (e & 15) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(w + (long long)(k0 + r) * Cout + o0 + c4));
      ws[r][c4] = v.x; ws[r][c4 + 1] = v.y; ws[r][c4 + 2] = v.z; ws[r][c4 + 3] = v.w;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < N) {
        g = __ldg(reinterpret_cast<const float4*>(dy + (long long)(n0 + r) * Cout + o0 + c4));
        if (lrelu) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(y + (long long)(n0 + r) * Cout + o0 + c4));
          g.x *= yv.x > 0.f ? 1.f : alpha;
          g.y *= yv.y > 0.f ? 1.f : alpha;
          g.z *= yv.z > 0.f ? 1.f : alpha;
          g.w *= yv.w > 0.f ? 1.f : alpha;
        }
      }
      zs[r][c4] = g.x; zs[r][c4 + 1] = g.y; zs[r][c4 + 2] = g.z; zs[r][c4 + 3] = g.w;
    }
    __syncthreads();
#pragma unroll 16
    for (int c = 0; c < 64; ++c) {
      const float wv = ws[kk][c];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(zs[ng * 4 + i][c], wv, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ng * 4 + i;
    if (n < N) dx[(long long)n * K + k0 + kk] = acc[i] * wscale;
  }
}

// Streaming path of the 8192 <- 512 head (Cout == 512, K % 64 == 0, N <= 32 per block): dx[n][k] = wscale *
// sum_o g[n][o] * w[k][o] with g = dy * act'(y).  The kernel above walks the 512 outputs in eight load / barrier /
// compute rounds with 8 KB of weights in flight per block (35 us for 16.8 MB).  Here g of all 32 samples sits in
// shared memory (64 KB), a block streams 64 weight rows in two passes of 32 - sixteen threads per row, each holding 8
// interleaved float4 of it in registers, the second pass already in flight while the first is multiplied - and a
// row's 32 dot products are finished by four xor-shuffles.
constexpr int kDbiRows = 64;
__global__ void __launch_bounds__(512, 1) dense_bwd_input_stream_kernel(const float* __restrict__ dy,
                                                                       const float* __restrict__ y,
                                                                       const float* __restrict__ w, float wscale,
                                                                       float* __restrict__ dx, int N, int K, int lrelu,
                                                                       float alpha) {
  // 512 threads = 32 rows x 16 threads per row (the 256-thread version issued 1.0 instructions per cycle: FFMA-issue
  // bound with eight warps per SM, ncu)
  extern __shared__ __align__(16) float g_s[];     // [32][512]
  constexpr int Cout = 512;
  const int t = threadIdx.x, r = t >> 4, sg = t & 15;
  const int k0 = blockIdx.x * kDbiRows, n0 = blockIdx.y * 32;
  const float4* w0 = reinterpret_cast<const float4*>(w + (long long)(k0 + r) * Cout) + sg;
  const float4* w1 = reinterpret_cast<const float4*>(w + (long long)(k0 + 32 + r) * Cout) + sg;
  float4 wa[8], wb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) wa[j] = __ldg(w0 + j * 16);           // float4 column j*16 + sg of the row
  // g = dy * act'(y) of the 32 samples: 8 float4 per thread, all loads in flight together (not a chain of L2 round trips)
  {
    const int nvalid4 = min(32, N - n0) * (Cout / 4);             // rows of dy / y are contiguous: [n0 .. n0+32) x Cout
    const float4* dyp = reinterpret_cast<const float4*>(dy + (long long)n0 * Cout);
    const float4* yp = reinterpret_cast<const float4*>(y + (long long)n0 * Cout);
    float4 gv[8], yv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = t + 512 * i;
      gv[i] = e < nvalid4 ? __ldg(dyp + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      yv[i] = (lrelu && e < nvalid4) ? __ldg(yp + e) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      gv[i].x *= yv[i].x > 0.f ? 1.f : alpha;
      gv[i].y *= yv[i].y > 0.f ? 1.f : alpha;
      gv[i].z *= yv[i].z > 0.f ? 1.f : alpha;
      gv[i].w *= yv[i].w > 0.f ? 1.f : alpha;
      reinterpret_cast<float4*>(g_s)[t + 512 * i] = gv[i];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) wb[j] = __ldg(w1 + j * 16);
  __syncthreads();
  auto pass = [&](const float4 (&wv)[8], int row) {
#pragma unroll 4
    for (int n = 0; n < 32; ++n) {
      const float4* gp = reinterpret_cast<const float4*>(g_s + n * Cout) + sg;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 gv = gp[j * 16];
        a0 = fmaf(gv.x, wv[j].x, a0);
        a1 = fmaf(gv.y, wv[j].y, a1);
        a2 = fmaf(gv.z, wv[j].z, a2);
        a3 = fmaf(gv.w, wv[j].w, a3);
      }
      float a = (a0 + a1) + (a2 + a3);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      if ((n & 15) == sg && n0 + n < N) dx[(long long)(n0 + n) * K + row] = a * wscale;
    }
  };
  pass(wa, k0 + r);
  pass(wb, k0 + 32 + r);
}

extern "C" int tmx_dense_bwd_input(tmx_handle_t h, const float* dy, const float* y, const float* w, float wscale,
                                   float* dx, int N, int K, int Cout, int lrelu, float alpha, tmx_stream_t s) {
  TMX_REQUIRE(h && dy && w && dx && (!lrelu || y), TMX_ERR_ARG, "tmx_dense_bwd_input: NULL argument");
  TMX_REQUIRE(N > 0 && K > 0 && Cout > 0, TMX_ERR_SHAPE, "tmx_dense_bwd_input: bad shape");
  if (Cout == 512 && K % kDbiRows == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)w) & 15) == 0 &&
      !tmx_env_flag("TMX_DENSE_NO_STREAM")) {
    auto kern = dense_bwd_input_stream_kernel;
    const int smem = 32 * 512 * (int)sizeof(float);
    static thread_local int configured_device = -1;
    if (configured_device != h->device) {
      TMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured_device = h->device;
    }
    kern<<<dim3(K / kDbiRows, tmx_ceil_div(N, 32)), 512, smem, (cudaStream_t)s>>>(dy, y, w, wscale, dx, N, K, lrelu, alpha);
    TMX_LAUNCHED(h, "dense_bwd_input_stream_kernel");
    return TMX_OK;
  }
  if (Cout % 64 == 0 && K % 32 == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)w) & 15) == 0) {
    dim3 grid4(K / 32, tmx_ceil_div(N, 32));
    dense_bwd_input4_kernel<<<grid4, 256, 0, (cudaStream_t)s>>>(dy, y, w, wscale, dx, N, K, Cout, lrelu, alpha);
    TMX_LAUNCHED(h, "dense_bwd_input_kernel");
    return TMX_OK;
  }
  dim3 grid(tmx_ceil_div(K, 64), tmx_ceil_div(N, 32));
  dense_bwd_input_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(dy, y, w, wscale, dx, N, K, Cout, lrelu, alpha);
  TMX_LAUNCHED(h, "dense_bwd_input_kernel");
  return TMX_OK;
}

// minibatch_stddev_layer backward (networks.py:177-189).  y = [x, s broadcast, 0...]:
//   dx = dy[..., :C] + ds[m] * d s[m] / dx,   ds[m] = sum over the group's samples and pixels of dy[..., C],
//   d s[m] / dx[g,m,e] = (x[g,m,e] - mean_g) / (G * sigma[m,e] * E),  sigma = sqrt(var_g + 1e-8), E = H*W*C.
__global__ void __launch_bounds__(256) mbstd_ds_kernel(const float* __restrict__ dy, float* __restrict__ ds, int G,
                                                       int M, int HW, int C, int C_total) {
  const int m = blockIdx.x;
  float acc = 0.f;
  for (int e = threadIdx.x; e < G * HW; e += blockDim.x) {
    const int g = e / HW, p = e % HW;
    acc += __ldg(dy + ((long long)(g * M + m) * HW + p) * C_total + C);
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) ds[m] = red[0];
}

__global__ void __launch_bounds__(256) mbstd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ ds, float* __restrict__ dx, int G,
                                                        int M, long long per_sample, int C, int C_total) {
  // one thread per (m, e): loops over the G samples of the group
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)M * per_sample) return;
  const int m = (int)(t / per_sample);
  const long long e = t % per_sample;
  const long long pix = e / C;
  const int c = (int)(e % C);
  float mean = 0.f;
  for (int g = 0; g < G; ++g) mean += __ldg(x + (long long)(g * M + m) * per_sample + e);
  mean /= (float)G;
  float var = 0.f;
  for (int g = 0; g < G; ++g) {
    const float d = __ldg(x + (long long)(g * M + m) * per_sample + e) - mean;
    var += d * d;
  }
  const float sigma = sqrtf(var / (float)G + 1e-8f);
  const float coef = __ldg(ds + m) / ((float)G * sigma * (float)per_sample);
  const long long HW = per_sample / C;
  for (int g = 0; g < G; ++g) {
    const long long n = g * M + m;
    const float xv = __ldg(x + n * per_sample + e);
    dx[n * per_sample + e] = __ldg(dy + (n * HW + pix) * C_total + c) + coef * (xv - mean);
  }
}

extern "C" int tmx_mbstd_bwd(tmx_handle_t h, const float* x, const float* dy, float* dx, float* ds, int N, int H, int W,
                             int C, int C_total, int group_size, tmx_stream_t s) {
  TMX_REQUIRE(h && x && dy && dx && ds, TMX_ERR_ARG, "tmx_mbstd_bwd: NULL argument");
  TMX_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C_total > C, TMX_ERR_SHAPE, "tmx_mbstd_bwd: bad shape");
  const int G = group_size < N ? group_size : N;
  TMX_REQUIRE(N % G == 0, TMX_ERR_SHAPE, "tmx_mbstd_bwd: batch %d is not a multiple of the group size %d", N, G);
  const int M = N / G;
  const long long per_sample = (long long)H * W * C;
  mbstd_ds_kernel<<<M, 256, 0, (cudaStream_t)s>>>(dy, ds, G, M, H * W, C, C_total);
  TMX_LAUNCHED(h, "mbstd_ds_kernel");
  mbstd_bwd_kernel<<<tmx_ceil_div((long long)M * per_sample, 256), 256, 0, (cudaStream_t)s>>>(x, dy, ds, dx, G, M,
                                                                                             per_sample, C, C_total);
  TMX_LAUNCHED(h, "mbstd_bwd_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- loss gradients (loss.py:105-259)
// L1 image term: loss_n = weight * mean_{c,h,w} |a - b| ; the optimizer takes the batch mean (run.py:321), so
// d/da = scale * sign(a - b) with scale = weight / (C*H*W*N); *loss_sum accumulates sum |a - b| for reporting.
__global__ void __launch_bounds__(256) l1_grad_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      float* __restrict__ grad, float* __restrict__ loss_sum,
                                                      long long n, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = __ldg(a + i) - __ldg(b + i);
    acc += fabsf(d);
    grad[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
  }
  if (loss_sum != nullptr) {
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss_sum, acc);
  }
}

extern "C" int tmx_loss_l1_grad(tmx_handle_t h, const float* a, const float* b, float* grad, float* loss_sum, int64_t n,
                                float scale, tmx_stream_t s) {
  TMX_REQUIRE(h && a && b && grad && n > 0, TMX_ERR_ARG, "tmx_loss_l1_grad: bad argument");
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)h->sm_count * 8) blocks = (long long)h->sm_count * 8;
  l1_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(a, b, grad, loss_sum, n, scale);
  TMX_LAUNCHED(h, "l1_grad_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- adjoints of the latent canvas ops
// tiling_permutation (+ corner re-pin) is a gather; its adjoint scatters: d_src[n][c][sy][sx] += d_canvas[n][c][i][j].
// d_src must be zero-initialised (or hold an earlier contribution).  Same index conventions as tmx_latent_blend COPY.
__global__ void __launch_bounds__(256) blend_copy_bwd_kernel(const float* __restrict__ dcanvas, float* __restrict__ dsrc,
                                                             const int32_t* __restrict__ idx_h,
                                                             const int32_t* __restrict__ idx_w, int N, int C, int h,
                                                             int w, int H, int W, unsigned long long pin_rows,
                                                             unsigned long long pin_cols, int reverse) {
  const long long total = (long long)N * C * H * W;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int j = (int)(t % W);
  long long q = t / W;
  const int i = (int)(q % H);
  q /= H;
  const int c = (int)(q % C);
  const int n = (int)(q / C);
  const bool pinned = pin_rows != 0 && ((pin_rows >> (i / h)) & 1ull) && ((pin_cols >> (j / w)) & 1ull);
  int yy = i, xx = j;
  if (!pinned) {
    if (idx_h) yy = __ldg(idx_h + (long long)n * H + i);
    if (idx_w) xx = __ldg(idx_w + (long long)n * W + j);
  }
  const int nn = reverse ? N - 1 - n : n;
  atomicAdd(dsrc + (((long long)nn * C + c) * h + yy % h) * w + xx % w, __ldg(dcanvas + t));
}

extern "C" int tmx_latent_gather_bwd(tmx_handle_t h, const float* dcanvas, float* dsrc, const int32_t* idx_h,
                                     const int32_t* idx_w, int N, int C, int sh, int sw, int H, int W, uint64_t pin_rows,
                                     uint64_t pin_cols, int reverse, tmx_stream_t s) {
  TMX_REQUIRE(h && dcanvas && dsrc, TMX_ERR_ARG, "tmx_latent_gather_bwd: NULL argument");
  TMX_REQUIRE(N > 0 && C > 0 && sh > 0 && sw > 0 && H > 0 && W > 0, TMX_ERR_SHAPE, "tmx_latent_gather_bwd: bad shape");
  const long long total = (long long)N * C * H * W;
  blend_copy_bwd_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(dcanvas, dsrc, idx_h, idx_w, N, C, sh, sw,
                                                                                H, W, pin_rows, pin_cols, reverse);
  TMX_LAUNCHED(h, "blend_copy_bwd_kernel");
  return TMX_OK;
}

// The same scatter restricted to a [wh x ww] window of the canvas at (oy, ox) - `dwin` is the gradient w.r.t. that
// window only (crop-aware G_fcn: everything outside it is zero), the offset comes from `off_dev` = {oy, ox} on the
// device when given (CUDA-graph replays): no zero canvas is built and only wh*ww of the H*W positions issue an atomic.
__global__ void __launch_bounds__(256) blend_copy_bwd_window_kernel(const float* __restrict__ dwin, float* __restrict__ dsrc,
                                                                    const int32_t* __restrict__ idx_h,
                                                                    const int32_t* __restrict__ idx_w, int N, int C,
                                                                    int h, int w, int H, int W, int wh, int ww, int oy,
                                                                    int ox, const int32_t* __restrict__ off_dev,
                                                                    unsigned long long pin_rows,
                                                                    unsigned long long pin_cols, int reverse) {
  const long long total = (long long)N * C * wh * ww;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  if (off_dev != nullptr) {
    oy = __ldg(off_dev);
    ox = __ldg(off_dev + 1);
  }
  const int j = ox + (int)(t % ww);
  long long q = t / ww;
  const int i = oy + (int)(q % wh);
  q /= wh;
  const int c = (int)(q % C);
  const int n = (int)(q / C);
  if (i < 0 || i >= H || j < 0 || j >= W) return;
  const bool pinned = pin_rows != 0 && ((pin_rows >> (i / h)) & 1ull) && ((pin_cols >> (j / w)) & 1ull);
  int yy = i, xx = j;
  if (!pinned) {
    if (idx_h) yy = __ldg(idx_h + (long long)n * H + i);
    if (idx_w) xx = __ldg(idx_w + (long long)n * W + j);
  }
  const int nn = reverse ? N - 1 - n : n;
  atomicAdd(dsrc + (((long long)nn * C + c) * h + yy % h) * w + xx % w, __ldg(dwin + t));
}

extern "C" int tmx_latent_gather_bwd_window(tmx_handle_t h, const float* dwin, float* dsrc, const int32_t* idx_h,
                                            const int32_t* idx_w, int N, int C, int sh, int sw, int H, int W, int wh,
                                            int ww, int oy, int ox, const int32_t* off_dev, uint64_t pin_rows,
                                            uint64_t pin_cols, int reverse, tmx_stream_t s) {
  TMX_REQUIRE(h && dwin && dsrc, TMX_ERR_ARG, "tmx_latent_gather_bwd_window: NULL argument");
  TMX_REQUIRE(N > 0 && C > 0 && sh > 0 && sw > 0 && H > 0 && W > 0 && wh > 0 && ww > 0 && wh <= H && ww <= W,
              TMX_ERR_SHAPE, "tmx_latent_gather_bwd_window: bad shape");
  const long long total = (long long)N * C * wh * ww;
  blend_copy_bwd_window_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(
      dwin, dsrc, idx_h, idx_w, N, C, sh, sw, H, W, wh, ww, oy, ox, off_dev, pin_rows, pin_cols, reverse);
  TMX_LAUNCHED(h, "blend_copy_bwd_window_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- sampled latent canvases (config-off interpolation modes)
// loss.py:176-193, 218-235 with zg_/zl_interp_variational = 'variational' / 'random' (the reference config uses 'hard' /
// 'permutational').  Sources mu, ls: [N][C][sh][sw] (the encoder's two outputs); eps: [N][C][eh][ew] standard-normal
// draws (eh x ew = 1 x 1 for the global code, H x W for the local one); canvas [N][C][H][W]; reverse: the sources are
// read batch-reversed (tf.reverse(axis=[0]) of the blend branch).
//   mode 1 'variational': out = eps[i % eh][j % ew] * exp(ls[i % sh][j % sw]) + mu[i % sh][j % sw]
//   mode 2 'random'     : out = mu[i % sh][j % sw] on the four corner tiles, eps[i][j] elsewhere
__global__ void __launch_bounds__(256) latent_noise_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls,
                                                               const float* __restrict__ eps, float* __restrict__ out,
                                                               long long total, int N, int C, int sh, int sw, int eh,
                                                               int ew, int H, int W, int mode, int reverse) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int j = (int)(t % W);
  long long q = t / W;
  const int i = (int)(q % H);
  q /= H;
  const int c = (int)(q % C);
  const int n = (int)(q / C);
  const int ns = reverse ? N - 1 - n : n;
  const long long so = (((long long)ns * C + c) * sh + i % sh) * sw + j % sw;
  const float e = __ldg(eps + (((long long)n * C + c) * eh + i % eh) * ew + j % ew);
  if (mode == 1) {
    out[t] = e * expf(__ldg(ls + so)) + __ldg(mu + so);
  } else {
    const bool corner = (i < sh || i >= H - sh) && (j < sw || j >= W - sw);
    out[t] = corner ? __ldg(mu + so) : e;
  }
}

// adjoint: one thread per source element (ns, c, y, x); dmu / dls are ACCUMULATED (+=), dls may be NULL in mode 2
__global__ void __launch_bounds__(256) latent_noise_bwd_kernel(const float* __restrict__ g, const float* __restrict__ ls,
                                                               const float* __restrict__ eps, float* __restrict__ dmu,
                                                               float* __restrict__ dls, long long total, int N, int C,
                                                               int sh, int sw, int eh, int ew, int H, int W, int mode,
                                                               int reverse) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int x = (int)(t % sw);
  long long q = t / sw;
  const int y = (int)(q % sh);
  q /= sh;
  const int c = (int)(q % C);
  const int ns = (int)(q / C);
  const int n = reverse ? N - 1 - ns : ns;
  const float* gn = g + ((long long)n * C + c) * H * W;
  const float* en = eps + ((long long)n * C + c) * eh * ew;
  float am = 0.f, al = 0.f;
  const float sig = mode == 1 ? expf(__ldg(ls + t)) : 0.f;
  for (int i = y; i < H; i += sh)
    for (int j = x; j < W; j += sw) {
      const float gv = __ldg(gn + (long long)i * W + j);
      if (mode == 1) {
        am += gv;
        al += gv * __ldg(en + (long long)(i % eh) * ew + j % ew) * sig;
      } else if ((i < sh || i >= H - sh) && (j < sw || j >= W - sw)) {
        am += gv;
      }
    }
  dmu[t] += am;
  if (mode == 1 && dls != nullptr) dls[t] += al;
}

extern "C" int tmx_latent_noise_fwd(tmx_handle_t h, int mode, const float* mu, const float* ls, const float* eps,
                                    float* out, int N, int C, int sh, int sw, int eh, int ew, int H, int W, int reverse,
                                    tmx_stream_t s) {
  TMX_REQUIRE(h && mu && eps && out && (mode == 2 || ls), TMX_ERR_ARG, "tmx_latent_noise_fwd: NULL argument");
  TMX_REQUIRE((mode == 1 || mode == 2) && N > 0 && C > 0 && sh > 0 && sw > 0 && eh > 0 && ew > 0 && H % sh == 0 &&
                  W % sw == 0 && H % eh == 0 && W % ew == 0 && (mode == 1 || (eh == H && ew == W)),
              TMX_ERR_SHAPE, "tmx_latent_noise_fwd: bad shape / mode");
  const long long total = (long long)N * C * H * W;
  latent_noise_fwd_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(mu, ls, eps, out, total, N, C, sh, sw, eh,
                                                                                 ew, H, W, mode, reverse);
  TMX_LAUNCHED(h, "latent_noise_fwd_kernel");
  return TMX_OK;
}

extern "C" int tmx_latent_noise_bwd(tmx_handle_t h, int mode, const float* g, const float* ls, const float* eps,
                                    float* dmu, float* dls, int N, int C, int sh, int sw, int eh, int ew, int H, int W,
                                    int reverse, tmx_stream_t s) {
  TMX_REQUIRE(h && g && eps && dmu && (mode == 2 || (ls && dls)), TMX_ERR_ARG, "tmx_latent_noise_bwd: NULL argument");
  TMX_REQUIRE((mode == 1 || mode == 2) && N > 0 && C > 0 && sh > 0 && sw > 0 && eh > 0 && ew > 0 && H % sh == 0 &&
                  W % sw == 0 && H % eh == 0 && W % ew == 0 && (mode == 1 || (eh == H && ew == W)),
              TMX_ERR_SHAPE, "tmx_latent_noise_bwd: bad shape / mode");
  const long long total = (long long)N * C * sh * sw;
  latent_noise_bwd_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(g, ls, eps, dmu, dls, total, N, C, sh, sw,
                                                                                 eh, ew, H, W, mode, reverse);
  TMX_LAUNCHED(h, "latent_noise_bwd_kernel");
  return TMX_OK;
}

// tf.tile of a [N][C][1][1] code over the canvas (loss.py:176): adjoint = sum over the canvas, one block per (n, c).
// out[row] (+)= scale * sum_i in[row][i]
__global__ void __launch_bounds__(256) row_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int len,
                                                      float scale, int accumulate, int square) {
  const long long row = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const float v = __ldg(in + row * len + i);
    acc += square ? v * v : v;
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[row] = (accumulate ? out[row] : 0.f) + red[0] * scale;
}

extern "C" int tmx_row_sum(tmx_handle_t h, const float* in, float* out, int rows, int len, float scale, int accumulate,
                           int square, tmx_stream_t s) {
  TMX_REQUIRE(h && in && out && rows > 0 && len > 0, TMX_ERR_ARG, "tmx_row_sum: bad argument");
  row_sum_kernel<<<rows, 256, 0, (cudaStream_t)s>>>(in, out, len, scale, accumulate, square);
  TMX_LAUNCHED(h, "row_sum_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- dense weight gradient, mbstd tangent / curvature
// dw[k][o] += wscale * sum_n x[n][k] * dz[n][o],  db[o] += sum_n dz[n][o],  dz = dy * lrelu'(y)
__global__ void __launch_bounds__(256) dense_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          const float* __restrict__ y, float* __restrict__ dw,
                                                          float* __restrict__ db, int N, int K, int Cout, float wscale,
                                                          int lrelu, float alpha) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)K * Cout) return;
  const int o = (int)(t % Cout);
  const int k = (int)(t / Cout);
  float acc = 0.f, bacc = 0.f;
  for (int n = 0; n < N; ++n) {
    float g = __ldg(dy + (long long)n * Cout + o);
    if (lrelu) g *= (__ldg(y + (long long)n * Cout + o) > 0.f) ? 1.f : alpha;
    acc = fmaf(__ldg(x + (long long)n * K + k), g, acc);
    bacc += g;
  }
  dw[t] += acc * wscale;
  if (db != nullptr && k == 0) db[o] += bacc;
}

// Fast path (Cout % 128 == 0, K % 64 == 0): block = 128 outputs x 64 features, thread = 4 outputs x 8 features
// (32 accumulators); the masked dz tile and the x tile of 32 samples at a time sit in shared memory, dw is updated
// with 128-bit read-modify-writes: the 16.8 MB gradient of the 8192 -> 512 head moves once.
__global__ void __launch_bounds__(256) dense_wgrad4_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ y, float* __restrict__ dw,
                                                           float* __restrict__ db, int N, int K, int Cout, float wscale,
                                                           int lrelu, float alpha) {
  __shared__ float zs[32][128];      // masked dz: [sample][output]
  __shared__ float xs[32][64 + 4];   // x: [sample][feature]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o0 = blockIdx.x * 128, k0 = blockIdx.y * 64;
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
  // the gradient rows this thread accumulates into: fetched now, so that the DRAM round trip of the read-modify-write
  // overlaps the staging and the products instead of following them
  float4 cur[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    cur[i] = *reinterpret_cast<const float4*>(dw + (long long)(k0 + ty * 8 + i) * Cout + o0 + tx * 4);
  for (int n0 = 0; n0 < N; n0 += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int n = e >> 5, c4 = (e & 31) * 4;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + n < N) {
        g = __ldg(reinterpret_cast<const float4*>(dy + (long long)(n0 + n) * Cout + o0 + c4));
        if (lrelu) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(y + (long long)(n0 + n) * Cout + o0 + c4));
          g.x *= yv.x > 0.f ? 1.f : alpha;
          g.y *= yv.y > 0.f ? 1.f : alpha;
          g.z *= yv.z > 0.f ? 1.f : alpha;
          g.w *= yv.w > 0.f ? 1.f : alpha;
        }
      }
      *reinterpret_cast<float4*>(&zs[n][c4]) = g;
    }
    for (int e = threadIdx.x; e < 32 * 16; e += 256) {
      const int n = e >> 4, k4 = (e & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + n < N) v = __ldg(reinterpret_cast<const float4*>(x + (long long)(n0 + n) * K + k0 + k4));
      *reinterpret_cast<float4*>(&xs[n][k4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int n = 0; n < 32; ++n) {
      const float4 g = *reinterpret_cast<const float4*>(&zs[n][tx * 4]);
      const float4 xa = *reinterpret_cast<const float4*>(&xs[n][ty * 8]);
      const float4 xb = *reinterpret_cast<const float4*>(&xs[n][ty * 8 + 4]);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i].x = fmaf(xv[i], g.x, acc[i].x);
        acc[i].y = fmaf(xv[i], g.y, acc[i].y);
        acc[i].z = fmaf(xv[i], g.z, acc[i].z);
        acc[i].w = fmaf(xv[i], g.w, acc[i].w);
      }
      bacc.x += g.x;
      bacc.y += g.y;
      bacc.z += g.z;
      bacc.w += g.w;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4* o = reinterpret_cast<float4*>(dw + (long long)(k0 + ty * 8 + i) * Cout + o0 + tx * 4);
    float4 c = cur[i];
    c.x = fmaf(acc[i].x, wscale, c.x);
    c.y = fmaf(acc[i].y, wscale, c.y);
    c.z = fmaf(acc[i].z, wscale, c.z);
    c.w = fmaf(acc[i].w, wscale, c.w);
    *o = c;
  }
  if (db != nullptr && blockIdx.y == 0 && ty == 0) {
    float4* o = reinterpret_cast<float4*>(db + o0 + tx * 4);
    float4 cur = *o;
    cur.x += bacc.x;
    cur.y += bacc.y;
    cur.z += bacc.z;
    cur.w += bacc.w;
    *o = cur;
  }
}

extern "C" int tmx_dense_wgrad(tmx_handle_t h, const float* x, const float* dy, const float* y, float* dw, float* db, int N,
                               int K, int Cout, float wscale, int lrelu, float alpha, tmx_stream_t s) {
  TMX_REQUIRE(h && x && dy && dw && (!lrelu || y), TMX_ERR_ARG, "tmx_dense_wgrad: NULL argument");
  TMX_REQUIRE(N > 0 && K > 0 && Cout > 0, TMX_ERR_SHAPE, "tmx_dense_wgrad: bad shape");
  if (Cout % 128 == 0 && K % 64 == 0 &&
      (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)y | (uintptr_t)dw | (uintptr_t)db) & 15) == 0) {
    dense_wgrad4_kernel<<<dim3(Cout / 128, K / 64), 256, 0, (cudaStream_t)s>>>(x, dy, y, dw, db, N, K, Cout, wscale, lrelu,
                                                                               alpha);
    TMX_LAUNCHED(h, "dense_wgrad_kernel");
    return TMX_OK;
  }
  dense_wgrad_kernel<<<tmx_ceil_div((long long)K * Cout, 256), 256, 0, (cudaStream_t)s>>>(x, dy, y, dw, db, N, K, Cout,
                                                                                         wscale, lrelu, alpha);
  TMX_LAUNCHED(h, "dense_wgrad_kernel");
  return TMX_OK;
}

// Forward-mode tangent and curvature of minibatch_stddev_layer (networks.py:177-189), for the WGAN-GP double
// backward (loss.py:332-336).  Per group m and element e over the G samples of the group:
//   sigma = sqrt(var_g(x) + 1e-8),  s[m] = mean_e sigma
//   d sigma / dx_g = (x_g - mu) / (G sigma)
//   tangent:   sdot[m] = (1/E) sum_e sum_g (x_g - mu) xdot_g / (G sigma)
//   curvature: (H xdot)_g = (xdot_g - mean_g xdot) / (G sigma) - (x_g - mu) * sum_h (x_h - mu) xdot_h / (G^2 sigma^3)
//   q[g][m][e] = lam[m] / E * (H xdot)_g      (gradient w.r.t. the PRIMAL x of  lam[m] * sdot[m])
// mode 0: ydot = [xdot, sdot broadcast, 0...] (NHWC [N][H][W][C_total]);  mode 1: q (NHWC [N][H][W][C]).
__global__ void __launch_bounds__(256) mbstd_sdot_kernel(const float* __restrict__ x, const float* __restrict__ xdot,
                                                         float* __restrict__ sdot, int G, int M, long long per_sample) {
  const int m = blockIdx.x;
  float acc = 0.f;
  for (long long e = threadIdx.x; e < per_sample; e += blockDim.x) {
    float mean = 0.f;
    for (int g = 0; g < G; ++g) mean += __ldg(x + (long long)(g * M + m) * per_sample + e);
    mean /= (float)G;
    float var = 0.f, dot = 0.f;
    for (int g = 0; g < G; ++g) {
      const float d = __ldg(x + (long long)(g * M + m) * per_sample + e) - mean;
      var += d * d;
      dot += d * __ldg(xdot + (long long)(g * M + m) * per_sample + e);
    }
    acc += dot / ((float)G * sqrtf(var / (float)G + 1e-8f));
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) sdot[m] = red[0] / (float)per_sample;
}

__global__ void __launch_bounds__(256) mbstd_hess_kernel(const float* __restrict__ x, const float* __restrict__ xdot,
                                                         const float* __restrict__ lam, float* __restrict__ q, int G,
                                                         int M, long long per_sample) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)M * per_sample) return;
  const int m = (int)(t / per_sample);
  const long long e = t % per_sample;
  float mean = 0.f, dmean = 0.f;
  for (int g = 0; g < G; ++g) {
    mean += __ldg(x + (long long)(g * M + m) * per_sample + e);
    dmean += __ldg(xdot + (long long)(g * M + m) * per_sample + e);
  }
  mean /= (float)G;
  dmean /= (float)G;
  float var = 0.f, dot = 0.f;
  for (int g = 0; g < G; ++g) {
    const float d = __ldg(x + (long long)(g * M + m) * per_sample + e) - mean;
    var += d * d;
    dot += d * __ldg(xdot + (long long)(g * M + m) * per_sample + e);
  }
  const float sigma = sqrtf(var / (float)G + 1e-8f);
  const float c1 = 1.f / ((float)G * sigma);
  const float c2 = dot / ((float)G * (float)G * sigma * sigma * sigma);
  const float scale = __ldg(lam + m) / (float)per_sample;
  for (int g = 0; g < G; ++g) {
    const long long i = (long long)(g * M + m) * per_sample + e;
    q[i] = scale * ((__ldg(xdot + i) - dmean) * c1 - (__ldg(x + i) - mean) * c2);
  }
}

__global__ void __launch_bounds__(256) mbstd_tangent_concat_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ stat,
                                                                   float* __restrict__ y, long long total, int C,
                                                                   int C_total, int M, long long pix_per_sample) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C_total);
  const long long pix = t / C_total;
  float v = 0.f;
  if (c < C) v = __ldg(x + pix * C + c);
  else if (c == C) v = __ldg(stat + (int)((pix / pix_per_sample) % M));
  y[t] = v;
}

extern "C" int tmx_mbstd_tangent(tmx_handle_t h, const float* x, const float* xdot, float* ydot, float* sdot, int N, int H,
                                 int W, int C, int C_total, int group_size, tmx_stream_t s) {
  TMX_REQUIRE(h && x && xdot && ydot && sdot, TMX_ERR_ARG, "tmx_mbstd_tangent: NULL argument");
  const int G = group_size < N ? group_size : N;
  TMX_REQUIRE(N > 0 && N % G == 0 && C_total > C, TMX_ERR_SHAPE, "tmx_mbstd_tangent: bad shape");
  const int M = N / G;
  const long long per_sample = (long long)H * W * C;
  mbstd_sdot_kernel<<<M, 256, 0, (cudaStream_t)s>>>(x, xdot, sdot, G, M, per_sample);
  TMX_LAUNCHED(h, "mbstd_sdot_kernel");
  const long long total = (long long)N * H * W * C_total;
  mbstd_tangent_concat_kernel<<<tmx_ceil_div(total, 256), 256, 0, (cudaStream_t)s>>>(xdot, sdot, ydot, total, C, C_total,
                                                                                     M, (long long)H * W);
  TMX_LAUNCHED(h, "mbstd_tangent_concat_kernel");
  return TMX_OK;
}

extern "C" int tmx_mbstd_curvature(tmx_handle_t h, const float* x, const float* xdot, const float* lam, float* q, int N,
                                   int H, int W, int C, int group_size, tmx_stream_t s) {
  TMX_REQUIRE(h && x && xdot && lam && q, TMX_ERR_ARG, "tmx_mbstd_curvature: NULL argument");
  const int G = group_size < N ? group_size : N;
  TMX_REQUIRE(N > 0 && N % G == 0, TMX_ERR_SHAPE, "tmx_mbstd_curvature: bad shape");
  const int M = N / G;
  const long long per_sample = (long long)H * W * C;
  mbstd_hess_kernel<<<tmx_ceil_div((long long)M * per_sample, 256), 256, 0, (cudaStream_t)s>>>(x, xdot, lam, q, G, M,
                                                                                              per_sample);
  TMX_LAUNCHED(h, "mbstd_hess_kernel");
  return TMX_OK;
}

// out[n][i] = scale[n] * in[n][i]   (per-sample coefficient of the gradient-penalty tangent seed)
__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ in, const float* __restrict__ scale,
                                                         float* __restrict__ out, long long len, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    out[i] = __ldg(in + i) * __ldg(scale + i / len);
}

// WGAN-GP per-sample terms from the squared gradient norms (loss.py:334-336): pen[n] = lambda (||g_n|| - target)^2 / target^2,
// coef[n] = d mean(pen) / d g_n direction scale = (1/N) 2 lambda (||g_n|| - target) / (target^2 ||g_n||)
__global__ void gp_coef_kernel(const float* __restrict__ sq, float* __restrict__ pen, float* __restrict__ coef, int N,
                               float lambda, float target) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float norm = sqrtf(__ldg(sq + n));
  const float d = norm - target;
  pen[n] = lambda * d * d / (target * target);
  coef[n] = (2.f * lambda * d / (target * target * norm)) / (float)N;
}

extern "C" int tmx_scale_rows(tmx_handle_t h, const float* in, const float* scale, float* out, int rows, int64_t len,
                              tmx_stream_t s) {
  TMX_REQUIRE(h && in && scale && out && rows > 0 && len > 0, TMX_ERR_ARG, "tmx_scale_rows: bad argument");
  const long long total = (long long)rows * len;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)h->sm_count * 8) blocks = (long long)h->sm_count * 8;
  scale_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(in, scale, out, len, total);
  TMX_LAUNCHED(h, "scale_rows_kernel");
  return TMX_OK;
}

extern "C" int tmx_gp_coefficients(tmx_handle_t h, const float* sq_norms, float* penalty, float* coef, int N, float lambda,
                                   float target, tmx_stream_t s) {
  TMX_REQUIRE(h && sq_norms && penalty && coef && N > 0, TMX_ERR_ARG, "tmx_gp_coefficients: bad argument");
  gp_coef_kernel<<<tmx_ceil_div(N, 128), 128, 0, (cudaStream_t)s>>>(sq_norms, penalty, coef, N, lambda, target);
  TMX_LAUNCHED(h, "gp_coef_kernel");
  return TMX_OK;
}

// out = a * in + b  (seeds of the critic losses: d loss / d score)
__global__ void __launch_bounds__(256) axpb_kernel(const float* __restrict__ in, float* __restrict__ out, long long n,
                                                   float a, float b) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaf(__ldg(in + i), a, b);
}

extern "C" int tmx_axpb(tmx_handle_t h, const float* in, float* out, int64_t n, float a, float b, tmx_stream_t s) {
  TMX_REQUIRE(h && in && out && n > 0, TMX_ERR_ARG, "tmx_axpb: bad argument");
  axpb_kernel<<<tmx_ceil_div(n, 256), 256, 0, (cudaStream_t)s>>>(in, out, n, a, b);
  TMX_LAUNCHED(h, "axpb_kernel");
  return TMX_OK;
}

// out = a + b (fp32): two gradient contributions meeting at one tensor (e.g. critic gradient + L1 gradient)
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                  float* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __ldg(a + i) + __ldg(b + i);
}

extern "C" int tmx_add_f32(tmx_handle_t h, const float* a, const float* b, float* out, int64_t n, tmx_stream_t s) {
  TMX_REQUIRE(h && a && b && out && n > 0, TMX_ERR_ARG, "tmx_add_f32: bad argument");
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)h->sm_count * 8) blocks = (long long)h->sm_count * 8;
  add_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(a, b, out, n);
  TMX_LAUNCHED(h, "add_kernel");
  return TMX_OK;
}

int tmx_conv2d_dgrad_lin_patch(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, const uint16_t* dz_hi,
                               const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                               cudaStream_t st, int* served, const tmx_grad_desc_t* gd = nullptr,
                               const tmx_grad_io_t* gio = nullptr);
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
  if (k == 3 && dz_hi && dz_lo && wt_hi && wt_lo && g_f32 && N > 0 && H > 0 && W > 0) {
    // thin layers (Cin, Cout in {16, 32, 64}): one haloed patch per tile instead of nine tap boxes (conv_lin.cu)
    int served = 0;
    const int rc = tmx_conv2d_dgrad_lin_patch(h, N, H, W, Cin, Cout, dz_hi, dz_lo, wt_hi, wt_lo, g_f32, (cudaStream_t)s,
                                              &served);
    if (rc != TMX_OK || served) return rc;
  }
  return tmx_conv2d_dgrad_tc(h, N, H, W, Cin, Cout, k, dz_hi, dz_lo, wt_hi, wt_lo, g_f32, (cudaStream_t)s);
}

// ---------------------------------------------------------------- fused data gradient + grad_prepare (GP)
// tmx_conv2d_dgrad_gp = tmx_conv2d_dgrad followed by tmx_grad_prepare(src_kind 0) with the second one folded into
// the epilogue of the first (conv_tc.cu, GP) for every interior pixel that receives no folded ring value; the border
// kernel below finishes the others - rows / columns {1, H-2} (REFLECT) or {0, H-1} (REPLICATE) - with the arithmetic
// of grad_prepare_kernel in the same order, so planes and fp32 output are bit-identical to the two-kernel path (the
// bias gradient is summed in another order).
int tmx_conv2d_dgrad_gp_tc(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                           const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                           const tmx_grad_desc_t* gd, const tmx_grad_io_t* gio, cudaStream_t st, int* served);

__global__ void __launch_bounds__(256) grad_border_kernel(const GradPrepParams P, int D) {
  extern __shared__ float red[];   // [blockDim / C8][C] bias-gradient partials
  const tmx_grad_desc_t& d = P.d;
  const int C8 = P.C8, C = d.C, H = d.H, W = d.W;
  const int Hq = H + 4, Wq = W + 4;
  const int cg = threadIdx.x % C8, co = cg * 8;
  const int lane_px = threadIdx.x / C8, lanes = blockDim.x / C8;
  const long long pixels = (long long)d.N * D;
  float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int e0 = d.fold == 0 ? 1 : 0;          // first folded row / column; the other one is H-1-e0 / W-1-e0
  for (long long px = (long long)blockIdx.x * lanes + lane_px; px < pixels; px += (long long)gridDim.x * lanes) {
    const int n = (int)(px / D);
    const int q = (int)(px - (long long)n * D);
    int r, c;
    if (q < W) {
      r = e0;
      c = q;
    } else if (q < 2 * W) {
      r = H - 1 - e0;
      c = q - W;
    } else {
      const int e = q - 2 * W, ri = e >> 1;       // ri-th row that is not a folded row
      r = d.fold == 0 ? (ri == 0 ? 0 : (ri <= H - 4 ? ri + 1 : H - 1)) : ri + 1;
      c = (e & 1) ? W - 1 - e0 : e0;
    }
    int rs[2], cs[2], nr = 0, nc = 0;
    rs[nr++] = r + 2;
    cs[nc++] = c + 2;
    if (d.fold == 0) {
      if (r == 1) rs[nr++] = 1;
      if (r == H - 2) rs[nr++] = H + 2;
      if (c == 1) cs[nc++] = 1;
      if (c == W - 2) cs[nc++] = W + 2;
    } else {
      if (r == 0) rs[nr++] = 1;
      if (r == H - 1) rs[nr++] = H + 2;
      if (c == 0) cs[nc++] = 1;
      if (c == W - 1) cs[nc++] = W + 2;
    }
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < nr; ++a)
      for (int b = 0; b < nc; ++b) {
        float t[8];
        ld8(P.io.g + (((long long)n * Hq + rs[a]) * Wq + cs[b]) * C + co, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
      }
    const long long in_pix = ((long long)n * H + r) * W + c;
    if (P.io.add != nullptr) {
      float t[8];
      ld8(P.io.add + in_pix * C + co, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    if (d.mask_kind == 1) {
      float y[8];
      ld8(reinterpret_cast<const float*>(P.io.y_mask) + in_pix * C + co, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= (y[j] > 0.f) ? 1.f : d.alpha;
    } else if (d.mask_kind == 2) {
      const uint4 yb = __ldg(reinterpret_cast<const uint4*>(
          reinterpret_cast<const uint16_t*>(P.io.y_mask) + (((long long)n * (H + 2) + r + 1) * (W + 2) + c + 1) * C + co));
      const uint32_t w4[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[2 * j] *= ((int)(w4[j] << 16) > 0) ? 1.f : d.alpha;
        v[2 * j + 1] *= ((int)(w4[j] & 0xffff0000u) > 0) ? 1.f : d.alpha;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] += v[j];
    if (P.io.dz_f32 != nullptr) {
      float4* o = reinterpret_cast<float4*>(P.io.dz_f32 + in_pix * C + co);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (P.io.dz_hi != nullptr) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tmx_split_bf16x2(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
      const long long o = (((long long)n * Hq + r + 2) * Wq + c + 2) * C + co;
      *reinterpret_cast<uint4*>(P.io.dz_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(P.io.dz_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
  if (P.io.dbias != nullptr) {
    float* mine = red + lane_px * C + co;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = bsum[j];
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * C + ch];
      if (s != 0.f) atomicAdd(P.io.dbias + ch, s * d.dbias_scale);
    }
  }
}

extern "C" int tmx_conv2d_dgrad_gp(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                                   const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                                   const tmx_grad_desc_t* gd, const tmx_grad_io_t* gio, int* served, tmx_stream_t s) {
  TMX_REQUIRE(h && dz_hi && dz_lo && wt_hi && wt_lo && g_f32 && gd && gio && served, TMX_ERR_ARG,
              "tmx_conv2d_dgrad_gp: NULL argument");
  TMX_REQUIRE(gd->N == N && gd->H == H && gd->W == W && gd->C == Cin && gd->src_kind == 0 && !gd->phase_pack &&
                  gd->fold >= 0 && gd->fold <= 2 && gd->mask_kind >= 0 && gd->mask_kind <= 2,
              TMX_ERR_ARG, "tmx_conv2d_dgrad_gp: the grad descriptor must describe the [N][H][W][Cin] input of this layer "
              "(src_kind 0, no phase_pack)");
  TMX_REQUIRE((gio->dz_hi == nullptr) == (gio->dz_lo == nullptr), TMX_ERR_ARG, "tmx_conv2d_dgrad_gp: dz_hi/dz_lo go together");
  TMX_REQUIRE(gio->dz_hi || gio->dz_f32, TMX_ERR_ARG, "tmx_conv2d_dgrad_gp: no output (dz planes and / or dz_f32)");
  TMX_REQUIRE(gd->mask_kind == 0 || gio->y_mask, TMX_ERR_ARG, "tmx_conv2d_dgrad_gp: mask requested without y_mask");
  TMX_REQUIRE((k == 1 || k == 3) && N > 0 && H > 0 && W > 0, TMX_ERR_SHAPE, "tmx_conv2d_dgrad_gp: bad shape");
  const void* ptrs[] = {dz_hi, dz_lo, wt_hi, wt_lo, g_f32, gio->add, gio->y_mask, gio->dz_hi, gio->dz_lo, gio->dz_f32,
                        gio->dbias};
  for (const void* q : ptrs)
    TMX_REQUIRE(((uintptr_t)q & 15) == 0, TMX_ERR_ARG, "tmx_conv2d_dgrad_gp: buffers must be 16-byte aligned (%p)", q);
  *served = 0;
  if (tmx_env_flag("TMX_NO_FUSED_GP") || (gd->fold != 2 && (H < 4 || W < 4))) return TMX_OK;
  int rc = TMX_OK;
  if (k == 3)      // thin layers (Cin, Cout in {16, 32, 64}): the LIN-PATCH kernel with the same epilogue
    rc = tmx_conv2d_dgrad_lin_patch(h, N, H, W, Cin, Cout, dz_hi, dz_lo, wt_hi, wt_lo, g_f32, (cudaStream_t)s, served, gd,
                                    gio);
  if (rc == TMX_OK && !*served)
    rc = tmx_conv2d_dgrad_gp_tc(h, N, H, W, Cin, Cout, k, dz_hi, dz_lo, wt_hi, wt_lo, g_f32, gd, gio, (cudaStream_t)s,
                                served);
  if (rc != TMX_OK || !*served || gd->fold == 2) return rc;
  GradPrepParams P;
  P.d = *gd;
  P.io = *gio;
  P.io.g = g_f32;
  P.C8 = Cin / 8;
  P.ppb = 256 / P.C8 > 0 ? 256 / P.C8 : 1;
  P.rpb = 0;
  P.c8_shift = -1;
  const int threads = P.ppb * P.C8;
  const int D = 2 * W + 2 * (H - 2);
  const long long pixels = (long long)N * D;
  long long blocks = (pixels + P.ppb - 1) / P.ppb;
  if (blocks > 4LL * h->sm_count) blocks = 4LL * h->sm_count;
  const size_t smem = gio->dbias ? (size_t)threads * 8 * sizeof(float) : 0;
  TMX_REQUIRE(threads <= 256 && smem <= 48 * 1024, TMX_ERR_SHAPE, "tmx_conv2d_dgrad_gp: C=%d not supported", Cin);
  grad_border_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)s>>>(P, D);
  TMX_LAUNCHED(h, "grad_border_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- KL regulariser of EG_wgan (loss.py:163-171)
// KL = -0.5 * kl_weight * mean(1 + 2 ls - mu^2 - exp(2 ls)) per sample; with gscale = kl_weight / (elements per
// sample * batch) the gradient of the batch mean is dmu = gscale * mu, dls = gscale * (exp(2 ls) - 1);
// val = 1 + 2 ls - mu^2 - exp(2 ls) per element (the caller reduces it with tmx_row_sum).
__global__ void __launch_bounds__(256) kl_terms_kernel(const float* __restrict__ mu, const float* __restrict__ ls,
                                                       float* __restrict__ dmu, float* __restrict__ dls,
                                                       float* __restrict__ val, long long n, float gscale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float m = __ldg(mu + i), l = __ldg(ls + i);
    const float e = expf(2.f * l);
    val[i] = 1.f + 2.f * l - m * m - e;
    dmu[i] = gscale * m;
    dls[i] = gscale * (e - 1.f);
  }
}

extern "C" int tmx_kl_terms(tmx_handle_t h, const float* mu, const float* log_sigma, float* dmu, float* dls, float* val,
                            int64_t n, float gscale, tmx_stream_t s) {
  TMX_REQUIRE(h && mu && log_sigma && dmu && dls && val && n > 0, TMX_ERR_ARG, "tmx_kl_terms: bad argument");
  const long long cap = (long long)h->sm_count * 16;
  const int grid = (int)(n / 256 + 1 < cap ? n / 256 + 1 : cap);
  kl_terms_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(mu, log_sigma, dmu, dls, val, n, gscale);
  TMX_LAUNCHED(h, "kl_terms_kernel");
  return TMX_OK;
}
