// conv_wgrad.cu — weight gradient of the 3x3 / 1x1 convolution on tcgen05:
//
//   dw[u][v][ci][co] = wscale * sum_{n,y,x} xpad[n][y+u][x+v][ci] * dz[n][y][x][co]
//
// Per tap (u,v) this is a GEMM  D[ci][co] += X^T[ci][pix] * DZ[pix][co]  whose contraction runs over PIXELS.
// Both operands are stored channel-contiguous (NHWC planes), i.e. they are MN-major for this GEMM: a TMA box
// [64 channels][64 pixels] lands in shared memory as 64 rows (pixels = K) of 128 B (channels = M or N), which
// is exactly the canonical MN-major SWIZZLE_128B operand layout of tcgen05.mma (8-row atoms, SBO = 1024 B,
// 64-channel column blocks LBO apart) - no transposition pass.  x comes from the forward input planes
// (SPLIT_BF16_HALO, shifted by the tap), dz from the zero-ringed gradient planes of tmx_grad_prepare.
// bf16x3 like the forward (x_lo*dz_hi + x_hi*dz_lo + x_hi*dz_hi).
//
// Channel counts that are not multiples of 64 are handled by the TMA unit: boxes reach past the channel extent
// and read zeros, so the thin layers (16/32 channels) run the same kernel with idle rows / columns.
//
// PACKED-M mode (3x3, Cin <= 64): a thin layer would leave most of the 128 UMMA rows (and of every 128-B TMA row)
// empty and pay the TMA row rate nine times per pixel - measured: the 16/32/64-channel layers at 128^2..256^2 took
// as long as the 256-channel trunk.  Instead the 128 rows of one MMA carry TWO 64-row blocks, each block = one
// vertical tap u and xwin = 64/Cin consecutive horizontal taps: block row m = dx*Cin + ci reads x[pix + dx] through
// an OVERLAPPING tensor map (innermost extent xwin*Cin elements, pixel stride Cin: one full 128-B row per pixel
// carries the horizontal neighbours too).  3*ceil(3/xwin) blocks -> 2 / 3 / 5 "tap groups" for Cin = 16 / 32 / 64
// instead of 9 taps: 4.5x / 3x / 1.8x fewer TMA rows and MMAs.  Rows whose horizontal tap would be >= 3 are junk
// (they read the next pixels / up to 96 B behind the plane - the caller vouches for that slack) and are dropped by
// the reduction.
//
// Work split: an output tile is (tap, 128 input channels, BN output channels); the pixel range is cut into
// `splits` slices so that tiles x splits fills the SMs (split-K).  Each CTA accumulates its slice in TMEM and
// writes a partial tile to the workspace; tmx_conv2d_wgrad then reduces the slices in a fixed order
// (deterministic) and ACCUMULATES wscale * sum into dw, so that several losses can add into one gradient buffer.
//
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#include "tc_common.cuh"

namespace {

constexpr int kWM = 128;      // input channels per tile (UMMA M)
constexpr int kWKP = 64;      // pixels per pipeline stage (4 UMMA K steps)
constexpr int kWThreads = 256;
constexpr int kBlk = 64 * kWKP * 2;   // one [64 ch][64 px] bf16 box = 8 KB

struct WgradParams {
  int N, H, W, Cin, Cout, k, taps;
  int bw, bh, bn;                  // pixel patch of one stage: bw*bh*bn == 64
  int tiles_x, tiles_y, tiles_n;   // patches per image row / column / batch
  int chunks;                      // total pixel patches
  int ci_tiles, co_tiles;          // Cin/128, Cout/BN, both rounded up (channels past the end read zeros)
  int co_pad;                      // co_tiles*BN: row length of a partial tile row
  int splits, chunks_per_split;
  float* partial;                  // [splits][taps][ci_tiles*128][Cout]
  int packed, xwin, bpu, nblocks;  // PACKED-M mode: taps == tap groups; blocks per vertical tap, total blocks
  // GRAM mode (tmx_gram_fwd_tc): the "tap" index is the SAMPLE, the pixel chunks stay inside that sample, and both
  // operands are the same feature planes (halo layout, interior at offset 1): partial[split][n] = F[n]^T F[n]
  int gram, zoff;                  // zoff: interior offset of the B operand's grid (2: zero-ringed dz, 1: halo planes)
};

// MN-major SWIZZLE_128B operand: 8 K-rows of 128 B per atom (SBO = 1024 B), 64-channel blocks LBO apart.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D f32, A/B bf16, BOTH operands MN-major (bits 15, 16), N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kWM >> 4) << 24);
}

template <int BN>
struct WCfg {
  static constexpr int kABytes = 2 * kBlk;            // 128 channels = two 64-channel blocks, per plane
  static constexpr int kBBytes = (BN / 64) * kBlk;    // per plane
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  // STACK (BN <= 128): two MMAs per K step instead of three - the dz tile's hi and lo column blocks are consecutive
  // 64-channel blocks LBO apart, so one MMA with N = 2*BN gives x_hi*dz_hi and x_hi*dz_lo; a second one x_lo*dz_hi;
  // 3*BN accumulator columns, added by the epilogue (small-N MMAs are bound by the A-operand read, see conv_lin.cu)
  static constexpr bool kStack = BN <= 128;
  static constexpr int kAccCols = kStack ? 3 * BN : BN;
  static constexpr int kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 1024;
  static_assert(kStages >= 2, "need at least a double buffer");
};

template <int BN>
__global__ void __launch_bounds__(kWThreads, 1)
    conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                      const __grid_constant__ CUtensorMap tm_z_hi, const __grid_constant__ CUtensorMap tm_z_lo,
                      const WgradParams p) {
  using Cfg = WCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* done_bar = empty_bar + S;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  tmx_pdl_trigger();

  // work item of this CTA: (split, tap, ci tile, co tile)
  int item = blockIdx.x;
  const int co_t = item % p.co_tiles;
  item /= p.co_tiles;
  const int ci_t = item % p.ci_tiles;
  item /= p.ci_tiles;
  const int tap = item % p.taps;
  const int split = item / p.taps;
  const int u = p.gram ? 0 : tap / p.k, v = p.gram ? 0 : tap - u * p.k;
  const int pad_off = 1 - p.k / 2;
  const int n_base = p.gram ? tap : 0;
  const int chunk0 = split * p.chunks_per_split;
  const int chunk1 = min(p.chunks, chunk0 + p.chunks_per_split);
  const int nchunks = max(0, chunk1 - chunk0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x_hi);
    tma_prefetch_desc(&tm_x_lo);
    tma_prefetch_desc(&tm_z_hi);
    tma_prefetch_desc(&tm_z_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_s);
  tmx_pdl_wait();      // set-up above overlaps the previous kernel's tail; its results are read from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = chunk0; c < chunk1; ++c) {
        int t = c;
        const int x0 = (t % p.tiles_x) * p.bw;
        t /= p.tiles_x;
        const int y0 = (t % p.tiles_y) * p.bh;
        const int n0 = (t / p.tiles_y) * p.bn + n_base;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        // A: x planes, halo layout [N][H+2][W+2][Cin], shifted by the tap; channels beyond Cin read zeros
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          int c0 = ci_t * kWM + b * 64, xs = x0 + v + pad_off, ys = y0 + u + pad_off;
          if (p.packed) {       // block bi = (vertical tap, first horizontal tap) through the overlapping-row map
            const int bi = min(2 * tap + b, p.nblocks - 1);
            c0 = 0;
            xs = x0 + (bi % p.bpu) * p.xwin;
            ys = y0 + bi / p.bpu;
          }
          tma_load_4d(sa + b * kBlk, &tm_x_hi, &full_bar[stage], c0, xs, ys, n0);
          tma_load_4d(sa + Cfg::kABytes + b * kBlk, &tm_x_lo, &full_bar[stage], c0, xs, ys, n0);
        }
        // B: dz planes on the zero-ringed grid [N][H+4][W+4][Cout], interior at offset 2
#pragma unroll
        for (int b = 0; b < BN / 64; ++b) {
          const int c0 = co_t * BN + b * 64;
          tma_load_4d(sa + 2 * Cfg::kABytes + b * kBlk, &tm_z_hi, &full_bar[stage], c0, x0 + p.zoff, y0 + p.zoff, n0);
          tma_load_4d(sa + 2 * Cfg::kABytes + Cfg::kBBytes + b * kBlk, &tm_z_lo, &full_bar[stage], c0, x0 + p.zoff,
                      y0 + p.zoff, n0);
        }
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_mn(BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
#pragma unroll
        for (int kk = 0; kk < kWKP / 16; ++kk) {
          const uint32_t koff = kk * 16 * 128;   // 16 pixel rows of 128 B
          const uint64_t a_hi = make_desc_mn(sa + koff, kBlk);
          const uint64_t a_lo = make_desc_mn(sa + Cfg::kABytes + koff, kBlk);
          const uint64_t b_hi = make_desc_mn(sa + 2 * Cfg::kABytes + koff, kBlk);
          const uint64_t b_lo = make_desc_mn(sa + 2 * Cfg::kABytes + Cfg::kBBytes + koff, kBlk);
          if (Cfg::kStack) {
            umma_bf16(tmem_base + 2 * BN, a_lo, b_hi, idesc, (c | kk) != 0);
            umma_bf16(tmem_base, a_hi, b_hi, make_idesc_mn(2 * BN), (c | kk) != 0);   // [dz_hi | dz_lo] blocks
          } else {
            umma_bf16(tmem_base, a_lo, b_hi, idesc, (c | kk) != 0);
            umma_bf16(tmem_base, a_hi, b_lo, idesc, 1);
            umma_bf16(tmem_base, a_hi, b_hi, idesc, 1);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: partial tile -> workspace =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                 // input channel within the tile
    const int ci = ci_t * kWM + row;
    const int cin_pad = p.ci_tiles * kWM;
    float* dst = p.partial + (((long long)split * p.taps + tap) * cin_pad + ci) * p.co_pad + co_t * BN;
    if (nchunks > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int g = 0; g < BN / 32; ++g) {
      uint32_t acc[32];
      if (nchunks > 0 && Cfg::kStack) {
        uint32_t tmp[32];
        tmem_ld32(taddr0 + 2 * BN + g * 32, acc);
        tmem_ld32(taddr0 + BN + g * 32, tmp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(tmp[j]));
        tmem_ld32(taddr0 + g * 32, tmp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(tmp[j]));
      } else if (nchunks > 0) {
        tmem_ld32(taddr0 + g * 32, acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0u;
      }
      float4* o = reinterpret_cast<float4*>(dst + g * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                           __uint_as_float(acc[4 * j + 3]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// dw[tap][ci][co] += scale * sum_s partial[s][tap][ci][co]   (fixed order -> deterministic)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                           int splits, int taps, int Cin, int cin_pad, int Cout,
                                                           int co_pad, float scale, int overwrite) {
  tmx_pdl_trigger();
  tmx_pdl_wait();
  const long long total4 = (long long)taps * Cin * Cout / 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total4) return;
  const long long e = t * 4;
  const int co = (int)(e % Cout);
  const long long q = e / Cout;
  const int ci = (int)(q % Cin);
  const int tap = (int)(q / Cin);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(
        partial + (((long long)s * taps + tap) * cin_pad + ci) * co_pad + co));
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  float4* o = reinterpret_cast<float4*>(dw + e);
  float4 cur = overwrite ? make_float4(0.f, 0.f, 0.f, 0.f) : *o;
  cur.x = fmaf(acc.x, scale, cur.x);
  cur.y = fmaf(acc.y, scale, cur.y);
  cur.z = fmaf(acc.z, scale, cur.z);
  cur.w = fmaf(acc.w, scale, cur.w);
  *o = cur;
}

// Many-slice form of the two reductions (thin layers: few outputs, up to ~70 pixel slices): one WARP per float4 of
// outputs, the lanes stride over the slices and a fixed xor-tree adds them (deterministic), lane 0 accumulates into dw.
// The one-thread-per-output kernels above serialise the slices: 17 us per launch for a 16 -> 16 layer.
template <bool PACKED>
__global__ void __launch_bounds__(256) wgrad_reduce_warp_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                int splits, int taps, int Cin, int cin_pad, int Cout,
                                                                int co_pad, int xwin, int bpu, float scale,
                                                                int overwrite) {
  tmx_pdl_trigger();
  tmx_pdl_wait();
  const long long total4 = (long long)(PACKED ? 9 : taps) * Cin * Cout / 4;
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= total4) return;
  const long long e = t * 4;
  const int co = (int)(e % Cout);
  const long long q = e / Cout;
  const int ci = (int)(q % Cin);
  const int tap = (int)(q / Cin);
  long long base, sstride;
  if (PACKED) {
    const int u = tap / 3, v = tap - u * 3;
    const int bi = u * bpu + v / xwin;
    const int row = (bi & 1) * 64 + (v % xwin) * Cin + ci;
    base = ((long long)(bi >> 1) * kWM + row) * co_pad + co;
    sstride = (long long)taps * kWM * co_pad;          // taps == tap groups
  } else {
    base = ((long long)tap * cin_pad + ci) * co_pad + co;
    sstride = (long long)taps * cin_pad * co_pad;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = lane; s < splits; s += 32) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(partial + s * sstride + base));
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, sft);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, sft);
  }
  if (lane == 0) {
    float4* o = reinterpret_cast<float4*>(dw + e);
    float4 cur = overwrite ? make_float4(0.f, 0.f, 0.f, 0.f) : *o;
    cur.x = fmaf(acc.x, scale, cur.x);
    cur.y = fmaf(acc.y, scale, cur.y);
    cur.z = fmaf(acc.z, scale, cur.z);
    cur.w = fmaf(acc.w, scale, cur.w);
    *o = cur;
  }
}

// PACKED-M reduction: dw[u][v][ci][co] += scale * sum_s partial[s][group][row][co] with
// block bi = u*bpu + v/xwin, group = bi/2, row = (bi&1)*64 + (v%xwin)*Cin + ci
__global__ void __launch_bounds__(256) wgrad_reduce_packed_kernel(const float* __restrict__ partial,
                                                                  float* __restrict__ dw, int splits, int groups, int Cin,
                                                                  int Cout, int co_pad, int xwin, int bpu, float scale) {
  tmx_pdl_trigger();
  tmx_pdl_wait();
  const long long total4 = 9LL * Cin * Cout / 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total4) return;
  const long long e = t * 4;
  const int co = (int)(e % Cout);
  const long long q = e / Cout;
  const int ci = (int)(q % Cin);
  const int tap = (int)(q / Cin);
  const int u = tap / 3, v = tap - u * 3;
  const int bi = u * bpu + v / xwin;
  const int row = (bi & 1) * 64 + (v % xwin) * Cin + ci;
  const int group = bi >> 1;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(
        partial + (((long long)s * groups + group) * kWM + row) * co_pad + co));
    acc.x += x.x;
    acc.y += x.y;
    acc.z += x.z;
    acc.w += x.w;
  }
  float4* o = reinterpret_cast<float4*>(dw + e);
  float4 cur = *o;
  cur.x = fmaf(acc.x, scale, cur.x);
  cur.y = fmaf(acc.y, scale, cur.y);
  cur.z = fmaf(acc.z, scale, cur.z);
  cur.w = fmaf(acc.w, scale, cur.w);
  *o = cur;
}

int encode_map4(tmx_handle_t h, CUtensorMap* m, const uint16_t* base, int N, int Hp, int Wp, int C, int bw, int bh,
                int bn, int xwin = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)C * xwin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = h->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tmx_fail(TMX_ERR_DRIVER, "cuTensorMapEncodeTiled(wgrad) failed: CUresult %d (C=%d Wp=%d Hp=%d N=%d)", (int)r,
                    C, Wp, Hp, N);
  return TMX_OK;
}

int pow2_div(int v, int cap) {
  int g = 1;
  while (g < cap && (v % (g * 2)) == 0) g *= 2;
  return g;
}

bool wgrad_packed(int Cin, int k, int flags) {
  if (tmx_env_flag("TMX_NO_WGRAD_PACK") || k != 3) return false;
  if (Cin == 64) return true;                                    // plain map: no slack needed
  return (Cin == 16 || Cin == 32) && (flags & TMX_WGRAD_X_SLACK) != 0;
}

void wgrad_plan(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, int bn_cols, int flags, WgradParams& p) {
  p.N = N;
  p.H = H;
  p.W = W;
  p.Cin = Cin;
  p.Cout = Cout;
  p.k = k;
  p.taps = k * k;
  p.bw = pow2_div(W, 32);
  p.bh = pow2_div(H, kWKP / p.bw);
  p.bn = kWKP / (p.bw * p.bh);
  p.tiles_x = W / p.bw;
  p.tiles_y = H / p.bh;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.chunks = p.tiles_x * p.tiles_y * p.tiles_n;
  p.ci_tiles = (Cin + kWM - 1) / kWM;
  p.co_tiles = (Cout + bn_cols - 1) / bn_cols;
  p.co_pad = p.co_tiles * bn_cols;
  p.gram = 0;
  p.zoff = 2;
  p.packed = wgrad_packed(Cin, k, flags) ? 1 : 0;
  p.xwin = p.bpu = p.nblocks = 1;
  if (p.packed) {
    p.xwin = 64 / Cin;
    p.bpu = (3 + p.xwin - 1) / p.xwin;
    p.nblocks = 3 * p.bpu;
    p.taps = (p.nblocks + 1) / 2;          // tap groups of two 64-row blocks
  }
  const int tiles = p.taps * p.ci_tiles * p.co_tiles;
  int splits = h->sm_count / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.chunks) splits = p.chunks;
  p.chunks_per_split = (p.chunks + splits - 1) / splits;
  p.splits = (p.chunks + p.chunks_per_split - 1) / p.chunks_per_split;
}

int wgrad_bn(int Cout) { return Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64); }

template <int BN>
int launch_wgrad(tmx_handle_t h, const CUtensorMap* maps, const WgradParams& p, cudaStream_t st) {
  using Cfg = WCfg<BN>;
  auto kern = conv_wgrad_kernel<BN>;
  static thread_local int configured_device = -1;
  if (configured_device != h->device) {
    TMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured_device = h->device;
  }
  const int grid = p.splits * p.taps * p.ci_tiles * p.co_tiles;
  TMX_CUDA(tmx_launch_pdl(kern, dim3(grid), dim3(kWThreads), (size_t)Cfg::kSmemBytes, st, 1, maps[0], maps[1], maps[2],
                          maps[3], p));
  TMX_LAUNCHED(h, "conv_wgrad_kernel");
  return TMX_OK;
}

}  // namespace

extern "C" int tmx_conv2d_wgrad_workspace_bytes(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, int flags,
                                                size_t* bytes) {
  TMX_REQUIRE(h && bytes, TMX_ERR_ARG, "tmx_conv2d_wgrad_workspace_bytes: NULL argument");
  TMX_REQUIRE((k == 1 || k == 3) && N > 0 && H > 0 && W > 0 && Cin % 8 == 0 && Cout % 8 == 0, TMX_ERR_SHAPE,
              "tmx_conv2d_wgrad: needs k in {1,3}, Cin and Cout multiples of 8 (got k=%d Cin=%d Cout=%d)", k, Cin, Cout);
  WgradParams p;
  wgrad_plan(h, N, H, W, Cin, Cout, k, wgrad_bn(Cout), flags, p);
  *bytes = (size_t)p.splits * p.taps * p.ci_tiles * kWM * p.co_pad * sizeof(float);
  return TMX_OK;
}

extern "C" int tmx_conv2d_wgrad(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, float wscale,
                                const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* dz_hi,
                                const uint16_t* dz_lo, float* dw, float* workspace, int flags, tmx_stream_t s) {
  TMX_REQUIRE(h && x_hi && x_lo && dz_hi && dz_lo && dw && workspace, TMX_ERR_ARG, "tmx_conv2d_wgrad: NULL argument");
  TMX_REQUIRE((k == 1 || k == 3) && N > 0 && H >= 2 && W >= 2 && Cin % 8 == 0 && Cout % 8 == 0, TMX_ERR_SHAPE,
              "tmx_conv2d_wgrad: needs k in {1,3}, H, W >= 2, Cin and Cout multiples of 8 (got k=%d %dx%d Cin=%d "
              "Cout=%d)", k, H, W, Cin, Cout);
  const void* ptrs[] = {x_hi, x_lo, dz_hi, dz_lo, dw, workspace};
  for (const void* q : ptrs)
    TMX_REQUIRE(((uintptr_t)q & 15) == 0, TMX_ERR_ARG, "tmx_conv2d_wgrad: buffers must be 16-byte aligned (%p)", q);
  const int bn_cols = wgrad_bn(Cout);
  WgradParams p;
  wgrad_plan(h, N, H, W, Cin, Cout, k, bn_cols, flags, p);
  p.partial = workspace;
  CUtensorMap maps[4];
  int rc;
  if ((rc = encode_map4(h, &maps[0], x_hi, N, H + 2, W + 2, Cin, p.bw, p.bh, p.bn, p.xwin))) return rc;
  if ((rc = encode_map4(h, &maps[1], x_lo, N, H + 2, W + 2, Cin, p.bw, p.bh, p.bn, p.xwin))) return rc;
  if ((rc = encode_map4(h, &maps[2], dz_hi, N, H + 4, W + 4, Cout, p.bw, p.bh, p.bn))) return rc;
  if ((rc = encode_map4(h, &maps[3], dz_lo, N, H + 4, W + 4, Cout, p.bw, p.bh, p.bn))) return rc;
  cudaStream_t st = (cudaStream_t)s;
  if (bn_cols == 256) rc = launch_wgrad<256>(h, maps, p, st);
  else if (bn_cols == 128) rc = launch_wgrad<128>(h, maps, p, st);
  else rc = launch_wgrad<64>(h, maps, p, st);
  if (rc) return rc;
  const long long total4 = (long long)k * k * Cin * Cout / 4;
  // thin layers (few outputs, many pixel slices): slices in parallel, one warp per float4 of outputs
  const bool many = p.splits >= 8 && total4 <= 16384;
  const float* ws_c = workspace;
  if (p.packed && many)
    TMX_CUDA(tmx_launch_pdl(wgrad_reduce_warp_kernel<true>, dim3(tmx_ceil_div(total4 * 32, 256)), dim3(256), 0, st, 1,
                            ws_c, dw, p.splits, p.taps, Cin, p.ci_tiles * kWM, Cout, p.co_pad, p.xwin, p.bpu, wscale, 0));
  else if (p.packed)
    TMX_CUDA(tmx_launch_pdl(wgrad_reduce_packed_kernel, dim3(tmx_ceil_div(total4, 256)), dim3(256), 0, st, 1, ws_c, dw,
                            p.splits, p.taps, Cin, Cout, p.co_pad, p.xwin, p.bpu, wscale));
  else if (many)
    TMX_CUDA(tmx_launch_pdl(wgrad_reduce_warp_kernel<false>, dim3(tmx_ceil_div(total4 * 32, 256)), dim3(256), 0, st, 1,
                            ws_c, dw, p.splits, p.taps, Cin, p.ci_tiles * kWM, Cout, p.co_pad, 1, 1, wscale, 0));
  else
    TMX_CUDA(tmx_launch_pdl(wgrad_reduce_kernel, dim3(tmx_ceil_div(total4, 256)), dim3(256), 0, st, 1, ws_c, dw, p.splits,
                            p.taps, Cin, p.ci_tiles * kWM, Cout, p.co_pad, wscale, 0));
  TMX_LAUNCHED(h, "wgrad_reduce_kernel");
  return TMX_OK;
}

// ---------------------------------------------------------------- Gram matrices on the tensor cores (loss.py:29-35)
// G[n][i][j] = sum_p F[n][p][i] F[n][p][j] / (H W) from the split-bf16 feature planes [N][H+2][W+2][C] the VGG conv
// epilogue wrote: the weight-gradient kernel with k = 1 where the "tap" is the sample and both operands are F.
namespace {
bool gram_plan(tmx_handle_t h, int N, int C, int H, int W, int bn_cols, WgradParams& p) {
  wgrad_plan(h, 1, H, W, C, C, 1, bn_cols, 0, p);
  if (p.bn != 1) return false;             // a 64-pixel chunk must lie inside one sample
  p.N = N;
  p.gram = 1;
  p.zoff = 1;
  p.taps = N;
  const int tiles = p.taps * p.ci_tiles * p.co_tiles;
  int splits = h->sm_count / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.chunks) splits = p.chunks;
  p.chunks_per_split = (p.chunks + splits - 1) / splits;
  p.splits = (p.chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  return true;
}
}  // namespace

extern "C" int tmx_gram_fwd_tc_workspace_bytes(tmx_handle_t h, int N, int C, int H, int W, size_t* bytes) {
  TMX_REQUIRE(h && bytes && N > 0 && C > 0 && H > 0 && W > 0, TMX_ERR_ARG, "tmx_gram_fwd_tc_workspace_bytes: bad argument");
  WgradParams p;
  // 0 bytes = this shape is not served by the tensor-core path (use tmx_gram_fwd on the NCHW features)
  if (C % 8 != 0 || (H * W) % kWKP != 0 || !gram_plan(h, N, C, H, W, wgrad_bn(C), p)) {
    *bytes = 0;
    return TMX_OK;
  }
  *bytes = (size_t)p.splits * p.taps * p.ci_tiles * kWM * p.co_pad * sizeof(float);
  return TMX_OK;
}

extern "C" int tmx_gram_fwd_tc(tmx_handle_t h, const uint16_t* f_hi, const uint16_t* f_lo, float* G, float* workspace,
                               int N, int C, int H, int W, tmx_stream_t s) {
  TMX_REQUIRE(h && f_hi && f_lo && G && workspace, TMX_ERR_ARG, "tmx_gram_fwd_tc: NULL argument");
  TMX_REQUIRE(N > 0 && C % 8 == 0 && H >= 2 && W >= 2 && (H * W) % kWKP == 0, TMX_ERR_SHAPE,
              "tmx_gram_fwd_tc: needs C %% 8 == 0 and H*W a multiple of %d (got C=%d %dx%d)", kWKP, C, H, W);
  const int bn_cols = wgrad_bn(C);
  WgradParams p;
  TMX_REQUIRE(gram_plan(h, N, C, H, W, bn_cols, p), TMX_ERR_UNSUPPORTED,
              "tmx_gram_fwd_tc: a 64-pixel chunk does not fit one %dx%d sample", H, W);
  p.partial = workspace;
  CUtensorMap maps[4];
  int rc;
  if ((rc = encode_map4(h, &maps[0], f_hi, N, H + 2, W + 2, C, p.bw, p.bh, p.bn))) return rc;
  if ((rc = encode_map4(h, &maps[1], f_lo, N, H + 2, W + 2, C, p.bw, p.bh, p.bn))) return rc;
  maps[2] = maps[0];
  maps[3] = maps[1];
  cudaStream_t st = (cudaStream_t)s;
  if (bn_cols == 256) rc = launch_wgrad<256>(h, maps, p, st);
  else if (bn_cols == 128) rc = launch_wgrad<128>(h, maps, p, st);
  else rc = launch_wgrad<64>(h, maps, p, st);
  if (rc) return rc;
  const long long total4 = (long long)N * C * C / 4;
  wgrad_reduce_kernel<<<tmx_ceil_div(total4, 256), 256, 0, st>>>(workspace, G, p.splits, p.taps, C, p.ci_tiles * kWM, C,
                                                                 p.co_pad, 1.0f / ((float)H * (float)W), 1);
  TMX_LAUNCHED(h, "wgrad_reduce_kernel");
  return TMX_OK;
}
