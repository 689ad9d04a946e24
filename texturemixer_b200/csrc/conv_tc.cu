// conv_tc.cu — tcgen05 implicit-GEMM convolution for sm_100a.
//
//   y = [residual +] lrelu(xcorr3x3_reflect(x, w) + bias)        (networks.py:48-75, :437)
//
// GEMM view per CTA tile: D[128 pixels][BN couts] += A[128][K] * B[BN][K]^T with
// K = taps * Cin walked as (tap, 64- or 32-channel chunk).
//   A  = activations in SPLIT_BF16_HALO layout ([N][H+2][W+2][C] bf16, hi and lo
//        planes).  A tap (u,v) of a bw x bh x bn pixel patch is ONE 4-D TMA box at
//        coordinates (c0, x0+v, y0+u, n0) - the reflect halo is materialised, so
//        no boundary logic exists in the main loop.  The box lands in shared
//        memory K-major with the 128B/64B hardware swizzle = the canonical UMMA
//        operand layout.
//   B  = weights prepared once as K-major [Cout][taps*Cin] bf16 hi/lo planes
//        (tmx_conv_weights_prepare), one 2-D TMA box per step.
//   D  = fp32 accumulators in TMEM (2 stages x BN columns) so the epilogue of
//        tile i overlaps the main loop of tile i+1.
// fp32 parity: x*w ~= xh*wh + xh*wl + xl*wh (bf16x3, fp32 accumulate): 16-17
// mantissa bits per product, three kind::f16 MMAs per K step; measured error is
// ~1e-5 relative on G_res vs 1.8e-3 for single-pass TF32.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles):
//   warp 0 lane 0 : TMA producer     warp 1 lane 0 : MMA issuer
//   warp 2        : TMEM allocator   warps 4-7     : epilogue (TMEM -> regs -> HBM)
// Epilogue fusions: bias, leaky-ReLU, residual add (fp32 stream), fp32 NHWC
// store, re-split into bf16 hi/lo planes INCLUDING the halo of the next layer
// (edge threads store their pixel to the mirrored / clamped halo slots too),
// optional nearest-neighbour x2 upsampling of the written planes, and the 1x1
// ToRGB head + tanh (networks.py:454-457, :483) written straight to NCHW.
//
// upscale2d + conv (networks.py:448-450) runs in sub-pixel form (UP2_IN): a 3x3
// conv of the x2 nearest-upsampled image equals, per output parity (a,b), a conv
// of the LOW-res image with taps summed pairwise, so the main loop is an
// ordinary 3x3 GEMM over the low-res planes (REPLICATE halo) with N = 4*Cout
// prepared weights and only the epilogue scatters column group (a,b) to pixel
// (2y+a, 2x+b): a quarter of the activation traffic of the materialised form.
#include "tc_common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, hardware swizzle:
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4
//   [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B)
// SBO = byte distance between 8-row groups = 8 * row_bytes (rows are dense).
template <int KC>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  constexpr uint64_t row_bytes = KC * 2;
  constexpr uint64_t layout = (KC == 64) ? 2 : (KC == 32 ? 4 : 6);  // SWIZZLE_128B / 64B / 32B
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (((8 * row_bytes) >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}
// Instruction descriptor (kind::f16): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n, int m = kTileM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct ConvTcParams {
  int N, H, W, Cin, Cout;
  int taps, k, pad_off;   // pad_off = 1 - k/2: halo offset of tap 0
  int bw, bh, bn;         // pixel patch of a tile: bw*bh*bn == 128
  int tiles_x, tiles_y, tiles_n, tiles_c, num_tiles;
  // work items: the first full_items tiles are whole (BN columns); the tiles of the last, partial wave
  // are split into `split` column slices of BN/split so that it still fills the SMs
  int full_items, split, num_items;
  int lrelu, has_res, up2_out;
  int phase;      // UP2_IN sub-pixel form: GEMM column = (a*2+b)*cout_log + c, output pixel (2y+a, 2x+b)
  int cout_log;   // logical Cout of the layer (== Cout unless phase)
  int halo_rep;   // halo of the written planes: 0 REFLECT, 1 REPLICATE (clamp), 2 ZERO (zeros written next to the border)
  // LIN mode (data gradient): pixels are the rows of ONE zero-ringed grid [N][H+4][W+4] shared by input and
  // output; a tile is 128 consecutive grid rows, tap (U,V) is the constant row shift (U-1)*pitch + (V-1), so the
  // A operand is a 2-D TMA box at row m0 + shift (out-of-range rows read zeros) and every grid position -
  // including the ring the reflect/replicate adjoint folds back - is produced.  Output: fp32 rows only.
  int lin, lin_pitch;
  long long lin_rows;
  // PATCH mode (X-MERGED 3-tap layers; TMX_NO_XMERGE_PATCH=1 turns it off for A/B runs): one stage = one tile; the A
  // operand of all three vertical taps is ONE haloed patch [(bh+2) rows][bw px][128 B] per plane, tap u reads it
  // from row offset u*bw (a whole number of 8-row swizzle atoms), the three weight boxes follow it.  These layers
  // run at the L2 throughput cap (~7.7 TB/s of L2->SM sectors, ncu): the patch cuts the sectors per tile by a third
  // (21.2 M -> 14.3 M for 16->16 at 128x128, batch 32: 87 -> 62 us; profiles/r02_ncu_thin_patch_summary.csv).
  int patch, patch_bytes, patch_stage_bytes, patch_stages;
  // TMX_CONV_W_PER_SAMPLE: the weight planes hold one [Cout][K] set per image; image n reads rows n*w_sample_rows + ..
  // (tiles lie inside one image: bn == 1; a CTA pair shares one weight tile, so tiles per image are even)
  int w_sample_rows;
  float alpha;
  const float* bias;
  const float* residual;
  float* y_f32;
  uint16_t* y_hi;
  uint16_t* y_lo;
  // GP epilogue (LIN mode, tmx_conv2d_dgrad_gp): tmx_grad_prepare fused into the data gradient (tc_common.cuh)
  GpParams gp;
  // fused ToRGB head
  int rgb_c, rgb_tanh;
  float rgb_wscale;
  const float* rgb_w;
  const float* rgb_b;
  float* y_rgb;
};

template <int BN, int KC, bool PAIR = false>
struct TcCfg {
  static constexpr int kABytes = kTileM * KC * 2;
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * KC * 2;  // per CTA: a pair splits the weight tile
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kSmemBudget = 200 * 1024;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 12 ? 12 : kStagesRaw;
  // STACK (single CTA, BN <= 64): bf16x3 as TWO MMAs per K step instead of three - the weight tile's hi and lo rows
  // sit next to each other in shared memory, so ONE MMA with N = 2*BN gives x_hi*w_hi and x_hi*w_lo, a second one
  // x_lo*w_hi; the accumulator has 3*BN columns which the epilogue adds.  A tcgen05.mma with a small N costs ~80-115
  // cycles whatever N is (the shared-memory read of its 4 KB A operand, measured in conv_lin.cu), so the thin layers'
  // main loop gets 1.5x shorter; at BN >= 128 the MMAs are math-bound and 2*BN would not fit a second accumulator.
  static constexpr bool kStack = !PAIR && BN <= 64;
  static constexpr int kAccCols = kStack ? 3 * BN : BN;
  static constexpr int kTmemRaw = 2 * kAccCols;
  static constexpr int kTmemCols = kTmemRaw <= 32 ? 32 : (kTmemRaw <= 64 ? 64 : (kTmemRaw <= 128 ? 128 : (kTmemRaw <= 256 ? 256 : 512)));
  static constexpr int kAuxBytes = 4096;  // barriers + tmem ptr, ToRGB weights (+256), GP bias-gradient partials (+1024: 512 floats)
  static constexpr int kSmemBytes = kStages * kStageBytes + kAuxBytes + 1024 /*align slack*/;
  static_assert(kStages >= 2, "need at least a double buffer");
  static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns: power of two <= 512");
};

__device__ __forceinline__ void decode_item(const ConvTcParams& p, int item, int& tile, int& part, int& parts) {
  if (item < p.full_items) {
    tile = item;
    part = 0;
    parts = 1;
  } else {
    const int q = item - p.full_items;
    tile = p.full_items + q / p.split;
    part = q - (q / p.split) * p.split;
    parts = p.split;
  }
}

template <int BN, int KC, int GW, bool PAIR, bool GP = false>
__global__ void __launch_bounds__(kThreads, 1)
    conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const __grid_constant__ CUtensorMap tm_bs_hi, const __grid_constant__ CUtensorMap tm_bs_lo,
                   const ConvTcParams p) {
  // PAIR: a cluster of two CTAs runs one 256-pixel x BN tile with cta_group::2 MMAs issued by the even
  // (leader) CTA.  Each CTA loads its own 128 pixels of A and HALF of the weight tile, so the L2->SM
  // weight traffic per pixel halves; accumulators live in both CTAs' TMEM (own 128 rows each).
  using Cfg = TcCfg<BN, KC, PAIR>;
  constexpr int S = Cfg::kStages;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int cta_lin = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // index of this CTA (pair) in the grid
  const int cta_cnt = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kBRowsFull = PAIR ? BN / 2 : BN;                          // weight rows this CTA loads for a whole tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* rgb_smem = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 256);   // [GW+1][4]: ToRGB weights + bias
  float* gp_bias_s = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 1024);  // [p.Cout <= 512] (GP)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cchunks = p.Cin / KC;
  const int ksteps = p.taps * cchunks;
  tmx_pdl_trigger();   // the next kernel of the stream may be scheduled as SMs free up (it blocks in its own wait)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    if (p.split > 1) {
      tma_prefetch_desc(&tm_bs_hi);
      tma_prefetch_desc(&tm_bs_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 8 : 4);  // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc2<Cfg::kTmemCols>(tmem_ptr_s);
    else tmem_alloc<Cfg::kTmemCols>(tmem_ptr_s);
  }
  // everything above touched no global memory: it overlaps the tail of the previous kernel (PDL); from here on the
  // previous kernel's results are read
  tmx_pdl_wait();
  if (GP && threadIdx.x >= 128)
    for (int c = threadIdx.x - 128; c < p.Cout; c += 128) gp_bias_s[c] = 0.f;
  if (p.y_rgb != nullptr && threadIdx.x >= 128) {
    for (int e = threadIdx.x - 128; e < (GW + 1) * 4; e += 128) {   // visible after the setup barrier below
      const int c = e >> 2, o = e & 3;
      float val = 0.f;
      if (o < p.rgb_c) val = c < GW ? __ldg(p.rgb_w + c * p.rgb_c + o) : (p.rgb_b ? __ldg(p.rgb_b + o) : 0.f);
      rgb_smem[e] = val;
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // barrier inits + TMEM allocation visible in both CTAs
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cta_lin; item < p.num_items; item += cta_cnt) {
        int tile, part, parts;
        decode_item(p, item, tile, part, parts);
        const int cblk = tile % p.tiles_c;
        int t = tile / p.tiles_c;
        if (PAIR) t = 2 * t + rank;   // the pair covers pixel tiles 2t (leader) and 2t+1; past-the-end tiles load zeros
        const int tlin = t;
        const int x0 = (t % p.tiles_x) * p.bw;
        t /= p.tiles_x;
        const int y0 = (t % p.tiles_y) * p.bh;
        const int n0 = (t / p.tiles_y) * p.bn;
        const CUtensorMap* mb_hi = parts > 1 ? &tm_bs_hi : &tm_b_hi;
        const CUtensorMap* mb_lo = parts > 1 ? &tm_bs_lo : &tm_b_lo;
        const int brows = kBRowsFull / parts;
        const int brow = cblk * BN + part * (BN / parts) + rank * brows + n0 * p.w_sample_rows;
        const uint32_t bytes = 2 * Cfg::kABytes + 2 * brows * KC * 2;
        if (!PAIR && p.patch) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + (size_t)stage * p.patch_stage_bytes;
          uint8_t* sb = sa + 2 * p.patch_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], 2 * p.patch_bytes + 6 * brows * KC * 2);
          tma_load_4d(sa, &tm_a_hi, &full_bar[stage], 0, x0, y0, n0);            // halo rows y0 .. y0 + bh + 1
          tma_load_4d(sa + p.patch_bytes, &tm_a_lo, &full_bar[stage], 0, x0, y0, n0);
#pragma unroll
          for (int u = 0; u < 3; ++u) {     // per vertical tap: [hi rows | lo rows] (one stacked B operand)
            tma_load_2d(sb + (2 * u) * brows * KC * 2, mb_hi, &full_bar[stage], u * p.Cin, brow);
            tma_load_2d(sb + (2 * u + 1) * brows * KC * 2, mb_lo, &full_bar[stage], u * p.Cin, brow);
          }
          if (++stage == p.patch_stages) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        for (int ks = 0; ks < ksteps; ++ks) {
          const int tap = ks / cchunks;
          const int c0 = (ks - tap * cchunks) * KC;
          const int u = tap / p.k, v = tap - u * p.k;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          const int ax = x0 + v + p.pad_off, ay = y0 + u + p.pad_off, bk = tap * p.Cin + c0;
          const int arow = tlin * kTileM + (p.k == 3 ? (u - 1) * p.lin_pitch + (v - 1) : 0);   // LIN mode
          if (PAIR) {
            // the leader's barrier collects the bytes of both CTAs
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * bytes);
            if (p.lin) {
              tma2_load_2d(sa, &tm_a_hi, &full_bar[stage], c0, arow);
              tma2_load_2d(sa + Cfg::kABytes, &tm_a_lo, &full_bar[stage], c0, arow);
            } else {
              tma2_load_4d(sa, &tm_a_hi, &full_bar[stage], c0, ax, ay, n0);
              tma2_load_4d(sa + Cfg::kABytes, &tm_a_lo, &full_bar[stage], c0, ax, ay, n0);
            }
            tma2_load_2d(sa + 2 * Cfg::kABytes, mb_hi, &full_bar[stage], bk, brow);
            tma2_load_2d(sa + 2 * Cfg::kABytes + Cfg::kBBytes, mb_lo, &full_bar[stage], bk, brow);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            if (p.lin) {
              tma_load_2d(sa, &tm_a_hi, &full_bar[stage], c0, arow);
              tma_load_2d(sa + Cfg::kABytes, &tm_a_lo, &full_bar[stage], c0, arow);
            } else {
              tma_load_4d(sa, &tm_a_hi, &full_bar[stage], c0, ax, ay, n0);
              tma_load_4d(sa + Cfg::kABytes, &tm_a_lo, &full_bar[stage], c0, ax, ay, n0);
            }
            // the lo rows directly behind the hi rows (a column slice of a split item has fewer rows than BN)
            tma_load_2d(sa + 2 * Cfg::kABytes, mb_hi, &full_bar[stage], bk, brow);
            tma_load_2d(sa + 2 * Cfg::kABytes + (Cfg::kStack ? brows * KC * 2 : Cfg::kBBytes), mb_lo, &full_bar[stage], bk,
                        brow);
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only in pair mode) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = cta_lin; item < p.num_items; item += cta_cnt, ++it) {
        const uint32_t idesc = make_idesc(item < p.full_items ? BN : BN / p.split, PAIR ? 2 * kTileM : kTileM);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * Cfg::kAccCols);
        const int bn_i = item < p.full_items ? BN : BN / p.split;          // columns of this item
        const uint32_t idesc2 = make_idesc(2 * bn_i, kTileM);                 // STACK: N = hi rows + lo rows
        if (!PAIR && p.patch) {
          const int brows_i = bn_i;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * p.patch_stage_bytes);
          const uint32_t sb = sa + 2 * p.patch_bytes;
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const uint32_t aoff = (uint32_t)(u * p.bw * KC * 2);           // u patch rows down: whole swizzle atoms
            const uint64_t a_hi = make_smem_desc<KC>(sa + aoff);
            const uint64_t a_lo = make_smem_desc<KC>(sa + p.patch_bytes + aoff);
            const uint64_t b_hi = make_smem_desc<KC>(sb + (2 * u) * brows_i * KC * 2);
#pragma unroll
            for (int kk = 0; kk < KC / 16; ++kk) {      // (patch mode implies BN <= 64: always STACK)
              const uint64_t adv = (uint64_t)(kk * 2);
              umma_bf16(tmem_d + 2 * bn_i, a_lo + adv, b_hi + adv, idesc, (u | kk) != 0);
              umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc2, (u | kk) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.patch_stages) {
            stage = 0;
            phase ^= 1;
          }
          umma_commit(&tfull_bar[as]);
          continue;
        }
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint64_t a_hi = make_smem_desc<KC>(sa);
          const uint64_t a_lo = make_smem_desc<KC>(sa + Cfg::kABytes);
          const uint64_t b_hi = make_smem_desc<KC>(sa + 2 * Cfg::kABytes);
          const uint64_t b_lo = make_smem_desc<KC>(sa + 2 * Cfg::kABytes + Cfg::kBBytes);
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);  // +32 B along K inside the swizzle row, >>4
            // small terms first, then the dominant hi*hi product
            if (PAIR) {
              umma2_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, (ks | kk) != 0);
              umma2_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
              umma2_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
            } else if (Cfg::kStack) {
              umma_bf16(tmem_d + 2 * bn_i, a_lo + adv, b_hi + adv, idesc, (ks | kk) != 0);
              umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc2, (ks | kk) != 0);      // b_hi + the lo rows behind it
            } else {
              umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, (ks | kk) != 0);
              umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
              umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
            }
          }
          // frees the smem slot (in both CTAs) when these MMAs retire
          if (PAIR) umma2_commit(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (PAIR) umma2_commit(&tfull_bar[as]);
        else umma_commit(&tfull_bar[as]);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    // One thread = one tile row (pixel); columns are processed in groups of GW
    // channels.  A group lies inside one output phase (a,b) when p.phase.
    const int quad = warp & 3;            // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;       // row of the tile = pixel
    const int in_ = r / (p.bh * p.bw);
    const int rem = r - in_ * (p.bh * p.bw);
    const int iy = rem / p.bw, ix = rem - iy * p.bw;
    const int Hl = p.phase ? 2 * p.H : p.H, Wl = p.phase ? 2 * p.W : p.W;      // logical output size
    const int Ho = p.up2_out ? 2 * Hl : Hl, Wo = p.up2_out ? 2 * Wl : Wl;      // size of the written planes
    const long long Hp = Ho + 2, Wp = Wo + 2;
    const int CL = p.cout_log;
    // source row/col copied into halo slot 0, and (size - hi_off) into slot size+1; ZERO halo: the border pixels
    // write zeros into the slots next to them (so the planes need no memset)
    const int lo_edge = p.halo_rep >= 1 ? 0 : 1;
    const int hi_off = p.halo_rep >= 1 ? 1 : 2;
    const bool halo_zero = p.halo_rep == 2;
    const bool has_bias = p.bias != nullptr;
    const float slope = p.lrelu ? p.alpha : 1.f;
    const float4* rgb_s = reinterpret_cast<const float4*>(rgb_smem);
    int it = 0;
    for (int item = cta_lin; item < p.num_items; item += cta_cnt, ++it) {
      int tile, part, parts;
      decode_item(p, item, tile, part, parts);
      const int cblk = tile % p.tiles_c;
      int t = tile / p.tiles_c;
      if (PAIR) t = 2 * t + rank;
      const long long mlin = (long long)t * kTileM + r;          // LIN mode: grid row of this thread
      const int x = (t % p.tiles_x) * p.bw + ix;
      t /= p.tiles_x;
      const int y = (t % p.tiles_y) * p.bh + iy;
      const int n = (t / p.tiles_y) * p.bn + in_;
      const bool valid = p.lin ? (mlin < p.lin_rows) : (n < p.N);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      GpRow gp_row;
      if (GP) gp_row = gp_classify(p.gp, mlin, valid);

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * Cfg::kAccCols);
      const int bn_i = BN / parts;
      const int ngroups = bn_i / GW;
      // residual stream (networks.py:437; never combined with the sub-pixel form): the loads of group g+1 are
      // issued before the TMEM read of group g so that their latency hides behind it
      const bool res_on = p.has_res && valid;
      const long long res_pix = p.lin ? mlin : ((long long)n * Hl + y) * Wl + x;
      const float4* res_base =
          reinterpret_cast<const float4*>(p.residual + res_pix * CL + cblk * BN + part * (BN / parts));
      float4 rcur[GW / 4], rnext[GW / 4];
      if (res_on) {
#pragma unroll
        for (int j = 0; j < GW / 4; ++j) rcur[j] = __ldg(res_base + j);
      }
#pragma unroll 1
      for (int g = 0; g < ngroups; ++g) {
        if (res_on && g + 1 < ngroups) {
#pragma unroll
          for (int j = 0; j < GW / 4; ++j) rnext[j] = __ldg(res_base + (g + 1) * (GW / 4) + j);
        }
        uint32_t acc[32];
        if (Cfg::kStack) {
          // three column groups of the accumulator: x_lo*w_hi + x_hi*w_lo first, then the dominant x_hi*w_hi
          uint32_t tmp[32];
          if (GW == 32) {
            tmem_ld32(taddr0 + 2 * bn_i + g * GW, acc);
            tmem_ld32(taddr0 + bn_i + g * GW, tmp);
          } else {
            tmem_ld16(taddr0 + 2 * bn_i + g * GW, acc);
            tmem_ld16(taddr0 + bn_i + g * GW, tmp);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < GW; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(tmp[j]));
          if (GW == 32) tmem_ld32(taddr0 + g * GW, tmp);
          else tmem_ld16(taddr0 + g * GW, tmp);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < GW; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(tmp[j]));
        } else {
          if (GW == 32) tmem_ld32(taddr0 + g * GW, acc);
          else tmem_ld16(taddr0 + g * GW, acc);
          tmem_ld_wait();
        }
        if constexpr (GP) {
          const int col0 = cblk * BN + part * (BN / parts) + g * GW;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = j < GW ? __uint_as_float(acc[j]) : 0.f;
          gp_group<GW>(p.gp, gp_row, mlin, CL, col0, v, gp_bias_s, lane);
        } else if (valid) {
          const int col0 = cblk * BN + part * (BN / parts) + g * GW;
          int cbase = col0, oy = y, ox = x;
          if (p.phase) {
            const int ph = col0 / CL;
            cbase = col0 - ph * CL;
            oy = 2 * y + (ph >> 1);
            ox = 2 * x + (ph & 1);
          }
          // bias as 128-bit loads, leaky-ReLU branch-free (slope 1 == identity: max(f, f))
          float v[GW];
          if (has_bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + cbase);
#pragma unroll
            for (int j = 0; j < GW / 4; ++j) {
              const float4 b4 = __ldg(bp + j);
              v[4 * j] = __uint_as_float(acc[4 * j]) + b4.x;
              v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b4.y;
              v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b4.z;
              v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < GW; ++j) v[j] = __uint_as_float(acc[j]);
          }
#pragma unroll
          for (int j = 0; j < GW; ++j) v[j] = fmaxf(v[j] * slope, v[j]);
          const long long pix = p.lin ? mlin : ((long long)n * Hl + oy) * Wl + ox;
          if (p.has_res) {
#pragma unroll
            for (int j = 0; j < GW / 4; ++j) {
              v[4 * j] += rcur[j].x;
              v[4 * j + 1] += rcur[j].y;
              v[4 * j + 2] += rcur[j].z;
              v[4 * j + 3] += rcur[j].w;
              rcur[j] = rnext[j];
            }
          }
          if (p.y_f32 != nullptr) {
            float4* op = reinterpret_cast<float4*>(p.y_f32 + pix * CL + cbase);
#pragma unroll
            for (int j = 0; j < GW / 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (p.y_hi != nullptr) {
            uint32_t ph[GW / 2], pl[GW / 2];
#pragma unroll
            for (int j = 0; j < GW / 2; ++j) {
              tmx_split_bf16x2(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
            }
            auto store_px = [&](int prow, int pcol, bool halo = true) {
              const long long o = (((long long)n * Hp + prow) * Wp + pcol) * CL + cbase;
              uint4* oh = reinterpret_cast<uint4*>(p.y_hi + o);
              uint4* ol = reinterpret_cast<uint4*>(p.y_lo + o);
              const bool z = halo && halo_zero;
#pragma unroll
              for (int j = 0; j < GW / 8; ++j) {
                oh[j] = z ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
                ol[j] = z ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
              }
            };
            if (!p.up2_out) {
              // interior slot plus at most one halo row and one halo column this pixel feeds (H, W >= 2)
              const int r0 = oy + 1, c0 = ox + 1;
              const int r1 = oy == lo_edge ? 0 : (oy == Ho - hi_off ? Ho + 1 : -1);
              const int c1 = ox == lo_edge ? 0 : (ox == Wo - hi_off ? Wo + 1 : -1);
              store_px(r0, c0, false);
              if (r1 >= 0) store_px(r1, c0);
              if (c1 >= 0) {
                store_px(r0, c1);
                if (r1 >= 0) store_px(r1, c1);
              }
            } else {
              // x2 nearest upsampling of the written planes: 2x2 copies, each with its own halo slots
              for (int d = 0; d < 2; ++d) {
                const int Y = 2 * oy + d;
                const int r1 = Y == lo_edge ? 0 : (Y == Ho - hi_off ? Ho + 1 : -1);
                for (int e = 0; e < 2; ++e) {
                  const int X = 2 * ox + e;
                  const int c1 = X == lo_edge ? 0 : (X == Wo - hi_off ? Wo + 1 : -1);
                  store_px(Y + 1, X + 1, false);
                  if (r1 >= 0) store_px(r1, X + 1);
                  if (c1 >= 0) {
                    store_px(Y + 1, c1);
                    if (r1 >= 0) store_px(r1, c1);
                  }
                }
              }
            }
          }
          if (p.y_rgb != nullptr) {
            // ToRGB 1x1 head on the fp32 activations of this pixel (host guarantees CL == GW, rgb_c <= 4);
            // the head's weights were staged in shared memory as [c][4] at kernel start
            float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
            for (int j = 0; j < GW; ++j) {
              const float4 w4 = rgb_s[j];
              r0 = fmaf(v[j], w4.x, r0);
              r1 = fmaf(v[j], w4.y, r1);
              r2 = fmaf(v[j], w4.z, r2);
              r3 = fmaf(v[j], w4.w, r3);
            }
            const float4 b4 = rgb_s[GW];
            const float rr[4] = {fmaf(r0, p.rgb_wscale, b4.x), fmaf(r1, p.rgb_wscale, b4.y),
                                 fmaf(r2, p.rgb_wscale, b4.z), fmaf(r3, p.rgb_wscale, b4.w)};
            const long long hw = (long long)Hl * Wl;
            float* dst = p.y_rgb + (long long)n * p.rgb_c * hw + (long long)oy * Wl + ox;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              if (o < p.rgb_c) dst[o * hw] = p.rgb_tanh ? tanhf(rr[o]) : rr[o];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tempty_bar[as]);
        else mbar_arrive(&tempty_bar[as]);
      }
    }
    if (GP) gp_flush_bias(p.gp, gp_bias_s, p.Cout);
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer can still signal / be signalled
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------- host side
CUtensorMapSwizzle swizzle_of(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// xwin > 1: X-MERGED view for thin layers - the innermost dimension spans `xwin` consecutive pixels (xwin*C
// channels), while the pixel stride stays C: overlapping rows, so that ONE box row of 128 B carries the three
// horizontal taps of a 16-channel layer (plus one zero-weighted neighbour) instead of three 32-B rows.
int encode_act_map(tmx_handle_t h, CUtensorMap* m, const uint16_t* base, int N, int Hp, int Wp, int C, int kc, int bw,
                   int bh, int bn, int xwin = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)C * xwin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_of(kc);
  CUresult r = h->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tmx_fail(TMX_ERR_DRIVER, "cuTensorMapEncodeTiled(activations) failed: CUresult %d (C=%d Wp=%d Hp=%d N=%d box %d,%d,%d,%d)",
                    (int)r, C, Wp, Hp, N, kc, bw, bh, bn);
  return TMX_OK;
}

int encode_wgt_map(tmx_handle_t h, CUtensorMap* m, const uint16_t* base, int Cout, int K, int kc, int bnc) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bnc};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = swizzle_of(kc);
  CUresult r = h->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tmx_fail(TMX_ERR_DRIVER, "cuTensorMapEncodeTiled(weights) failed: CUresult %d (K=%d Cout=%d box %d,%d)",
                    (int)r, K, Cout, kc, bnc);
  return TMX_OK;
}

int gcd_pow2(int v, int cap) {  // largest power of two dividing v, at most cap
  int g = 1;
  while (g < cap && (v % (g * 2)) == 0) g *= 2;
  return g;
}

template <int BN, int KC, int GW, bool PAIR, bool GP = false>
int launch_tc(tmx_handle_t h, const CUtensorMap* maps, const ConvTcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN, KC, PAIR>;
  auto kern = conv_tc_kernel<BN, KC, GW, PAIR, GP>;
  static thread_local int configured_device = -1;
  if (configured_device != h->device) {
    TMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured_device = h->device;
  }
  TMX_REQUIRE(Cfg::kSmemBytes <= h->max_smem_optin, TMX_ERR_UNSUPPORTED,
              "tmx_conv2d_fwd[TC]: kernel needs %d B shared memory, device allows %d", Cfg::kSmemBytes,
              h->max_smem_optin);
  const int units = PAIR ? h->sm_count / 2 : h->sm_count;   // CTAs, or CTA pairs
  const int n = p.num_items < units ? p.num_items : units;
  TMX_CUDA(tmx_launch_pdl(kern, dim3(PAIR ? 2 * n : n), dim3(kThreads), (size_t)Cfg::kSmemBytes, st, PAIR ? 2 : 1,
                          maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p));
  TMX_LAUNCHED(h, "conv_tc_kernel");
  return TMX_OK;
}

// Activations as a plain 2-D matrix [rows][C] (LIN mode): box = 128 rows x kc channels.
int encode_rows_map(tmx_handle_t h, CUtensorMap* m, const uint16_t* base, long long rows, int C, int kc) {
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)kTileM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tmx_fail(TMX_ERR_DRIVER, "cuTensorMapEncodeTiled(rows) failed: CUresult %d (rows=%lld C=%d)", (int)r, rows, C);
  return TMX_OK;
}

// Common tail of the forward and data-gradient launches: tile schedule (CTA pairs, last-wave split), weight
// tensor maps, kernel selection.  `tiles_m` = number of 128-row tiles, Ng = GEMM N, Kc = GEMM K per tap.
int schedule_and_launch(tmx_handle_t h, ConvTcParams& p, CUtensorMap* maps, const uint16_t* w_hi, const uint16_t* w_lo,
                        long long tiles_m, int Ng, int Kc, int bnc, int kc, int gw, cudaStream_t st, int w_sets = 1,
                        bool gp = false) {
  p.tiles_c = Ng / bnc;
  // CTA pairs (cta_group::2) for the wide layers: two pixel tiles share one weight tile split over the pair
  const bool pair = bnc == 256 && kc == 64 && tiles_m >= 2 && !tmx_env_flag("TMX_NO_PAIR");
  if (pair) tiles_m = (tiles_m + 1) / 2;
  long long nt = tiles_m * p.tiles_c;
  TMX_REQUIRE(nt < (1ll << 31), TMX_ERR_SHAPE, "conv_tc: too many tiles");
  p.num_tiles = (int)nt;
  // last-wave split: T tiles over G SMs leave R = T mod G tiles for a final, partly empty wave; cut those
  // into 2 or 4 column slices (>= 32 columns, whole epilogue groups) when that lets the wave fill the SMs
  {
    const int G = pair ? h->sm_count / 2 : h->sm_count;
    const int R = p.num_tiles % G;
    p.split = 1;
    if (R > 0 && !tmx_env_flag("TMX_NO_SPLIT")) {
      for (int sp = 4; sp >= 2; sp /= 2) {
        const int bs = bnc / sp;
        if (bs >= 32 && bs % gw == 0 && (long long)R * sp <= G) {
          p.split = sp;
          break;
        }
      }
    }
    p.full_items = p.split > 1 ? p.num_tiles - R : p.num_tiles;
    p.num_items = p.full_items + (p.split > 1 ? R * p.split : 0);
  }
  int rc;
  const int brows = pair ? bnc / 2 : bnc;   // weight rows one CTA loads per stage
  if ((rc = encode_wgt_map(h, &maps[2], w_hi, Ng * w_sets, p.taps * Kc, kc, brows))) return rc;
  if ((rc = encode_wgt_map(h, &maps[3], w_lo, Ng * w_sets, p.taps * Kc, kc, brows))) return rc;
  if ((rc = encode_wgt_map(h, &maps[4], w_hi, Ng * w_sets, p.taps * Kc, kc, brows / p.split))) return rc;
  if ((rc = encode_wgt_map(h, &maps[5], w_lo, Ng * w_sets, p.taps * Kc, kc, brows / p.split))) return rc;
  if (gp) {   // LIN mode with the fused tmx_grad_prepare epilogue (the caller checked gp_variant_exists)
    if (pair) return launch_tc<256, 64, 32, true, true>(h, maps, p, st);
#define TMX_GP_CASE(BN_, KC_) \
  if (bnc == BN_ && kc == KC_ && gw == 32) return launch_tc<BN_, KC_, 32, false, true>(h, maps, p, st);
    TMX_GP_CASE(256, 64)
    TMX_GP_CASE(256, 32)
    TMX_GP_CASE(128, 64)
    TMX_GP_CASE(128, 32)
    TMX_GP_CASE(64, 64)
    TMX_GP_CASE(64, 32)
    TMX_GP_CASE(32, 64)
    TMX_GP_CASE(32, 32)
#undef TMX_GP_CASE
    return tmx_fail(TMX_ERR_UNSUPPORTED, "conv_tc: no GP kernel for BN=%d KC=%d GW=%d", bnc, kc, gw);
  }
  if (pair) return launch_tc<256, 64, 32, true>(h, maps, p, st);

#define TMX_TC_CASE(BN_, KC_, GW_) \
  if (bnc == BN_ && kc == KC_ && gw == GW_) return launch_tc<BN_, KC_, GW_, false>(h, maps, p, st);
  TMX_TC_CASE(256, 64, 32)
  TMX_TC_CASE(256, 32, 32)
  TMX_TC_CASE(128, 64, 32)
  TMX_TC_CASE(128, 32, 32)
  TMX_TC_CASE(128, 16, 32)
  TMX_TC_CASE(64, 64, 32)
  TMX_TC_CASE(64, 32, 32)
  TMX_TC_CASE(64, 16, 32)
  TMX_TC_CASE(32, 64, 32)
  TMX_TC_CASE(32, 32, 32)
  TMX_TC_CASE(32, 16, 32)
  TMX_TC_CASE(64, 64, 16)
  TMX_TC_CASE(64, 32, 16)
  TMX_TC_CASE(64, 16, 16)
  TMX_TC_CASE(16, 64, 16)
  TMX_TC_CASE(16, 32, 16)
  TMX_TC_CASE(16, 16, 16)
#undef TMX_TC_CASE
  return tmx_fail(TMX_ERR_UNSUPPORTED, "conv_tc: no kernel for BN=%d KC=%d GW=%d (K=%d N=%d)", bnc, kc, gw, Kc, Ng);
}

}  // namespace

int tmx_conv2d_fwd_tc(tmx_handle_t h, const tmx_conv_desc_t* d, const tmx_conv_io_t* io, int kc_max, cudaStream_t st) {
  const bool phase = (d->flags & TMX_CONV_UP2_IN) != 0;
  const bool torgb = (d->flags & TMX_CONV_TORGB) != 0;
  TMX_REQUIRE(io->x_hi && io->x_lo && io->w_hi && io->w_lo, TMX_ERR_ARG,
              "tmx_conv2d_fwd[TC]: x_hi/x_lo (SPLIT_BF16_HALO) and w_hi/w_lo (prepared) are required");
  TMX_REQUIRE(io->y_f32 || (io->y_hi && io->y_lo) || (torgb && io->y_rgb), TMX_ERR_ARG,
              "tmx_conv2d_fwd[TC]: no output");
  TMX_REQUIRE((io->y_hi == nullptr) == (io->y_lo == nullptr), TMX_ERR_ARG,
              "tmx_conv2d_fwd[TC]: y_hi and y_lo go together");
  TMX_REQUIRE(!(d->flags & TMX_CONV_UP2_OUT) || io->y_hi, TMX_ERR_ARG,
              "tmx_conv2d_fwd[TC]: UP2_OUT applies to the split-plane output");
  TMX_REQUIRE(!(phase && (d->flags & TMX_CONV_UP2_OUT)), TMX_ERR_UNSUPPORTED,
              "tmx_conv2d_fwd[TC]: UP2_IN and UP2_OUT cannot be combined");
  TMX_REQUIRE(!((d->flags & (TMX_CONV_HALO_REPLICATE | TMX_CONV_HALO_ZERO)) && (d->flags & TMX_CONV_UP2_OUT)),
              TMX_ERR_UNSUPPORTED, "tmx_conv2d_fwd[TC]: HALO_REPLICATE / HALO_ZERO and UP2_OUT cannot be combined");
  TMX_REQUIRE(!((d->flags & TMX_CONV_HALO_REPLICATE) && (d->flags & TMX_CONV_HALO_ZERO)), TMX_ERR_ARG,
              "tmx_conv2d_fwd[TC]: HALO_REPLICATE and HALO_ZERO exclude each other");
  TMX_REQUIRE(!(d->flags & TMX_CONV_RESIDUAL) || io->residual, TMX_ERR_ARG,
              "tmx_conv2d_fwd[TC]: RESIDUAL flag without residual pointer");
  TMX_REQUIRE(!(phase && (d->flags & TMX_CONV_RESIDUAL)), TMX_ERR_UNSUPPORTED,
              "tmx_conv2d_fwd[TC]: RESIDUAL and UP2_IN cannot be combined");
  TMX_REQUIRE(!phase || (d->k == 3 && d->H % 2 == 0 && d->W % 2 == 0), TMX_ERR_SHAPE,
              "tmx_conv2d_fwd[TC]: UP2_IN needs k == 3 and even H, W (got k=%d, %d x %d)", d->k, d->H, d->W);
  TMX_REQUIRE(d->Cin % 16 == 0 && d->Cout % 16 == 0, TMX_ERR_SHAPE,
              "tmx_conv2d_fwd[TC]: Cin=%d and Cout=%d must be multiples of 16", d->Cin, d->Cout);
  int kc = d->Cin % 64 == 0 ? 64 : (d->Cin % 32 == 0 ? 32 : 16);
  if (kc > kc_max) kc = kc_max;
  // thin layers: three horizontal taps of a 16-channel input in one 128-B operand row (weights prepared with
  // tmx_conv_weights_prepare xmerge layout [Cout][3][64]); the caller opts in because the plane buffers must be
  // readable (and finite) 3 pixels past their end
  const bool xmerge = (d->flags & TMX_CONV_XMERGE) != 0;
  TMX_REQUIRE(!xmerge || (d->Cin == 16 && d->k == 3 && !phase), TMX_ERR_SHAPE,
              "tmx_conv2d_fwd[TC]: XMERGE needs Cin == 16, k == 3, no UP2_IN");
  if (xmerge) kc = 64;
  // stored (GEMM-side) geometry
  const int Hs = phase ? d->H / 2 : d->H, Ws = phase ? d->W / 2 : d->W;
  const int Ng = phase ? 4 * d->Cout : d->Cout;  // GEMM N
  TMX_REQUIRE(Hs >= 2 && Ws >= 2, TMX_ERR_SHAPE, "tmx_conv2d_fwd[TC]: stored H, W >= 2 (halo)");
  const int gw = d->Cout % 32 == 0 ? 32 : 16;
  int bnc;
  if (Ng % 256 == 0) bnc = 256;
  else if (Ng % 128 == 0) bnc = 128;
  else if (Ng % 64 == 0) bnc = 64;
  else if (Ng % 32 == 0) bnc = 32;
  else bnc = 16;
  if (torgb) {
    TMX_REQUIRE(io->rgb_w && io->y_rgb && d->rgb_cout >= 1 && d->rgb_cout <= 4, TMX_ERR_ARG,
                "tmx_conv2d_fwd[TC]: TORGB needs rgb_w, y_rgb and 1 <= rgb_cout <= 4");
    TMX_REQUIRE(d->Cout == gw, TMX_ERR_SHAPE,
                "tmx_conv2d_fwd[TC]: TORGB needs Cout in {16, 32} (one column group per pixel), got %d", d->Cout);
  }

  ConvTcParams p;
  p.N = d->N;
  p.H = Hs;
  p.W = Ws;
  p.Cin = xmerge ? 64 : d->Cin;          // contraction length per (merged) tap
  p.Cout = Ng;
  p.k = xmerge ? 1 : d->k;               // merged: tap index == vertical tap u, no horizontal shift
  p.taps = xmerge ? 3 : d->k * d->k;
  p.pad_off = xmerge ? 0 : 1 - d->k / 2;
  p.bw = gcd_pow2(Ws, 32);
  p.bh = gcd_pow2(Hs, kTileM / p.bw);
  p.bn = kTileM / (p.bw * p.bh);
  p.tiles_x = Ws / p.bw;
  p.tiles_y = Hs / p.bh;
  p.tiles_n = (d->N + p.bn - 1) / p.bn;
  p.lin = 0;
  p.lin_pitch = 0;
  p.lin_rows = 0;
  const long long tiles_m = (long long)p.tiles_x * p.tiles_y * p.tiles_n;
  const bool per_sample = (d->flags & TMX_CONV_W_PER_SAMPLE) != 0;
  TMX_REQUIRE(!per_sample || (p.bn == 1 && !phase && !xmerge && !torgb &&
                              (bnc != 256 || kc != 64 || (p.tiles_x * p.tiles_y) % 2 == 0)),
              TMX_ERR_UNSUPPORTED,
              "tmx_conv2d_fwd[TC]: W_PER_SAMPLE needs >= 128 pixels per image (an even number of 128-pixel tiles for "
              "the 256-column kernel), no UP2_IN / XMERGE / TORGB (got %d x %d)", Hs, Ws);
  p.w_sample_rows = per_sample ? Ng : 0;
  p.lrelu = (d->flags & TMX_CONV_LRELU) != 0;
  p.has_res = (d->flags & TMX_CONV_RESIDUAL) != 0;
  p.up2_out = (d->flags & TMX_CONV_UP2_OUT) != 0;
  p.phase = phase;
  p.cout_log = d->Cout;
  p.halo_rep = (d->flags & TMX_CONV_HALO_REPLICATE) ? 1 : ((d->flags & TMX_CONV_HALO_ZERO) ? 2 : 0);
  p.alpha = d->lrelu_alpha;
  p.bias = io->bias;
  p.residual = io->residual;
  p.y_f32 = io->y_f32;
  p.y_hi = io->y_hi;
  p.y_lo = io->y_lo;
  p.rgb_c = torgb ? d->rgb_cout : 0;
  p.rgb_tanh = d->rgb_tanh;
  p.rgb_wscale = d->rgb_wscale;
  p.rgb_w = torgb ? io->rgb_w : nullptr;
  p.rgb_b = torgb ? io->rgb_b : nullptr;
  p.y_rgb = torgb ? io->y_rgb : nullptr;

  CUtensorMap maps[6];
  int rc;
  const int xw = xmerge ? 4 : 1;
  // PATCH mode: see ConvTcParams::patch
  p.patch = p.patch_bytes = p.patch_stage_bytes = p.patch_stages = 0;
  int box_h = p.bh;
  if (xmerge && !tmx_env_flag("TMX_NO_XMERGE_PATCH") && p.bn == 1 && p.bw % 8 == 0 && bnc <= 64) {
    const int k_a = kTileM * 64 * 2, k_b = bnc * 64 * 2;                  // TcCfg<bnc, 64>: bytes per plane
    const int stage = 2 * k_a + 2 * k_b;
    int stages = (200 * 1024) / stage;
    if (stages > 12) stages = 12;
    const int patch_bytes = (p.bh + 2) * p.bw * 64 * 2;
    const int patch_stage = 2 * patch_bytes + 6 * k_b;
    int patch_stages = (stages * stage) / patch_stage;     // inside the launch's stage area (TcCfg::kStages * kStageBytes)
    if (patch_stages > stages) patch_stages = stages;      // one full/empty barrier pair per stage exists
    if (patch_stages >= 2) {
      p.patch = 1;
      p.patch_bytes = patch_bytes;
      p.patch_stage_bytes = patch_stage;
      p.patch_stages = patch_stages > 12 ? 12 : patch_stages;
      box_h = p.bh + 2;
    }
  }
  if ((rc = encode_act_map(h, &maps[0], io->x_hi, d->N, Hs + 2, Ws + 2, d->Cin, kc, p.bw, box_h, p.bn, xw))) return rc;
  if ((rc = encode_act_map(h, &maps[1], io->x_lo, d->N, Hs + 2, Ws + 2, d->Cin, kc, p.bw, box_h, p.bn, xw))) return rc;
  return schedule_and_launch(h, p, maps, io->w_hi, io->w_lo, tiles_m, Ng, p.Cin, bnc, kc, gw, st,
                             per_sample ? d->N : 1);
}

// ---------------------------------------------------------------- data gradient (LIN mode)
// dL/dx of y = conv3x3(pad(x)) (or conv1x1) before the padding adjoint: for EVERY position of the zero-ringed
// grid [N][H+4][W+4] (interior at offset 2)   g[r][c][ci] = sum_{u,v,co} dz[r+1-u][c+1-v][co] * w[u][v][ci][co],
// i.e. the forward kernel with K = Cout, N = Cin, flipped taps (tmx_conv_weights_prepare_dgrad) and a 2-D
// row-shifted A operand.  The ring 1 <= r <= H+2 holds the values the REFLECT / REPLICATE adjoint folds onto
// the interior (tmx_grad_prepare); the outermost ring is garbage by construction and never read.
int tmx_conv2d_dgrad_tc(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                        const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                        cudaStream_t st) {
  TMX_REQUIRE(dz_hi && dz_lo && wt_hi && wt_lo && g_f32, TMX_ERR_ARG, "tmx_conv2d_dgrad: NULL argument");
  TMX_REQUIRE((k == 1 || k == 3) && N > 0 && H > 0 && W > 0, TMX_ERR_SHAPE, "tmx_conv2d_dgrad: bad shape");
  TMX_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, TMX_ERR_SHAPE,
              "tmx_conv2d_dgrad: Cin=%d and Cout=%d must be multiples of 16", Cin, Cout);
  const int kc = Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16);   // contraction runs over Cout
  const int gw = Cin % 32 == 0 ? 32 : 16;
  int bnc;
  if (Cin % 256 == 0) bnc = 256;
  else if (Cin % 128 == 0) bnc = 128;
  else if (Cin % 64 == 0) bnc = 64;
  else if (Cin % 32 == 0) bnc = 32;
  else bnc = 16;
  const long long rows = (long long)N * (H + 4) * (W + 4);
  TMX_REQUIRE(rows < (1ll << 31) - 4096, TMX_ERR_SHAPE, "tmx_conv2d_dgrad: grid too large");
  ConvTcParams p = {};
  p.N = N;
  p.H = H;
  p.W = W;
  p.Cin = Cout;      // GEMM K per tap
  p.Cout = Cin;      // GEMM N
  p.k = k;
  p.taps = k * k;
  p.pad_off = 0;
  p.bw = kTileM;     // unused geometry of the 4-D mode: keep the divisions harmless
  p.bh = 1;
  p.bn = 1;
  p.tiles_x = 1;
  p.tiles_y = 1;
  p.tiles_n = 1;
  p.cout_log = Cin;
  p.lin = 1;
  p.lin_pitch = W + 4;
  p.lin_rows = rows;
  p.y_f32 = g_f32;
  const long long tiles_m = (rows + kTileM - 1) / kTileM;
  CUtensorMap maps[6];
  int rc;
  if ((rc = encode_rows_map(h, &maps[0], dz_hi, rows, Cout, kc))) return rc;
  if ((rc = encode_rows_map(h, &maps[1], dz_lo, rows, Cout, kc))) return rc;
  return schedule_and_launch(h, p, maps, wt_hi, wt_lo, tiles_m, Cin, Cout, bnc, kc, gw, st);
}

// ---------------------------------------------------------------- data gradient + tmx_grad_prepare in one kernel
// The LIN-mode data gradient whose epilogue finishes what tmx_grad_prepare would do with its output for every
// interior pixel no ring value folds onto: second addend, leaky-ReLU mask of the producing layer, bias gradient, re-split
// into the dz planes (+ fp32 copy).  Ring values and the interior rows / columns the REFLECT / REPLICATE adjoint folds
// them onto still go to the fp32 grid buffer; tmx_grad_border (backward.cu) finishes those 2(H + W) - 4 pixels per
// image.  *served = 0 (nothing launched): shape not covered here - run tmx_conv2d_dgrad + tmx_grad_prepare.
int tmx_conv2d_dgrad_gp_tc(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                           const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                           const tmx_grad_desc_t* gd, const tmx_grad_io_t* gio, cudaStream_t st, int* served) {
  *served = 0;
  if (!(k == 1 || k == 3) || Cin % 32 != 0 || Cin > 512 || Cout % 32 != 0) return TMX_OK;
  const int kc = Cout % 64 == 0 ? 64 : 32;
  int bnc;
  if (Cin % 256 == 0) bnc = 256;
  else if (Cin % 128 == 0) bnc = 128;
  else if (Cin % 64 == 0) bnc = 64;
  else bnc = 32;
  const long long rows = (long long)N * (H + 4) * (W + 4);
  TMX_REQUIRE(rows < (1ll << 31) - 4096, TMX_ERR_SHAPE, "tmx_conv2d_dgrad_gp: grid too large");
  ConvTcParams p = {};
  p.N = N;
  p.H = H;
  p.W = W;
  p.Cin = Cout;      // GEMM K per tap
  p.Cout = Cin;      // GEMM N
  p.k = k;
  p.taps = k * k;
  p.pad_off = 0;
  p.bw = kTileM;
  p.bh = 1;
  p.bn = 1;
  p.tiles_x = 1;
  p.tiles_y = 1;
  p.tiles_n = 1;
  p.cout_log = Cin;
  p.lin = 1;
  p.lin_pitch = W + 4;
  p.lin_rows = rows;
  p.gp = tmx_gp_params(H, W, g_f32, gd, gio);
  const long long tiles_m = (rows + kTileM - 1) / kTileM;
  CUtensorMap maps[6];
  int rc;
  if ((rc = encode_rows_map(h, &maps[0], dz_hi, rows, Cout, kc))) return rc;
  if ((rc = encode_rows_map(h, &maps[1], dz_lo, rows, Cout, kc))) return rc;
  if ((rc = schedule_and_launch(h, p, maps, wt_hi, wt_lo, tiles_m, Cin, Cout, bnc, kc, 32, st, 1, true))) return rc;
  *served = 1;
  return TMX_OK;
}
