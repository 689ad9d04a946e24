// conv_lin.cu — data gradient of the THIN 3x3 layers (Cout <= 64, Cin <= 64) on tcgen05: LIN-PATCH form.
//
// Same GEMM as LIN mode of conv_tc.cu (tmx_conv2d_dgrad): for every row m of the zero-ringed grid [N][H+4][W+4]
//     g[m][ci] = sum_{tap (U,V)} sum_co dz[m + (U-1)*pitch + (V-1)][co] * wt[ci][tap*Cout + co]
// but there a 128-row tile loads NINE row-shifted copies of its operand (one TMA box per tap), which made the
// 16-channel layers load-bound (103 us for 16 -> 16 at 128^2, batch 32).  Here a tile loads ONE contiguous patch of
// R = 128 + 2*pitch + 2 grid rows per plane and every tap is a different START ROW inside it:
//   * operands live in shared memory in the canonical NO-SWIZZLE K-major layout  [K/8 chunks][rows][16 B]
//     (core matrix = 8 consecutive rows x 16 B; SBO = 128 B between 8-row groups, LBO = rows*16 B between the two
//     K-chunks of one MMA), which - unlike the swizzled layouts - accepts any 16-byte aligned start address, i.e. any
//     row shift; TMA writes it with one [8 channels][rows] box per chunk (out-of-range rows arrive as zeros);
//   * the weights (all 9 taps, both planes: <= 74 KB) are loaded ONCE per CTA and stay resident, hi and lo rows of a
//     chunk next to each other so that ONE MMA with N = 2*Cin yields x_hi*w_hi and x_hi*w_lo (see below).
// What bounds these layers after that (measured with the loads / MMAs / stores switched off one at a time): every
// tcgen05.mma with a small N costs ~80-115 cycles whatever N is - the shared-memory read of its 4 KB A operand - and
// neither more accumulators nor a different loader (cp.async warps were tried) change it.  bf16x3 therefore runs as
// TWO MMAs per K step instead of three:  D[:, 0:2N] += A_hi * [B_hi; B_lo]^T  and  D2[:, 0:N] += A_lo * B_hi^T ; the
// epilogue adds the three column groups.  Output: fp32 rows [rows][Cin].
#include "tc_common.cuh"

namespace {

constexpr int kLinM = 128;
constexpr int kLinThreads = 256;

struct LinPatchParams {
  int K;               // contraction per tap = Cout of the layer
  int pitch;           // W + 4
  long long rows;      // N * (H+4) * (W+4)
  int tiles;
  int rbox, nbox;      // the patch is loaded as nbox boxes of rbox rows (TMA box dimension <= 256)
  int ralloc;          // rbox * nbox >= R: rows per chunk in shared memory
  int a_plane_bytes;   // ralloc * K * 2
  int b_bytes;         // both planes: 2 * BN * 9 * K * 2
  int stages;
  float* y;
  GpParams gp;         // GP epilogue (tmx_conv2d_dgrad_gp): tmx_grad_prepare fused in, see tc_common.cuh
};

// no-swizzle K-major shared-memory descriptor (cute::UMMA Major-K INTERLEAVE: ((8,n),2):((1,SBO),LBO) in 16-B units)
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc_lin(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kLinM >> 4) << 24);
}

template <int BN, bool GP = false>
__global__ void __launch_bounds__(kLinThreads, 1)
    conv_lin_patch_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                          const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                          const LinPatchParams p) {
  constexpr int kAccCols = 3 * BN;      // [x_hi*w_hi | x_hi*w_lo | x_lo*w_hi]
  constexpr int kTmemCols = (2 * kAccCols) <= 128 ? 128 : ((2 * kAccCols) <= 256 ? 256 : 512);
  constexpr int GW = BN >= 32 ? 32 : 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sb = smem;                                    // weights: [9K/8 chunks][hi rows | lo rows][16 B]
  uint8_t* sa0 = smem + ((p.b_bytes + 1023) & ~1023);    // stages of [A hi | A lo]
  const int a_stage = 2 * p.a_plane_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sa0 + (size_t)p.stages * a_stage);
  uint64_t* empty_bar = full_bar + 16;
  uint64_t* tfull_bar = empty_bar + 16;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* b_bar = tempty_bar + 2;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(b_bar + 1);
  float* gp_bias_s = reinterpret_cast<float*>(full_bar) + 128;      // 512 B into the 1-KB barrier block: [BN <= 64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = p.K / 8;
  tmx_pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr_s);
  if (GP && threadIdx.x >= 128 && threadIdx.x < 128 + BN) gp_bias_s[threadIdx.x - 128] = 0.f;
  tmx_pdl_wait();      // set-up above overlaps the previous kernel's tail; its results are read from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // weights: all taps, resident for the whole kernel; chunk q -> [BN hi rows | BN lo rows][16 B]
      mbar_arrive_expect_tx(b_bar, p.b_bytes);
      for (int q = 0; q < 9 * kchunks; ++q) {
        tma_load_2d(sb + (size_t)q * 2 * BN * 16, &tm_b_hi, b_bar, q * 8, 0);
        tma_load_2d(sb + (size_t)q * 2 * BN * 16 + BN * 16, &tm_b_lo, b_bar, q * 8, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int row0 = tile * kLinM - p.pitch - 1;             // first grid row of the patch (may be < 0: zeros)
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = sa0 + (size_t)stage * a_stage;
        mbar_arrive_expect_tx(&full_bar[stage], a_stage);
        for (int q = 0; q < kchunks; ++q) {
          for (int b = 0; b < p.nbox; ++b) {
            uint8_t* dst = sa + ((size_t)q * p.ralloc + (size_t)b * p.rbox) * 16;
            tma_load_2d(dst, &tm_a_hi, &full_bar[stage], q * 8, row0 + b * p.rbox);
            tma_load_2d(dst + p.a_plane_bytes, &tm_a_lo, &full_bar[stage], q * 8, row0 + b * p.rbox);
          }
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_lin(BN), idesc2 = make_idesc_lin(2 * BN);
      const uint32_t a_lbo = (uint32_t)p.ralloc * 16u, b_lbo = (uint32_t)(2 * BN) * 16u;
      const uint32_t sb_u = smem_u32(sb);
      mbar_wait(b_bar, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * kAccCols);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(sa0 + (size_t)stage * a_stage);
        uint32_t first = 0;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int u = tap / 3, v = tap - u * 3;
          const uint32_t arow = (uint32_t)(u * p.pitch + v) * 16u;      // start row of this tap inside the patch
          for (int kk = 0; kk < p.K / 16; ++kk) {
            const uint32_t aoff = arow + (uint32_t)(2 * kk) * a_lbo;
            const uint32_t boff = (uint32_t)((tap * p.K) / 8 + 2 * kk) * b_lbo;
            const uint64_t a_hi = make_desc_nosw(sa + aoff, a_lbo);
            const uint64_t a_lo = make_desc_nosw(sa + p.a_plane_bytes + aoff, a_lbo);
            const uint64_t b_hl = make_desc_nosw(sb_u + boff, b_lbo);     // rows 0..BN-1 = hi, BN..2BN-1 = lo
            umma_bf16(tmem_d + 2 * BN, a_lo, b_hl, idesc1, first);        // x_lo * w_hi
            umma_bf16(tmem_d, a_hi, b_hl, idesc2, first);                 // x_hi * [w_hi; w_lo]
            first = 1;
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[as]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: fp32 rows =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const long long mlin = (long long)tile * kLinM + r;
      const bool valid = mlin < p.rows;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      GpRow gp_row;
      if (GP) gp_row = gp_classify(p.gp, mlin, valid);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kAccCols);
#pragma unroll 1
      for (int g = 0; g < BN / GW; ++g) {
        uint32_t a0[32], a1[32], a2[32];
        if (GW == 32) {
          tmem_ld32(taddr0 + g * GW, a0);
          tmem_ld32(taddr0 + BN + g * GW, a1);
          tmem_ld32(taddr0 + 2 * BN + g * GW, a2);
        } else {
          tmem_ld16(taddr0 + g * GW, a0);
          tmem_ld16(taddr0 + BN + g * GW, a1);
          tmem_ld16(taddr0 + 2 * BN + g * GW, a2);
        }
        tmem_ld_wait();
        if constexpr (GP) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = j < GW ? (__uint_as_float(a2[j]) + __uint_as_float(a1[j])) + __uint_as_float(a0[j]) : 0.f;
          gp_group<GW>(p.gp, gp_row, mlin, BN, g * GW, v, gp_bias_s, lane);
        } else if (valid) {
          float v[GW];
#pragma unroll
          for (int j = 0; j < GW; ++j)      // small terms first, then the dominant hi*hi product
            v[j] = (__uint_as_float(a2[j]) + __uint_as_float(a1[j])) + __uint_as_float(a0[j]);
          float4* op = reinterpret_cast<float4*>(p.y + mlin * BN + g * GW);
#pragma unroll
          for (int j = 0; j < GW / 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (GP) gp_flush_bias(p.gp, gp_bias_s, BN);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// [rows][C] bf16 matrix, box = 8 channels (16 B) x `brows` rows, no swizzle: lands as [brows][16 B]
int encode_chunk_map(tmx_handle_t h, CUtensorMap* m, const uint16_t* base, long long rows, int C, int brows) {
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {8u, (cuuint32_t)brows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tmx_fail(TMX_ERR_DRIVER, "cuTensorMapEncodeTiled(lin patch) failed: CUresult %d (rows=%lld C=%d box rows %d)",
                    (int)r, rows, C, brows);
  return TMX_OK;
}

constexpr int kLinSmemBudget = 220 * 1024;

bool lin_patch_plan(int W, int Cin, int Cout, long long rows, LinPatchParams& p, int& smem_bytes) {
  if (!(Cin == 16 || Cin == 32 || Cin == 64) || !(Cout == 16 || Cout == 32 || Cout == 64)) return false;
  p.K = Cout;
  p.pitch = W + 4;
  p.rows = rows;
  p.tiles = (int)((rows + kLinM - 1) / kLinM);
  const int R = kLinM + 2 * p.pitch + 2;
  p.nbox = (R + 255) / 256;
  p.rbox = (R + p.nbox - 1) / p.nbox;
  p.rbox = (p.rbox + 7) & ~7;                    // whole 8-row groups per box
  p.ralloc = p.rbox * p.nbox;
  p.a_plane_bytes = p.ralloc * p.K * 2;
  p.b_bytes = 2 * Cin * 9 * p.K * 2;
  const int b_bytes = (p.b_bytes + 1023) & ~1023;
  const int a_stage = 2 * p.a_plane_bytes;
  int stages = (kLinSmemBudget - b_bytes - 2048) / a_stage;
  if (stages > 8) stages = 8;
  if (stages < 2) return false;
  p.stages = stages;
  smem_bytes = b_bytes + stages * a_stage + 1024 /*barriers*/ + 1024 /*align slack*/;
  return true;
}

template <int BN, bool GP = false>
int launch_lin_patch(tmx_handle_t h, const CUtensorMap* maps, const LinPatchParams& p, int smem_bytes, cudaStream_t st) {
  auto kern = conv_lin_patch_kernel<BN, GP>;
  static thread_local int configured_device = -1;
  if (configured_device != h->device) {
    TMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinSmemBudget + 4096));
    configured_device = h->device;
  }
  const int grid = p.tiles < h->sm_count ? p.tiles : h->sm_count;
  TMX_CUDA(tmx_launch_pdl(kern, dim3(grid), dim3(kLinThreads), (size_t)smem_bytes, st, 1, maps[0], maps[1], maps[2],
                          maps[3], p));
  TMX_LAUNCHED(h, "conv_lin_patch_kernel");
  return TMX_OK;
}

}  // namespace

// returns TMX_OK and sets *served = 1 when the shape was handled here; *served = 0: use the general LIN mode
// gd / gio != NULL: GP form (tmx_conv2d_dgrad_gp) - the epilogue runs tmx_grad_prepare on the accumulators
int tmx_conv2d_dgrad_lin_patch(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, const uint16_t* dz_hi,
                               const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                               cudaStream_t st, int* served, const tmx_grad_desc_t* gd, const tmx_grad_io_t* gio) {
  *served = 0;
  if (tmx_env_flag("TMX_NO_LIN_PATCH")) return TMX_OK;
  const long long rows = (long long)N * (H + 4) * (W + 4);
  LinPatchParams p;
  int smem_bytes = 0;
  if (!lin_patch_plan(W, Cin, Cout, rows, p, smem_bytes)) return TMX_OK;
  if (smem_bytes > h->max_smem_optin || smem_bytes > kLinSmemBudget + 4096) return TMX_OK;
  p.y = g_f32;
  CUtensorMap maps[4];
  int rc;
  if ((rc = encode_chunk_map(h, &maps[0], dz_hi, rows, Cout, p.rbox))) return rc;
  if ((rc = encode_chunk_map(h, &maps[1], dz_lo, rows, Cout, p.rbox))) return rc;
  if ((rc = encode_chunk_map(h, &maps[2], wt_hi, Cin, 9 * Cout, Cin))) return rc;
  if ((rc = encode_chunk_map(h, &maps[3], wt_lo, Cin, 9 * Cout, Cin))) return rc;
  if (gd != nullptr) {
    p.gp = tmx_gp_params(H, W, g_f32, gd, gio);
    if (Cin == 16) rc = launch_lin_patch<16, true>(h, maps, p, smem_bytes, st);
    else if (Cin == 32) rc = launch_lin_patch<32, true>(h, maps, p, smem_bytes, st);
    else rc = launch_lin_patch<64, true>(h, maps, p, smem_bytes, st);
  } else if (Cin == 16) rc = launch_lin_patch<16>(h, maps, p, smem_bytes, st);
  else if (Cin == 32) rc = launch_lin_patch<32>(h, maps, p, smem_bytes, st);
  else rc = launch_lin_patch<64>(h, maps, p, smem_bytes, st);
  if (rc) return rc;
  *served = 1;
  return TMX_OK;
}
