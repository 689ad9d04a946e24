// tc_common.cuh — inline-PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 alloc / mma / commit / ld, and their CTA-pair (cta_group::2) forms.
#pragma once

#include "common.cuh"

namespace {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// --- CTA-pair (cta_group::2) variants.  Addresses of CTA-local shared memory double as shared::cluster
// addresses of the own CTA; clearing bit 24 (the peer bit of a 2-CTA cluster) names the same offset in
// the even (leader) CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {  // data into the own CTA, completion on the leader's barrier
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N rows: N/2 per CTA]^T
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once all prior MMAs of the pair have completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---------------------------------------------------------------- GP: tmx_grad_prepare inside a data-gradient epilogue
// Shared by conv_tc.cu (LIN mode) and conv_lin.cu (LIN-PATCH): the thread that owns grid row `mlin` of the zero-ringed
// grid [N][H+4][W+4] finishes, GW channels at a time, what tmx_grad_prepare(src_kind 0) would do with the accumulator:
// see tmx_conv2d_dgrad_gp in include/tmx.h.
struct GpParams {
  int H, W, fold, mask_kind;     // fold: 0 REFLECT, 1 REPLICATE, 2 none; mask_kind: 0 none, 1 fp32 NHWC, 2 bf16 hi plane
  float alpha, dbias_scale;
  const float* add;              // fp32 NHWC [N][H][W][C] or NULL
  const void* mask;
  float* grid;                   // fp32 [N][H+4][W+4][C]: ring next to the interior + the pixels it folds onto
  uint16_t* dz_hi;               // planes on the grid (may be NULL: fp32 output only)
  uint16_t* dz_lo;
  float* f32;                    // fp32 NHWC output or NULL
  float* dbias;                  // [C] or NULL
};

// what a grid row is: -1 past the end, 0 outer / unused ring (zero planes), 1 ring next to the interior that the padding
// adjoint reads (raw value to the grid buffer + zero planes), 2 interior pixel a ring value folds onto (raw value to
// the grid buffer; grad_border_kernel finishes it), 3 any other interior pixel (finished by the epilogue)
struct GpRow {
  int cls;
  long long pix, mpix;           // NHWC pixel index; pixel index inside the haloed [N][H+2][W+2] mask planes
};

__device__ __forceinline__ GpRow gp_classify(const GpParams& q, long long mlin, bool valid) {
  GpRow row;
  row.cls = -1;
  row.pix = row.mpix = 0;
  if (!valid) return row;
  const int Wq = q.W + 4, HWq = (q.H + 4) * Wq;
  const int n_ = (int)(mlin / HWq);
  const int rem = (int)(mlin - (long long)n_ * HWq);
  const int rr = rem / Wq;
  const int r_ = rr - 2, c_ = rem - rr * Wq - 2;
  if (r_ >= 0 && r_ < q.H && c_ >= 0 && c_ < q.W) {
    const bool dirty = q.fold == 0 ? (r_ == 1 || r_ == q.H - 2 || c_ == 1 || c_ == q.W - 2)
                                   : (q.fold == 1 ? (r_ == 0 || r_ == q.H - 1 || c_ == 0 || c_ == q.W - 1) : false);
    row.cls = dirty ? 2 : 3;
    row.pix = ((long long)n_ * q.H + r_) * q.W + c_;
    row.mpix = ((long long)n_ * (q.H + 2) + r_ + 1) * (q.W + 2) + c_ + 1;
  } else {
    row.cls = (q.fold != 2 && r_ >= -1 && r_ <= q.H && c_ >= -1 && c_ <= q.W) ? 1 : 0;
  }
  return row;
}

// column sums of a 32 x 32 tile held one row per lane: lane L returns sum over the warp's lanes of v[L] (31 shuffles:
// each stage keeps the half of the columns whose index has the stage's bit equal to the lane's)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < o; ++j) {
      const float send = up ? v[j] : v[j + o];
      const float keep = up ? v[j + o] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// One group of GW channels [col0, col0 + GW) of grid row mlin; v[0..GW) = the data gradient (v[GW..32) = 0), C = channel
// count of the activation; bias_s = the CTA's shared bias-gradient partials [C].  Called by all 32 lanes of a warp.
template <int GW>
__device__ __forceinline__ void gp_group(const GpParams& q, const GpRow& row, long long mlin, int C, int col0,
                                         float (&v)[32], float* bias_s, int lane) {
  if (row.cls == 1 || row.cls == 2) {
    float4* op = reinterpret_cast<float4*>(q.grid + mlin * C + col0);
#pragma unroll
    for (int j = 0; j < GW / 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if ((row.cls == 0 || row.cls == 1) && q.dz_hi != nullptr) {
    uint4* oh = reinterpret_cast<uint4*>(q.dz_hi + mlin * C + col0);
    uint4* ol = reinterpret_cast<uint4*>(q.dz_lo + mlin * C + col0);
#pragma unroll
    for (int j = 0; j < GW / 8; ++j) {
      oh[j] = make_uint4(0u, 0u, 0u, 0u);
      ol[j] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (row.cls == 3) {
    if (q.add != nullptr) {
      const float4* ap = reinterpret_cast<const float4*>(q.add + row.pix * C + col0);
#pragma unroll
      for (int j = 0; j < GW / 4; ++j) {
        const float4 a4 = __ldg(ap + j);
        v[4 * j] += a4.x;
        v[4 * j + 1] += a4.y;
        v[4 * j + 2] += a4.z;
        v[4 * j + 3] += a4.w;
      }
    }
    if (q.mask_kind == 1) {
      const float4* yp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q.mask) + row.pix * C + col0);
#pragma unroll
      for (int j = 0; j < GW / 4; ++j) {
        const float4 y4 = __ldg(yp + j);
        v[4 * j] *= y4.x > 0.f ? 1.f : q.alpha;
        v[4 * j + 1] *= y4.y > 0.f ? 1.f : q.alpha;
        v[4 * j + 2] *= y4.z > 0.f ? 1.f : q.alpha;
        v[4 * j + 3] *= y4.w > 0.f ? 1.f : q.alpha;
      }
    } else if (q.mask_kind == 2) {
      const uint4* yp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(q.mask) + row.mpix * C + col0);
#pragma unroll
      for (int j = 0; j < GW / 8; ++j) {
        const uint4 yb = __ldg(yp + j);
        const uint32_t w4[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {   // y > 0  <=>  the bf16 word, moved to the top half, is a positive int32
          v[8 * j + 2 * t] *= ((int)(w4[t] << 16) > 0) ? 1.f : q.alpha;
          v[8 * j + 2 * t + 1] *= ((int)(w4[t] & 0xffff0000u) > 0) ? 1.f : q.alpha;
        }
      }
    }
    if (q.f32 != nullptr) {
      float4* op = reinterpret_cast<float4*>(q.f32 + row.pix * C + col0);
#pragma unroll
      for (int j = 0; j < GW / 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    if (q.dz_hi != nullptr) {
      uint4* oh = reinterpret_cast<uint4*>(q.dz_hi + mlin * C + col0);
      uint4* ol = reinterpret_cast<uint4*>(q.dz_lo + mlin * C + col0);
#pragma unroll
      for (int j = 0; j < GW / 8; ++j) {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) tmx_split_bf16x2(v[8 * j + 2 * t], v[8 * j + 2 * t + 1], ph[t], pl[t]);
        oh[j] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        ol[j] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if (q.dbias != nullptr) {      // (uniform branch: every lane takes part in the shuffles)
    const float cs = warp_colsum32(v, lane);
    if (lane < GW) atomicAdd(&bias_s[col0 + lane], cs);
  }
}

// after the last tile: one global atomic per channel and CTA (the 128 epilogue threads, warps 4-7, meet first)
__device__ __forceinline__ void gp_flush_bias(const GpParams& q, const float* bias_s, int C) {
  if (q.dbias == nullptr) return;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int c = threadIdx.x - 128; c < C; c += 128) {
    const float bs = bias_s[c];
    if (bs != 0.f) atomicAdd(q.dbias + c, bs * q.dbias_scale);
  }
}

static inline GpParams tmx_gp_params(int H, int W, float* grid, const tmx_grad_desc_t* gd, const tmx_grad_io_t* gio) {
  GpParams q;
  q.H = H;
  q.W = W;
  q.fold = gd->fold;
  q.mask_kind = gd->mask_kind;
  q.alpha = gd->alpha;
  q.dbias_scale = gd->dbias_scale;
  q.add = gio->add;
  q.mask = gio->y_mask;
  q.grid = grid;
  q.dz_hi = gio->dz_hi;
  q.dz_lo = gio->dz_lo;
  q.f32 = gio->dz_f32;
  q.dbias = gio->dbias;
  return q;
}

}  // namespace
