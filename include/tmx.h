/* tmx.h — C ABI of libtmx.so: the B200 (sm_100a) implementation of the
 * TextureMixer hot path (conv encoders E_zg/E_zl -> latent tile blend ->
 * residual generator G_res, discriminator D_patch).
 *
 * This is the drop-in boundary.  The reference has no FFI of its own (it is
 * pure Python on TF 1.12); each entry point names the reference op it stands
 * in for (file:line into ningyu1991/TextureMixer).  INTEGRATION.md shows the
 * ctypes stub a reference maintainer adds.
 *
 * Conventions
 *  - every function returns int: 0 = OK, <0 = argument/shape/arch error,
 *    >0 = cudaError_t; never throws, never aborts.  tmx_last_error() returns a
 *    thread-local human readable message for the last non-zero return.
 *  - the CALLER owns every buffer (device pointers, plain sizes); the library
 *    allocates nothing persistent except the opaque handle.
 *  - all work is enqueued asynchronously on the caller's cudaStream_t (passed
 *    as void*); no hidden synchronisation.  One handle per (device, thread).
 *  - activations are fp32.  API-side layout is the reference's NCHW; the
 *    internal layouts are
 *       NHWC f32         [N][H][W][C]
 *       SPLIT_BF16_HALO  two bf16 planes hi, lo with x ~= hi + lo (16 mantissa
 *                        bits), each [N][H+2][W+2][C] with the REFLECT halo of
 *                        networks.py:55 materialised (row -1 = row 1, ...), so
 *                        that TMA boxes can fetch shifted 3x3 taps.
 */
#ifndef TMX_H_
#define TMX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMX_ABI_VERSION 6

typedef struct tmx_ctx* tmx_handle_t;
typedef void* tmx_stream_t; /* cudaStream_t */

/* error codes (<0) */
#define TMX_OK 0
#define TMX_ERR_ARG (-1)      /* null pointer / bad enum */
#define TMX_ERR_SHAPE (-2)    /* unsupported or inconsistent shape */
#define TMX_ERR_ARCH (-3)     /* device is not sm_100 */
#define TMX_ERR_UNSUPPORTED (-4)
#define TMX_ERR_DRIVER (-5)   /* driver entry point (cuTensorMapEncodeTiled) failed */

int tmx_abi_version(void);
const char* tmx_last_error(void);
/* Create/destroy the per-device context (queries SM count, resolves the
 * tensor-map encoder).  Fails with TMX_ERR_ARCH on anything but sm_100. */
int tmx_create(int device, tmx_handle_t* out);
int tmx_destroy(tmx_handle_t h);
int tmx_device_info(tmx_handle_t h, int* sm_count, int* cc_major, int* cc_minor);
/* Number of kernel launches this handle has enqueued so far (bench.py's gpu_launches). */
int tmx_launch_count(tmx_handle_t h, uint64_t* count);

/* ------------------------------------------------------------------ conv2d
 * networks.py:48-56 conv2d (+ :26-33 wscale, :61-67 apply_bias, :72-75
 * leaky_relu, :80-88 upscale2d in front, residual add of :437 behind).
 * y = [residual +] lrelu( wscale * xcorr(reflect_pad(up2?(x)), w) + bias )      */
#define TMX_CONV_LRELU 1u    /* max(alpha*x, x) after bias */
#define TMX_CONV_RESIDUAL 2u /* y += residual (after activation; G_res :437 has none) */
#define TMX_CONV_UP2_IN 4u   /* logical input = nearest-neighbour x2 of the stored input (networks.py:448).
                              * FFMA: the gather reads through the upsampling.
                              * TC: sub-pixel form - conv3x3(up2(x)) == four 2x2-tap convs of x, run as ONE
                              *     3x3-tap GEMM with N = 4*Cout over the stored low-res planes, which must
                              *     carry a REPLICATE halo (REFLECT of the upsampled image == clamp of the
                              *     stored one) and weights prepared with up2_phase = 1. */
#define TMX_CONV_UP2_OUT 8u  /* split-plane output is written x2 upsampled (then REFLECT halo of the big image) */
#define TMX_CONV_HALO_REPLICATE 16u /* split-plane output gets a REPLICATE halo (consumer is a TC UP2_IN conv) */
#define TMX_CONV_TORGB 32u   /* TC, Cout <= 32: additionally y_rgb = [tanh](rgb_wscale * (y . rgb_w) + rgb_b) as NCHW
                              * (networks.py:454-457 torgb + :483 tanh fused behind the last conv) */

#define TMX_CONV_XMERGE 64u  /* TC, Cin == 16, k == 3: operand rows carry the 3 horizontal taps (4 pixels = 128 B) of
                              * an overlapping-stride view; weights from tmx_conv_weights_prepare with xmerge = 1.
                              * The x planes must be allocated (and zero-filled) 64 elements past their end. */
#define TMX_CONV_HALO_ZERO 128u /* split-plane output gets a ZERO halo (written by the border pixels' threads; the
                                 * consumer is a SAME-padded conv: VGG-19, fused_scale) */

#define TMX_CONV_W_PER_SAMPLE 256u /* TC: w_hi / w_lo hold N weight sets [N][Cout][k*k*Cin], image n is convolved with
                                    * set n (a batched GEMM: the Gram-loss gradient dF[n] = F[n] (S[n] + S[n]^T)).  Needs
                                    * >= 128 pixels per image; no UP2_IN / XMERGE / TORGB. */

#define TMX_ALGO_AUTO 0
#define TMX_ALGO_FFMA 1 /* CUDA-core fp32 implicit GEMM, NHWC f32 in/out */
#define TMX_ALGO_TC 2   /* tcgen05 bf16x3 implicit GEMM, SPLIT_BF16_HALO in; K chunk = 64/32/16 channels by Cin */
#define TMX_ALGO_TC_K32 3 /* same, K chunks of at most 32 channels (64B swizzle, deeper pipeline) */

typedef struct {
  int32_t N, H, W;   /* output size == logical input size (stride 1, 'same') */
  int32_t Cin, Cout;
  int32_t k;         /* 1 (VALID) or 3 (REFLECT pad 1) */
  uint32_t flags;    /* TMX_CONV_* */
  int32_t algo;      /* TMX_ALGO_* */
  float wscale;      /* gain / sqrt(k*k*Cin), networks.py:28-30 (FFMA applies it; TC expects it folded by tmx_conv_weights_prepare) */
  float lrelu_alpha; /* 0.2 */
  int32_t rgb_cout;  /* TORGB: number of image channels (<= 4) */
  int32_t rgb_tanh;  /* TORGB: apply tanh */
  float rgb_wscale;  /* TORGB: gain / sqrt(Cout) of the 1x1 head */
} tmx_conv_desc_t;

typedef struct {
  const float* x_f32;    /* FFMA: NHWC f32 [N][H][W][Cin] ([N][H/2][W/2][Cin] with UP2_IN) */
  const uint16_t* x_hi;  /* TC: SPLIT_BF16_HALO planes [N][H+2][W+2][Cin] ([N][H/2+2][W/2+2][Cin] with UP2_IN) */
  const uint16_t* x_lo;
  const float* w;        /* FFMA: raw variable, HWIO [k][k][Cin][Cout] */
  const uint16_t* w_hi;  /* TC: prepared planes [Cout][k*k*Cin] ([4*Cout][9*Cin] with UP2_IN) */
  const uint16_t* w_lo;
  const float* bias;     /* [Cout] or NULL */
  const float* residual; /* NHWC f32 [N][H][W][Cout] or NULL */
  float* y_f32;          /* NHWC f32 [N][H][W][Cout] or NULL */
  uint16_t* y_hi;        /* SPLIT_BF16_HALO out [N][H+2][W+2][Cout] ([N][2H+2][2W+2][Cout] with UP2_OUT) or NULL */
  uint16_t* y_lo;
  const float* rgb_w;    /* TORGB: raw 1x1 head variable [Cout][rgb_cout] */
  const float* rgb_b;    /* TORGB: [rgb_cout] or NULL */
  float* y_rgb;          /* TORGB: NCHW f32 [N][rgb_cout][H][W] */
} tmx_conv_io_t;

int tmx_conv2d_fwd(tmx_handle_t h, const tmx_conv_desc_t* d, const tmx_conv_io_t* io, tmx_stream_t s);

/* get_weight (networks.py:26-33) for the tensor-core path: w_hwio * wscale ->
 * bf16 hi/lo planes laid out K-major [Cout][k*k*Cin], K index = (u*k+v)*Cin + c.
 * up2_phase (k == 3 only): the sub-pixel weights of conv3x3(upscale2d(x)) as [4*Cout][9*Cin]: row
 * (a*2+b)*Cout + o holds, for output phase (a,b) = (row parity, column parity), the 3x3 low-res taps
 * (U,V): sum of w[u][v] over the upsampled taps that fall on low-res offset (U-1, V-1)
 * (a=0: {0},{1,2},{} ; a=1: {},{0,1},{2}); fp32 sums, then * wscale, then split.
 * Cin_pad >= Cin: the planes are laid out for an activation whose channel count was zero-padded to
 * Cin_pad (a multiple of 16), e.g. the 513-channel minibatch-stddev output; K index = tap*Cin_pad + c. */
int tmx_conv_weights_prepare(tmx_handle_t h, const float* w_hwio, float wscale, int k, int Cin, int Cin_pad, int Cout,
                             int up2_phase, uint16_t* w_hi, uint16_t* w_lo, tmx_stream_t s);
/* XMERGE layout for 16-channel 3x3 layers: [Cout][3 (u)][64], K index u*64 + v*16 + c for v < 3, zeros for v == 3
 * and for c >= Cin (w_hwio is [3][3][Cin][Cout], Cin <= 16: an input zero-padded to 16 channels). */
int tmx_conv_weights_prepare_xmerge(tmx_handle_t h, const float* w_hwio, float wscale, int Cin, int Cout,
                                    uint16_t* w_hi, uint16_t* w_lo, tmx_stream_t s);

/* NHWC f32 -> SPLIT_BF16_HALO (tf.pad REFLECT of networks.py:55 materialised; replicate != 0: edge-clamped halo). */
int tmx_split_halo_pack(tmx_handle_t h, const float* x_nhwc, uint16_t* hi, uint16_t* lo, int N, int H, int W, int C,
                        int replicate, tmx_stream_t s);
/* SPLIT_BF16_HALO interior -> NHWC f32 (hi + lo), for tests and FFMA consumers. */
int tmx_split_halo_unpack(tmx_handle_t h, const uint16_t* hi, const uint16_t* lo, float* y_nhwc, int N, int H, int W,
                          int C, tmx_stream_t s);

/* ------------------------------------------------------------------ pointwise / layout
 * FromRGB: 1x1 conv from NCHW images (networks.py:226-228 fromrgb): y NHWC f32. */
int tmx_fromrgb_fwd(tmx_handle_t h, const float* x_nchw, const float* w /*[Cin][Cout]*/, const float* bias,
                    float wscale, float* y_nhwc, int N, int Cin, int H, int W, int Cout, int lrelu, float alpha,
                    tmx_stream_t s);
/* ToRGB: 1x1 conv to NCHW images + optional tanh (networks.py:454-457, :483). */
int tmx_torgb_fwd(tmx_handle_t h, const float* x_nhwc, const float* w /*[Cin][Cout]*/, const float* bias,
                  float wscale, float* y_nchw, int N, int H, int W, int Cin, int Cout, int apply_tanh,
                  tmx_stream_t s);
/* downscale2d (networks.py:131-136): 2x2 mean, NHWC f32 [N][H][W][C] -> [N][H/2][W/2][C]. */
int tmx_avgpool2_fwd(tmx_handle_t h, const float* x, float* y, int N, int H, int W, int C, tmx_stream_t s);
/* the same pooling with the result ALSO (y may be NULL: only) written as SPLIT_BF16_HALO planes [N][H/2+2][W/2+2][C] for a
 * tensor-core consumer (halo_kind as in tmx_split_halo_pack): downscale2d -> conv2d of networks.py:131-136, 48-56 without
 * the intermediate layout pass.  H, W even and >= 4, C % 8 == 0. */
int tmx_avgpool2_pack(tmx_handle_t h, const float* x, float* y, uint16_t* hi, uint16_t* lo, int N, int H, int W, int C,
                      int halo_kind, tmx_stream_t s);
/* Layout moves between the reference's NCHW and the internal NHWC.  The NHWC
 * side may be a channel slice [c_off, c_off+C) of a tensor with C_total channels
 * (tf.concat of networks.py:423; mu/log_sigma split of :289-290, :381-382).
 * If bcast_hw != 0 the NCHW source is [N][C][1][1] broadcast over H x W
 * (tf.tile of loss.py:130). */
int tmx_nchw_to_nhwc(tmx_handle_t h, const float* x_nchw, float* y_nhwc, int N, int C, int H, int W, int c_off,
                     int C_total, int bcast_hw, tmx_stream_t s);
int tmx_nhwc_to_nchw(tmx_handle_t h, const float* x_nhwc, float* y_nchw, int N, int C, int H, int W, int c_off,
                     int C_total, tmx_stream_t s);

/* ------------------------------------------------------------------ discriminator head (D_patch)
 * minibatch_stddev_layer (networks.py:177-189): x NHWC [N][H][W][C] -> y NHWC [N][H][W][C_total] with
 * y[..., :C] = x, y[..., C] = group statistic (groups of min(group_size, N) samples {m, m+M, ...}),
 * y[..., C+1:] = 0 (channel padding for the consumer conv).  stat: [N / G] floats of scratch/output. */
int tmx_mbstd_fwd(tmx_handle_t h, const float* x_nhwc, float* y_nhwc, float* stat, int N, int H, int W, int C,
                  int C_total, int group_size, tmx_stream_t s);
/* dense + apply_bias + leaky_relu (networks.py:38-43, 61-67, 72-75): y[N][Cout] = act(wscale * x[N][K] . w[K][Cout] + b).
 * Deterministic split-K through a caller-owned workspace of tmx_dense_workspace_bytes(). */
int tmx_dense_workspace_bytes(int N, int K, int Cout, size_t* bytes);
int tmx_dense_fwd(tmx_handle_t h, const float* x, const float* w, const float* bias, float wscale, float* y,
                  float* workspace, int N, int K, int Cout, int lrelu, float alpha, tmx_stream_t s);

/* ------------------------------------------------------------------ latent tile blend (K6)
 * loss.py:92-100 tiling_permutation + tfutil.py:41-43 lerp + the matte sums of
 * util_scripts.py:496,525 in one pass.  For every canvas element
 *    out[n][c][i][j] = sum_k  m_k(i,j) * src_k[n][c][sy_k(i)][sx_k(j)]
 * with K <= 4 sources, sy_k(i) = idx_h[k][n][i] % h, sx_k(j) = idx_w[k][n][j] % w
 * (identity i % h, j % w when the index pointer is NULL).  Re-pinned tiles:
 * canvas positions whose tile row i/h is in the bit set pin_rows AND whose tile
 * column j/w is in pin_cols read src[i % h][j % w] instead - the four corners
 * of loss.py:96-100 / util_scripts.py:747-750 are rows {0,sh-1} x cols {0,sw-1};
 * the interpolation app (util_scripts.py:1433-1437) pins row sh/2 x cols {0,sw-1}.
 * Sources are NCHW f32 [N][C][h][w] ([N][C][1][1] if src_bcast: the tiled
 * global code).
 * Weights:
 *   mode TMX_BLEND_MATTE: m_k(i,j) = ramp_h[k][i] * ramp_w[k][j] as float64
 *       products/sums in the reference's left-to-right order, one final
 *       rounding to f32 (numpy promotion in util_scripts.py:496) - bit exact.
 *       With math_f32 the ramps hold float32 values and every op is a
 *       separately rounded f32 op (the float32 matte of util_scripts.py:85-89
 *       used at :1337-1342, :1438-1439).
 *   mode TMX_BLEND_LERP : K == 2, out = a + (b - a) * t[n] in f32 with
 *       separately rounded sub/mul/add (tfutil.py:41-43), a = source 0.
 *   mode TMX_BLEND_COPY : K == 1, plain gather.
 * Output: NCHW f32 [N][C][H][W] (the reference-facing canvas) when out_nchw !=
 * NULL, and/or NHWC f32 slice (c_off, C_total) when out_nhwc != NULL. */
#define TMX_BLEND_COPY 0
#define TMX_BLEND_MATTE 1
#define TMX_BLEND_LERP 2

typedef struct {
  int32_t N, C, h, w; /* source tile */
  int32_t H, W;       /* canvas */
  int32_t K;          /* sources, 1..4 */
  int32_t mode;       /* TMX_BLEND_* */
  int32_t math_f32;   /* MATTE: f32 arithmetic instead of f64 */
  int32_t src_bcast;  /* sources are [N][C][1][1] */
  uint32_t src_reverse; /* bit k: source k reads sample N-1-n (tf.reverse(axis=[0]) of loss.py:218,224,236) */
  uint64_t pin_rows, pin_cols; /* tile-row / tile-column bit sets of re-pinned tiles (0 = none) */
  int32_t c_off, C_total; /* NHWC output slice */
} tmx_blend_desc_t;

typedef struct {
  const float* src[4];
  const int32_t* idx_h[4]; /* [N][H] or NULL */
  const int32_t* idx_w[4]; /* [N][W] or NULL */
  const double* ramp_h[4]; /* [H] (MATTE) */
  const double* ramp_w[4]; /* [W] (MATTE) */
  const float* t;          /* [N]  (LERP) */
  float* out_nchw;
  float* out_nhwc;
} tmx_blend_io_t;

int tmx_latent_blend(tmx_handle_t h, const tmx_blend_desc_t* d, const tmx_blend_io_t* io, tmx_stream_t s);

/* ------------------------------------------------------------------ backward (tf.gradients of the ops above,
 * tfutil.py:299).  Gradients w.r.t. layer outputs travel as fp32 NHWC or as bf16 hi/lo planes on a
 * ZERO-RINGED grid [N][H+4][W+4][C] (interior at offset 2) shared by the input and the output of the data
 * gradient, so that a 3x3 tap is a constant row shift of that grid.
 *
 * tmx_conv2d_dgrad: g[n][r][c][ci] = sum_{u,v,co} dz[n][r+1-u][c+1-v][co] * w[u][v][ci][co] * wscale at EVERY grid
 *   position (fp32 [N][H+4][W+4][Cin]); the ring 1..H+2 holds what the padding adjoint folds back, the outermost
 *   ring is garbage.  wt_hi/wt_lo = tmx_conv_weights_transpose of the forward planes: [Cin][k*k*Cout].
 *   For a UP2_IN (sub-pixel) layer pass the low-res H, W, Cout := 4*Cout and phase-packed dz planes. */
int tmx_conv2d_dgrad(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                     const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32, tmx_stream_t s);
/* prepared forward planes [rows][taps*K] -> data-gradient planes [K][taps*rows], taps flipped. */
int tmx_conv_weights_transpose(tmx_handle_t h, const uint16_t* w_hi, const uint16_t* w_lo, int rows, int taps, int K,
                               uint16_t* wt_hi, uint16_t* wt_lo, tmx_stream_t s);

/* tmx_conv2d_wgrad: dw[u][v][ci][co] += wscale * sum_{n,y,x} xpad[n][y+u][x+v][ci] * dz[n][y][x][co]  (ACCUMULATES
 *   into the HWIO gradient of the raw variable).  x_hi/x_lo: the layer's forward input planes (SPLIT_BF16_HALO
 *   [N][H+2][W+2][Cin], the halo kind the forward used); dz_hi/dz_lo: tmx_grad_prepare planes on the zero-ringed
 *   grid [N][H+4][W+4][Cout].  Tensor cores, Cin and Cout multiples of 64; deterministic split-K through a
 *   caller-owned workspace.  For a UP2_IN layer pass low-res H, W, Cout := 4*Cout, phase-packed dz and reduce the
 *   phase gradient with tmx_conv_wgrad_unphase.
 *   flags: TMX_WGRAD_X_SLACK = at least 128 readable bytes follow each x plane; lets the 16/32-channel 3x3 layers
 *   read their horizontal taps through overlapping tensor-map rows (PACKED-M mode, conv_wgrad.cu).  Pass the same
 *   flags to the workspace query. */
#define TMX_WGRAD_X_SLACK 1
int tmx_conv2d_wgrad_workspace_bytes(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, int flags,
                                     size_t* bytes);
int tmx_conv2d_wgrad(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, float wscale, const uint16_t* x_hi,
                     const uint16_t* x_lo, const uint16_t* dz_hi, const uint16_t* dz_lo, float* dw, float* workspace,
                     int flags, tmx_stream_t s);

/* dwp = gradient of the sub-pixel weights as [9][Cin][4*Cout] (what tmx_conv2d_wgrad writes for a UP2_IN layer);
 * dw[u][v][ci][co] += sum of the phase entries tap (u,v) contributed to (adjoint of tmx_conv_weights_prepare up2_phase,
 * without its wscale). */
int tmx_conv_wgrad_unphase(tmx_handle_t h, const float* dwp, float* dw, int Cin, int Cout, tmx_stream_t s);
/* ToRGB + tanh backward (networks.py:454-457, :483): dimg/img NCHW [N][Cimg][H][W], y NHWC [N][H][W][Cin] (the head's
 * input), w raw [Cin][Cimg]; writes dy NHWC, accumulates dw (x wscale) and db. */
int tmx_torgb_bwd(tmx_handle_t h, const float* dimg, const float* img, const float* y, const float* w, float wscale,
                  float* dy, float* dw, float* db, int N, int H, int W, int Cin, int Cimg, int use_tanh, tmx_stream_t s);
/* FromRGB backward (networks.py:226-228): dz = masked gradient NHWC [N][H][W][Cout]; accumulates dw [Cimg][Cout]
 * (x wscale); dimg (NCHW, optional) = gradient w.r.t. the image. */
int tmx_fromrgb_bwd(tmx_handle_t h, const float* img, const float* dz, const float* w, float wscale, float* dw,
                    float* dimg, int N, int Cimg, int H, int W, int Cout, tmx_stream_t s);

/* D_patch head, input gradients: dense (networks.py:38-43; dz = dy * lrelu'(y) when lrelu) and
 * minibatch_stddev_layer (networks.py:177-189; x, dx NHWC [N][H][W][C], dy NHWC [N][H][W][C_total], ds: [N/G] scratch). */
int tmx_dense_bwd_input(tmx_handle_t h, const float* dy, const float* y, const float* w, float wscale, float* dx, int N,
                        int K, int Cout, int lrelu, float alpha, tmx_stream_t s);
int tmx_mbstd_bwd(tmx_handle_t h, const float* x, const float* dy, float* dx, float* ds, int N, int H, int W, int C,
                  int C_total, int group_size, tmx_stream_t s);
/* L1 image loss of EG_wgan (loss.py:142-146): grad = scale * sign(a - b), *loss_sum += sum |a - b| (may be NULL). */
int tmx_loss_l1_grad(tmx_handle_t h, const float* a, const float* b, float* grad, float* loss_sum, int64_t n, float scale,
                     tmx_stream_t s);
/* Adjoint of tmx_latent_blend COPY mode (tiling_permutation, loss.py:92-100): scatter-add of the canvas gradient
 * NCHW [N][C][H][W] into d_src NCHW [N][C][sh][sw] (caller zero-initialises); reverse: source sample N-1-n. */
int tmx_latent_gather_bwd(tmx_handle_t h, const float* dcanvas, float* dsrc, const int32_t* idx_h, const int32_t* idx_w,
                          int N, int C, int sh, int sw, int H, int W, uint64_t pin_rows, uint64_t pin_cols, int reverse,
                          tmx_stream_t s);
/* Sampled latent canvases of the config-off interpolation modes (loss.py:176-193, 218-235:
 * zg_/zl_interp_variational = 'variational' | 'random'; the reference config uses 'hard' / 'permutational').
 * mu, ls: [N][C][sh][sw]; eps: [N][C][eh][ew] standard-normal draws (1 x 1 or H x W); out / g: [N][C][H][W].
 * mode 1: out = eps * exp(ls) + mu (sources tiled);  mode 2: mu on the four corner tiles, eps elsewhere.
 * reverse: sources read batch-reversed.  _bwd ACCUMULATES d out / d mu and (mode 1) d out / d ls. */
int tmx_latent_noise_fwd(tmx_handle_t h, int mode, const float* mu, const float* ls, const float* eps, float* out, int N,
                         int C, int sh, int sw, int eh, int ew, int H, int W, int reverse, tmx_stream_t s);
int tmx_latent_noise_bwd(tmx_handle_t h, int mode, const float* g, const float* ls, const float* eps, float* dmu,
                         float* dls, int N, int C, int sh, int sw, int eh, int ew, int H, int W, int reverse,
                         tmx_stream_t s);
/* The same for a gradient that is non-zero only inside the [wh x ww] window of the canvas at (oy, ox) (crop-aware
 * G_fcn, loss.crop_window): dwin is NCHW [N][C][wh][ww]; off_dev, when non-NULL, holds {oy, ox} on the device. */
int tmx_latent_gather_bwd_window(tmx_handle_t h, const float* dwin, float* dsrc, const int32_t* idx_h,
                                 const int32_t* idx_w, int N, int C, int sh, int sw, int H, int W, int wh, int ww, int oy,
                                 int ox, const int32_t* off_dev, uint64_t pin_rows, uint64_t pin_cols, int reverse,
                                 tmx_stream_t s);
/* out[row] (+)= scale * sum_i f(in[row][i]), f = identity or square: adjoint of tiling a [N][C][1][1] code over a
 * canvas (loss.py:176), loss means, per-sample squared gradient norms (loss.py:334). */
int tmx_row_sum(tmx_handle_t h, const float* in, float* out, int rows, int len, float scale, int accumulate, int square,
                tmx_stream_t s);
/* out = a * in + b over n fp32 elements. */
int tmx_axpb(tmx_handle_t h, const float* in, float* out, int64_t n, float a, float b, tmx_stream_t s);

/* ---- WGAN-GP double backward (loss.py:332-336).  For the piecewise-linear layers the second-order
 * term needs no new conv kernel: d/dtheta [ v . grad_x D ] = weight gradients with the TANGENT activations
 * (forward of the bias-free, mask-frozen network on v) against the adjoints of the first backward.  Only the
 * minibatch-stddev layer has curvature:
 * tmx_mbstd_tangent:  ydot = J_mbstd(x) xdot   ([xdot, sdot broadcast, 0...], NHWC [N][H][W][C_total]; sdot: [N/G])
 * tmx_mbstd_curvature: q = d/dx ( lam[m] * sdot[m] )  with xdot fixed (NHWC [N][H][W][C]; lam: [N/G])
 * tmx_dense_wgrad:    dw[k][o] += wscale * sum_n x[n][k] dz[n][o],  db[o] += sum_n dz[n][o],  dz = dy * lrelu'(y)
 * tmx_scale_rows:     out[n][:] = scale[n] * in[n][:]
 * tmx_gp_coefficients: from squared gradient norms sq[n]: penalty[n] = lambda (sqrt(sq)-target)^2 / target^2 and
 *                     coef[n] = d mean_n(penalty) / d||g_n|| / ||g_n|| (the tangent seed is coef[n] * g_n). */
int tmx_mbstd_tangent(tmx_handle_t h, const float* x, const float* xdot, float* ydot, float* sdot, int N, int H, int W,
                      int C, int C_total, int group_size, tmx_stream_t s);
int tmx_mbstd_curvature(tmx_handle_t h, const float* x, const float* xdot, const float* lam, float* q, int N, int H, int W,
                        int C, int group_size, tmx_stream_t s);
int tmx_dense_wgrad(tmx_handle_t h, const float* x, const float* dy, const float* y, float* dw, float* db, int N, int K,
                    int Cout, float wscale, int lrelu, float alpha, tmx_stream_t s);
int tmx_scale_rows(tmx_handle_t h, const float* in, const float* scale, float* out, int rows, int64_t len, tmx_stream_t s);
int tmx_gp_coefficients(tmx_handle_t h, const float* sq_norms, float* penalty, float* coef, int N, float lambda,
                        float target, tmx_stream_t s);

/* Network.run output conversion (tfutil.py:649-659): y = saturate_cast(round(avg_pool_shrink(x * mul + add))) over
 * `planes` = N*C image planes [H][W] fp32.  out_kind 0: fp32, no rounding; 1: uint8 (round half to even, clamp to
 * [0,255], NaN -> 0); 2: fp32 holding rounded values (the caller narrows to the other integer types). */
int tmx_convert_output(tmx_handle_t h, const float* x, void* y, int64_t planes, int H, int W, float mul, float add,
                       int shrink, int out_kind, tmx_stream_t s);
/* App-level matte compositing (util_scripts.py:1262,1268 `np.sum(latents * weights, axis=0)`, :1337,1342
 * `left * matt + right * (1 - matt)`): out[n][c][p] = sum_k srcs[k][n][c][p] * weights[k][p] over K NCHW sources
 * (device array of K device pointers; bcast[k] != 0: source k is [N][C][1][1]).  float64 products and sums rounded
 * once to float32 (numpy promotion), or float32 throughout (math_f32). */
int tmx_weighted_sum(tmx_handle_t h, const float* const* srcs, const int* bcast, const double* weights, float* out, int K,
                     int N, int C, int H, int W, int math_f32, tmx_stream_t s);
/* out = tanh(in) over n fp32 elements: the generator's image head when lod != 0 (networks.py:482-483). */
int tmx_tanh_f32(tmx_handle_t h, const float* in, float* out, int64_t n, tmx_stream_t s);
/* its adjoint: dx = dy * (1 - y^2) with y = tanh(x). */
int tmx_tanh_bwd(tmx_handle_t h, const float* dy, const float* y, float* dx, int64_t n, tmx_stream_t s);
/* y = [lrelu](x + bias[c]) on NHWC fp32 [npix][C]: apply_bias + leaky_relu behind the fused conv2d_downscale2d
 * (networks.py:142-148), whose bias and activation follow the 2x2 average. */
int tmx_bias_act(tmx_handle_t h, const float* x, const float* bias, float* y, int64_t npix, int C, int lrelu,
                 float alpha, tmx_stream_t s);
/* pixel_norm (networks.py:170-172) on NHWC fp32 [npix][C]: y = x * rsqrt(mean_c x^2 + eps). */
int tmx_pixel_norm(tmx_handle_t h, const float* x, float* y, int64_t npix, int C, float eps, tmx_stream_t s);
/* its adjoint: dx = r * dy - r^3 / C * x * sum_c(dy * x), r = rsqrt(mean_c x^2 + eps). */
int tmx_pixel_norm_bwd(tmx_handle_t h, const float* x, const float* dy, float* dx, int64_t npix, int C, float eps,
                       tmx_stream_t s);

/* KL regulariser of EG_wgan (loss.py:163-171) on one encoder's (mu, log_sigma) pair, n elements in all:
 * val = 1 + 2 ls - mu^2 - exp(2 ls) (reduce with tmx_row_sum, scale -0.5 * kl_weight / n), and the gradient of the
 * batch mean, dmu = gscale * mu, dls = gscale * (exp(2 ls) - 1) with gscale = kl_weight / n. */
int tmx_kl_terms(tmx_handle_t h, const float* mu, const float* log_sigma, float* dmu, float* dls, float* val, int64_t n,
                 float gscale, tmx_stream_t s);

/* Window of a [A][H][W][B] fp32 tensor (NCHW: A = N*C, B = 1; NHWC: A = N, B = C) at an offset that may live in
 * DEVICE memory (off_dev = {oy, ox} int32, overrides oy / ox when not NULL), so that the launch is identical from
 * one train step to the next and the step can be replayed as a CUDA graph: random_crop (loss.py:78-90) and the
 * crop-aware latent windows derived from it.
 *   embed = 0: dst[A][wh][ww][B] = src[A][oy+y][ox+x][B];   embed = 1: the adjoint - dst[A][H][W][B] = the window
 *   src[A][wh][ww][B] placed at (oy, ox), zeros elsewhere (every element of dst is written). */
int tmx_window_copy(tmx_handle_t h, const float* src, float* dst, int64_t A, int H, int W, int B, int wh, int ww,
                    int oy, int ox, const int32_t* off_dev, int embed, tmx_stream_t s);

/* ------------------------------------------------------------------ VGG-19 Gram-matrix loss
 * custom_vgg19.py:31-40: out[n][y][x][0..2] = ((rgb + 1) / 2 * 255)[BGR] - (103.939, 116.779, 123.68), channels 3..15
 *   zero (NHWC, 16 channels = one tensor-core K chunk), from an NCHW image in [-1, 1]; _bwd: its adjoint.
 * tmx_gram_fwd   (loss.py:29-35): G[n][i][j] = sum_p F[n][i][p] F[n][j][p] / H / W, F = NCHW features [N][C][H*W].
 * tmx_gram_l1    (loss.py:68-75 multi_layer_diff + its gradient): per sample n, sums[n] += val_scale * sum |G[n] - T[n']|
 *                and S[n] = (accumulate ? S[n] : 0) + coef * sign(G[n] - T[n']), n' = N-1-n when reverse_t; both are
 *                further scaled by w = 1 / *wdev / 1 - *wdev (wmode 0 / 1 / 2; the blend weight of loss.py:253-254).
 * tmx_gram_bwd   dF[n][i][p] = sum_j (S[n][i][j] + S[n][j][i]) F[n][j][p] / H / W. */
int tmx_vgg_preprocess(tmx_handle_t h, const float* img_nchw, float* out_nhwc16, int N, int H, int W, tmx_stream_t s);
int tmx_vgg_preprocess_bwd(tmx_handle_t h, const float* dout_nhwc16, float* dimg_nchw, int N, int H, int W,
                           tmx_stream_t s);
int tmx_gram_fwd(tmx_handle_t h, const float* F, float* G, int N, int C, int H, int W, tmx_stream_t s);
/* The same Gram matrices on the tensor cores from the SPLIT_BF16_HALO feature planes [N][H+2][W+2][C] (any halo):
 * the weight-gradient kernel with the sample as the tap and F as both operands (bf16x3, fp32 accumulate, split-K
 * over pixels, fixed-order reduction).  *bytes == 0: shape not served (H*W % 64 != 0 or < 64) - use tmx_gram_fwd. */
int tmx_gram_fwd_tc_workspace_bytes(tmx_handle_t h, int N, int C, int H, int W, size_t* bytes);
int tmx_gram_fwd_tc(tmx_handle_t h, const uint16_t* f_hi, const uint16_t* f_lo, float* G, float* workspace, int N,
                    int C, int H, int W, tmx_stream_t s);
/* Per-sample weight planes of the Gram-loss gradient as a 1x1 conv (TMX_CONV_W_PER_SAMPLE):
 * w[n][i][j] = (S[n][i][j] + S[n][j][i]) * scale, split into bf16 hi / lo, [N][C][C]. */
int tmx_gram_sym_split(tmx_handle_t h, const float* S, uint16_t* w_hi, uint16_t* w_lo, int N, int C, float scale,
                       tmx_stream_t s);
int tmx_gram_l1(tmx_handle_t h, const float* G, const float* T, float* S, float* sums, int N, int C, int reverse_t,
                float coef, float val_scale, int accumulate, const float* wdev, int wmode, tmx_stream_t s);
int tmx_gram_bwd(tmx_handle_t h, const float* S, const float* F, float* dF, int N, int C, int H, int W, tmx_stream_t s);

/* out = a + b over n fp32 elements (two gradient contributions meeting at one tensor). */
int tmx_add_f32(tmx_handle_t h, const float* a, const float* b, float* out, int64_t n, tmx_stream_t s);

/* tmx_grad_prepare: gradient w.r.t. a layer OUTPUT y [N][H][W][C] -> operand of that layer's dgrad / wgrad:
 *   v = src (+ add) ; v *= (y > 0 ? 1 : alpha) if mask ; dbias[c] += dbias_scale * sum v ; write v.
 *   src_kind 0: g on the zero-ringed grid (a consumer's tmx_conv2d_dgrad output) folded by `fold`
 *               (0 REFLECT adjoint of networks.py:55, 1 REPLICATE adjoint, 2 none);
 *            1: plain NHWC fp32 [N][H][W][C];   2: NHWC fp32 at half resolution = downscale2d adjoint (x 0.25).
 *   mask_kind 0 none, 1: y as NHWC fp32, 2: y as the hi plane of SPLIT_BF16_HALO [N][H+2][W+2][C] (sign only).
 *   Outputs (any subset): dz_hi/dz_lo planes on the zero-ringed grid (ring written as zeros) - or, with
 *   phase_pack, on the HALF-resolution grid [N][H/2+4][W/2+4][4C] with channel (a*2+b)*C + c for pixel
 *   (2y+a, 2x+b) (caller zeroes that buffer's ring once); dz_f32 NHWC; dbias[C] (atomic accumulate). */
typedef struct {
  int32_t N, H, W, C;
  int32_t src_kind, fold, mask_kind, phase_pack;
  float alpha;
  float dbias_scale;
} tmx_grad_desc_t;

typedef struct {
  const float* g;
  const float* add;     /* NHWC fp32 [N][H][W][C] or NULL */
  const void* y_mask;
  uint16_t* dz_hi;
  uint16_t* dz_lo;
  float* dz_f32;
  float* dbias;
} tmx_grad_io_t;

int tmx_grad_prepare(tmx_handle_t h, const tmx_grad_desc_t* d, const tmx_grad_io_t* io, tmx_stream_t s);

/* tmx_conv2d_dgrad_gp: tmx_conv2d_dgrad (same first twelve arguments) FOLLOWED BY tmx_grad_prepare(src_kind 0) on its
 * output, as one tensor-core kernel plus a border pass: `gd` / `gio` describe the [N][H][W][Cin] INPUT activation of this
 * layer exactly as they would for tmx_grad_prepare (fold, mask of the layer that produced it, add, dz_hi / dz_lo
 * [required], dz_f32, dbias; gio->g is ignored - the gradient comes from the accumulators).  The epilogue of the data
 * gradient finishes every interior pixel that receives no folded ring value (add, mask, bias gradient, re-split into
 * the planes); g_f32 - still a full [N][H+4][W+4][Cin] scratch buffer - only receives the ring and the rows / columns
 * the REFLECT / REPLICATE adjoint folds onto, which a small second kernel finishes.  Planes and dz_f32 are bit-identical
 * to the two-call sequence; dbias is summed in another order.  Replaces what tf.gradients emits for
 * pad -> conv2d -> bias -> leaky_relu chains (networks.py:48-75; tfutil.py:299).
 * *served = 1: done.  *served = 0: nothing was launched (shape served better elsewhere: Cin, Cout <= 64 thin layers,
 * Cin not a multiple of 32, maps smaller than 4 x 4 with a fold, TMX_NO_FUSED_GP=1) - call the two functions. */
int tmx_conv2d_dgrad_gp(tmx_handle_t h, int N, int H, int W, int Cin, int Cout, int k, const uint16_t* dz_hi,
                        const uint16_t* dz_lo, const uint16_t* wt_hi, const uint16_t* wt_lo, float* g_f32,
                        const tmx_grad_desc_t* gd, const tmx_grad_io_t* gio, int* served, tmx_stream_t s);

/* ------------------------------------------------------------------ optimizer (tfutil.py:246-399, 611-621)
 * All over flat fp32 buffers (a network's variables are one contiguous, 256-B aligned buffer).
 * tmx_nonfinite_check: *flag = 1 if any g[i] is inf/nan (never clears it; tfutil.py:347-355).
 * tmx_adam_step: TF1 Adam on (w, g*grad_scale, m, v); lr_t = lr*sqrt(1-powers[1])/(1-powers[0]) with the running
 *   beta powers on the device (initialise to {beta1, beta2}); the whole step - including the power update - is
 *   skipped when skip_flag != NULL and *skip_flag != 0.
 * tmx_ema_update: dst = src + (dst - src) * beta  (Network.setup_as_moving_average_of). */
int tmx_nonfinite_check(tmx_handle_t h, const float* g, int64_t n, int* flag, tmx_stream_t s);
int tmx_adam_step(tmx_handle_t h, float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float grad_scale, float* powers, const int* skip_flag, tmx_stream_t s);
int tmx_ema_update(tmx_handle_t h, const float* src, float* dst, int64_t n, float beta, tmx_stream_t s);
/* Data-parallel form (replaces the per-variable nccl.all_sum + per-tower finite check of
 * tfutil.py:326-355): every rank marks its LOCAL non-finite gradients in a float slot that travels in the tail of
 * the gradient bucket (tmx_nonfinite_mark: *mark = 1 if any g[i] is inf/nan, never clears it); after the SUM
 * all-reduce the slot is > 0 on every rank iff any rank overflowed.
 * tmx_adam_update: the Adam step of tmx_adam_step WITHOUT the beta-power update, skipped when *skip_flag != 0 or
 *   *skip_mark != 0 (either may be NULL) - several networks of one optimizer step with the same powers;
 * tmx_adam_advance: the beta-power update, once per optimizer, under the same skip condition. */
int tmx_nonfinite_mark(tmx_handle_t h, const float* g, int64_t n, float* mark, tmx_stream_t s);
int tmx_adam_update(tmx_handle_t h, float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                    float beta2, float eps, float grad_scale, const float* powers, const int* skip_flag,
                    const float* skip_mark, tmx_stream_t s);
int tmx_adam_advance(tmx_handle_t h, float* powers, float beta1, float beta2, const int* skip_flag,
                     const float* skip_mark, tmx_stream_t s);

/* ------------------------------------------------------------------ permutation sampler (host)
 * run.py:107-182 (my_swap_h / my_swap_w / block_permutation) as driven by
 * run.py:436-507, in index-vector form: writes `count` int32 vectors of `length`
 * entries; vector r means P_h[i, r[i]] = 1 (row map) or P_w[r[j], j] = 1
 * (column map) - both obey the same recurrence.  `u` holds uniforms pre-drawn
 * from the caller's RNG in the reference's draw order (2*(length>>l) per level l
 * with length>>l > 1); *consumed returns how many were used.  Host only, no GPU. */
int tmx_perm_indices_from_uniforms(const double* u, int64_t n_u, int length, int levels, int count, int32_t* out,
                                   int64_t* consumed);

#ifdef __cplusplus
}
#endif
#endif /* TMX_H_ */
