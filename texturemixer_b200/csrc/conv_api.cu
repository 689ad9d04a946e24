// conv_api.cu — tmx_conv2d_fwd: argument checking (the reference's asserts,
// networks.py:49) and kernel selection by shape.  There is exactly one
// hardware target (sm_100a); "algo" picks between the CUDA-core exact-fp32
// kernel and the tcgen05 bf16x3 kernel, it is not a backend switch.
#include "common.cuh"

int tmx_conv2d_fwd_ffma(tmx_handle_t h, const tmx_conv_desc_t* d, const tmx_conv_io_t* io, cudaStream_t st);
int tmx_conv2d_fwd_tc(tmx_handle_t h, const tmx_conv_desc_t* d, const tmx_conv_io_t* io, int kc_max, cudaStream_t st);

extern "C" int tmx_conv2d_fwd(tmx_handle_t h, const tmx_conv_desc_t* d, const tmx_conv_io_t* io, tmx_stream_t s) {
  TMX_REQUIRE(h && d && io, TMX_ERR_ARG, "tmx_conv2d_fwd: NULL argument");
  TMX_REQUIRE(d->k == 1 || d->k == 3, TMX_ERR_SHAPE, "tmx_conv2d_fwd: kernel=%d (only 1 and 3 occur on the path)", d->k);
  TMX_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, TMX_ERR_SHAPE,
              "tmx_conv2d_fwd: bad shape N=%d H=%d W=%d Cin=%d Cout=%d", d->N, d->H, d->W, d->Cin, d->Cout);
  TMX_REQUIRE(d->k == 1 || (d->H >= 2 && d->W >= 2), TMX_ERR_SHAPE,
              "tmx_conv2d_fwd: REFLECT pad needs H, W >= 2 (got %d x %d)", d->H, d->W);
  const void* ptrs[] = {io->x_f32, io->x_hi, io->x_lo, io->w, io->w_hi, io->w_lo, io->bias, io->residual,
                        io->y_f32, io->y_hi, io->y_lo, io->y_rgb};
  for (const void* q : ptrs)
    TMX_REQUIRE(((uintptr_t)q & 15) == 0, TMX_ERR_ARG,
                "tmx_conv2d_fwd: every buffer must be 16-byte aligned (got %p) - the kernels use 128-bit accesses", q);
  cudaStream_t st = (cudaStream_t)s;
  int algo = d->algo;
  if (algo == TMX_ALGO_AUTO) algo = (io->x_hi != nullptr) ? TMX_ALGO_TC : TMX_ALGO_FFMA;
  switch (algo) {
    case TMX_ALGO_FFMA: return tmx_conv2d_fwd_ffma(h, d, io, st);
    case TMX_ALGO_TC: return tmx_conv2d_fwd_tc(h, d, io, 64, st);
    case TMX_ALGO_TC_K32: return tmx_conv2d_fwd_tc(h, d, io, 32, st);
    default: return tmx_fail(TMX_ERR_ARG, "tmx_conv2d_fwd: unknown algo %d", d->algo);
  }
}
