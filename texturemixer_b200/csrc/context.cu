// context.cu — handle lifetime, error reporting (include/tmx.h conventions).
#include <cstring>
#include <new>

#include "common.cuh"

static thread_local char g_err[512] = "";

int tmx_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int tmx_cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

extern "C" int tmx_abi_version(void) { return TMX_ABI_VERSION; }
extern "C" const char* tmx_last_error(void) { return g_err; }

extern "C" int tmx_create(int device, tmx_handle_t* out) {
  TMX_REQUIRE(out != nullptr, TMX_ERR_ARG, "tmx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  TMX_CUDA(cudaGetDeviceCount(&count));
  TMX_REQUIRE(device >= 0 && device < count, TMX_ERR_ARG, "tmx_create: device %d out of range (%d devices)", device,
              count);
  cudaDeviceProp prop;
  TMX_CUDA(cudaGetDeviceProperties(&prop, device));
  TMX_REQUIRE(prop.major == 10, TMX_ERR_ARCH,
              "tmx_create: device %d is sm_%d%d; libtmx is built for sm_100a (B200) only - there is no fallback",
              device, prop.major, prop.minor);
  TMX_CUDA(cudaSetDevice(device));
  tmx_ctx* c = new (std::nothrow) tmx_ctx();
  TMX_REQUIRE(c != nullptr, TMX_ERR_ARG, "tmx_create: out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->launches = 0;
  c->encode_tiled = nullptr;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete c;
    return tmx_fail(TMX_ERR_DRIVER, "tmx_create: cuTensorMapEncodeTiled not available from the driver");
  }
  c->encode_tiled = (tmx_encode_tiled_fn)fn;
  *out = c;
  return TMX_OK;
}

extern "C" int tmx_destroy(tmx_handle_t h) {
  TMX_REQUIRE(h != nullptr, TMX_ERR_ARG, "tmx_destroy: NULL handle");
  delete h;
  return TMX_OK;
}

extern "C" int tmx_device_info(tmx_handle_t h, int* sm_count, int* cc_major, int* cc_minor) {
  TMX_REQUIRE(h != nullptr, TMX_ERR_ARG, "tmx_device_info: NULL handle");
  if (sm_count) *sm_count = h->sm_count;
  if (cc_major) *cc_major = h->cc_major;
  if (cc_minor) *cc_minor = h->cc_minor;
  return TMX_OK;
}

extern "C" int tmx_launch_count(tmx_handle_t h, uint64_t* count) {
  TMX_REQUIRE(h != nullptr && count != nullptr, TMX_ERR_ARG, "tmx_launch_count: NULL argument");
  *count = h->launches;
  return TMX_OK;
}
