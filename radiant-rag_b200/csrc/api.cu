// Library bookkeeping for librr_b200.so: device selection, error strings.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace rr {

static thread_local char g_err[512] = "";
static int g_sm_count = 0;
static int g_smem_optin = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() { return g_sm_count; }
int max_smem_optin() { return g_smem_optin; }

}  // namespace rr

extern "C" int rr_abi_version(void) { return RR_ABI_VERSION; }

extern "C" const char* rr_last_error(void) { return rr::g_err; }

extern "C" int rr_sm_count(void) { return rr::g_sm_count; }

extern "C" int rr_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    rr::set_error("rr_init: no CUDA device (%s)", cudaGetErrorString(e));
    return RR_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    rr::set_error("rr_init: device %d out of range (%d devices)", device, count);
    return RR_ERR_INVALID;
  }
  RR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    rr::set_error("rr_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
                  prop.major, prop.minor);
    return RR_ERR_NO_DEVICE;
  }
  rr::g_sm_count = prop.multiProcessorCount;
  rr::g_smem_optin = (int)prop.sharedMemPerBlockOptin;
  return RR_OK;
}
