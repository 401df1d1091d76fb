// Library-wide plumbing of the C-ABI: error strings, version, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace avssl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace avssl

extern "C" int avssl_abi_version(void) { return 2; }

extern "C" const char* avssl_last_error(void) { return avssl::g_err; }

extern "C" int avssl_device_sm_count(void) {
  int n = avssl::sm_count();
  if (n < 0) {
    cudaGetLastError();
    avssl::set_error("no CUDA device available (this library has no CPU fallback)");
  }
  return n;
}
