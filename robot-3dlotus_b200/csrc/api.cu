// Library-level entry points: version, error channel.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void lotus_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

LOTUS_API const char* lotus_last_error(void) { return g_err; }
LOTUS_API int lotus_version(void) { return 100; }  // 0.1.0
LOTUS_API int lotus_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}
