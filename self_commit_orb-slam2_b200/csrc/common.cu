// common.cu — error text, device selection, version for libb200slam.so
#include <stdarg.h>

#include "common.cuh"

namespace b2s {
thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int select_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s): libb200slam has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return B2S_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (have %d)", device, n);
    return B2S_ERR_BAD_ARG;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return B2S_ERR_CUDA;
  }
  return B2S_OK;
}
}  // namespace b2s

extern "C" const char* b2s_last_error(void) { return b2s::g_last_error.c_str(); }
extern "C" int b2s_version(void) { return 100; }
extern "C" int b2s_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
