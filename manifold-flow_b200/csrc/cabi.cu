// Error reporting / version entry points of the C ABI (include/mflow_b200.h).
#include <stdarg.h>
#include <string.h>

#include "mflow_common.cuh"

namespace mflow {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace mflow

extern "C" MFLOW_API int mflow_version(void) { return MFLOW_ABI_VERSION; }

extern "C" MFLOW_API const char* mflow_last_error_string(void) { return mflow::g_error; }

extern "C" MFLOW_API int mflow_check_device(void) {
  int dev = -1;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    mflow::set_error("no CUDA device is current");
    return MFLOW_E_NODEVICE;
  }
  if (prop.major != 10) {
    mflow::set_error("libmflow_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return MFLOW_E_NODEVICE;
  }
  return MFLOW_OK;
}
