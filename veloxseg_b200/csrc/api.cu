// Error reporting and version entry points of libveloxseg_sm100.
#include "vx_kernels.h"

#include <stdarg.h>
#include <stdio.h>

namespace vx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VX_ERR_LAUNCH;
  }
  return VX_OK;
}

}  // namespace vx

extern "C" int vx_version(void) { return 100; }
extern "C" const char* vx_last_error_string(void) { return vx::g_err; }
