// Error plumbing of the C ABI: return codes + thread-local message (include/s2c.h).
#include <stdarg.h>

#include "s2c_common.cuh"

namespace s2c {
static thread_local char g_err[512] = {0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return S2C_ERR_CUDA;
}
}  // namespace s2c

extern "C" int s2c_version(void) { return 100; }
extern "C" const char *s2c_last_error(void) { return s2c::g_err; }
