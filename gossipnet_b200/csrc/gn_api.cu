// Error string, version, device queries for the C ABI (include/gossipnet_b200.h).
#include <stdarg.h>

#include "gn_common.cuh"

namespace gn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  return n;
}

}  // namespace gn

extern "C" {

const char* gn_last_error(void) { return gn::g_err; }
int gn_abi_version(void) { return 1; }
int gn_sm_count(void) { return gn::sm_count(); }

}  // extern "C"
