// Library-level entry points: version, error string, device probe.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace nr

extern "C" int nr_version(void) { return NR_ABI_VERSION; }
extern "C" const char* nr_last_error(void) { return nr::g_err; }
extern "C" int nr_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}
