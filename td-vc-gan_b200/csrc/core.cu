#include <cstdlib>
// Error reporting, device queries.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace tdvc {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
double g_flops[FLOP_FAMILIES] = {0, 0, 0, 0, 0, 0, 0, 0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TDVC_PDL");
    on = e ? (atoi(e) != 0) : 1;
  }
  return on != 0;
}
}  // namespace tdvc

extern "C" const char* tdvc_last_error(void) { return tdvc::g_err; }
extern "C" int tdvc_version(void) { return 100; }
extern "C" int64_t tdvc_launch_count(void) { return (int64_t)tdvc::g_launches; }
extern "C" double tdvc_flop_count(int family) {
  return (family >= 0 && family < tdvc::FLOP_FAMILIES) ? tdvc::g_flops[family] : -1.0;
}
extern "C" int tdvc_device_is_sm100(void) {
  int dev = 0, major = 0;
  TDVC_CUDA(cudaGetDevice(&dev));
  TDVC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}
