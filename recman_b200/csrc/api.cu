// Error plumbing, version and device check for librecman_b200.so.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <string.h>

#include "common.cuh"

namespace rm {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int tune_variant(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rm

extern "C" {

int rm_version(void) { return RM_ABI_VERSION; }

const char* rm_last_error(void) { return rm::g_err; }

int64_t rm_launch_count(void) { return (int64_t)rm::g_launches.load(std::memory_order_relaxed); }

int rm_device_check(int device) {
  cudaDeviceProp prop;
  RM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    rm::set_error("rm_device_check: device %d is sm_%d%d; librecman_b200 is built for sm_100a only", device,
                  prop.major, prop.minor);
    return RM_E_ARCH;
  }
  // tuning run only: L2 sector-promotion granularity (32 / 64 / 128 bytes) for the k=1 random lookups
  const int l2g = rm::tune_variant("RM_TUNE_L2_FETCH", 0);
  if (l2g > 0) RM_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)l2g));
  return 0;
}

}  // extern "C"
