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
  return 0;
}

}  // extern "C"
