// Shared device/host helpers for librecman_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/recman_b200.h"

#define RM_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

namespace rm {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define RM_CHECK_ARG(cond, msg)                                       \
  do {                                                                \
    if (!(cond)) {                                                    \
      rm::set_error("%s: %s (%s)", __func__, msg, #cond);             \
      return RM_E_INVALID;                                            \
    }                                                                 \
  } while (0)

#define RM_UNSUPPORTED(cond, msg)                                     \
  do {                                                                \
    if (!(cond)) {                                                    \
      rm::set_error("%s: unsupported: %s (%s)", __func__, msg, #cond); \
      return RM_E_UNSUPPORTED;                                        \
    }                                                                 \
  } while (0)

#define RM_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      rm::set_error("%s: %s -> %s", __func__, #expr, cudaGetErrorString(e__));          \
      return (int)e__;                                                                  \
    }                                                                                   \
  } while (0)

#define RM_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    rm::count_launch();                                                                 \
    if (e__ != cudaSuccess) {                                                           \
      rm::set_error("%s: kernel launch -> %s", __func__, cudaGetErrorString(e__));      \
      return (int)e__;                                                                  \
    }                                                                                   \
  } while (0)

// opt a kernel in to `bytes` of dynamic shared memory: once per call site (and template instantiation), raised only
// when a larger size is requested - a launch-time constant, not a per-call driver round trip
#define RM_SMEM_ATTR_ONCE(bytes, ...)                                                                       \
  do {                                                                                                      \
    static int smem_attr_set__ = 0;                                                                         \
    if (smem_attr_set__ < (int)(bytes)) {                                                                   \
      RM_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      smem_attr_set__ = (int)(bytes);                                                                       \
    }                                                                                                       \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Row-sharded tables: global row `id` of a table lives on rank id mod W at local row id div W.  wshift >= 0: W is the
// power of two 1 << wshift (mask / shift); wshift < 0: any other W (one integer division per id).
inline int world_shift(int W) {
  int s = 0;
  while ((1 << s) < W) ++s;
  return (1 << s) == W ? s : -1;
}
__host__ __device__ __forceinline__ void shard_of(int64_t id, int W, int wshift, int& owner, int64_t& local) {
  if (wshift >= 0) {
    owner = (int)(id & (((int64_t)1 << wshift) - 1));
    local = id >> wshift;
  } else {
    const uint64_t q = (uint64_t)id / (uint32_t)W;
    owner = (int)((uint64_t)id - q * (uint32_t)W);
    local = (int64_t)q;
  }
}
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Grid for a grid-stride kernel: enough CTAs for the work, capped at a multiple of
// the SM count so that every wave is full.
inline int grid_for(int64_t work_items, int items_per_cta, int ctas_per_sm) {
  int64_t need = ceil_div(work_items, items_per_cta);
  int64_t cap = (int64_t)RM_NUM_SMS * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

#ifdef __CUDACC__

// 128-bit read-only streaming load (table rows are touched once per step: do not
// allocate in L1).
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over an aligned group of W lanes (W power of two <= 32); every lane gets the result
template <int W>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__

}  // namespace rm
