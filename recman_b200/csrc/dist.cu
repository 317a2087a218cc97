// (e) Row-sharded embedding tables: staging kernels either side of the NCCL all-to-all.
//
// Rank i routes every (sample, field) id to its owner (row r of a table lives on rank r mod W at local row r div W),
// sorted by destination.  Owners gather straight into the return buffer (rm_gather_fwd with out_stride = KP), so
// the only extra passes are these two:
//   rm_unpack_rows     received rows (routed order)  -> DNN input row buffer x + per-position bias / linear values
//   rm_pack_grad_rows  d(x) + FM backward            -> gradient rows in routed order (the a2a send buffer); the
//                      owner's K2 (rm_segment_plan/reduce) consumes the receive buffer directly.
// A routed row is KP = k + 4 floats: [e_0..e_{k-1} | bias | lin | 0 | 0] (16-byte aligned rows).
#include "common.cuh"

namespace rm {

template <int LPR>
__global__ void __launch_bounds__(256) unpack_rows_kernel(const float* __restrict__ recv, int64_t n, int KP,
                                                          const int32_t* __restrict__ pos, uint32_t m, int k,
                                                          float* __restrict__ x, int64_t ld, float* __restrict__ bias_out,
                                                          float* __restrict__ lin_out) {
  const int lir = threadIdx.x % LPR;
  const int k4 = k >> 2;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  for (int64_t j = group; j < n; j += n_groups) {
    const uint32_t p = (uint32_t)pos[j];
    const uint32_t b = p / m, f = p - b * m;
    const float* src = recv + j * KP;
    float* dst = x + (int64_t)b * ld + (int64_t)f * k;
    for (int c = lir; c < k4; c += LPR) st4(dst + 4 * c, ld4(src + 4 * c));
    if (lir == 0) {
      if (bias_out) bias_out[p] = src[k];
      if (lin_out) lin_out[p] = src[k + 1];
    }
  }
}

template <int LPR>
__global__ void __launch_bounds__(256) pack_grad_rows_kernel(const float* __restrict__ dx, const float* __restrict__ x,
                                                             int64_t ld, const float* __restrict__ sum,
                                                             const float* __restrict__ g_fm,
                                                             const float* __restrict__ g_lin, int64_t n, int KP,
                                                             const int32_t* __restrict__ pos, uint32_t m, int k,
                                                             float* __restrict__ send) {
  const int lir = threadIdx.x % LPR;
  const int k4 = k >> 2;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  for (int64_t j = group; j < n; j += n_groups) {
    const uint32_t p = pos ? (uint32_t)pos[j] : (uint32_t)j;  // pos == NULL: rows stay in position order
    const uint32_t b = p / m, f = p - b * m;
    const int64_t o = (int64_t)b * ld + (int64_t)f * k;
    const float gf = g_fm ? g_fm[b] : 0.f;
    float* dst = send + j * KP;
    for (int c = lir; c < k4; c += LPR) {
      float4 v = dx ? ld4(dx + o + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (g_fm) {
        const float4 xv = ld4(x + o + 4 * c);
        const float4 sv = ld4(sum + (int64_t)b * k + 4 * c);
        v.x += gf * (sv.x - xv.x);
        v.y += gf * (sv.y - xv.y);
        v.z += gf * (sv.z - xv.z);
        v.w += gf * (sv.w - xv.w);
      }
      st4(dst + 4 * c, v);
    }
    if (lir == 0 && KP >= k + 4) st4(dst + k, make_float4(gf, g_lin ? g_lin[b] : 0.f, 0.f, 0.f));
  }
}

static inline int lpr_for(int k) {
  int p = 1;
  while (p < k / 4) p <<= 1;
  return p > 32 ? 32 : p;
}

}  // namespace rm

extern "C" {

int rm_unpack_rows(const float* recv, int64_t n, int32_t KP, const int32_t* pos, int32_t m, int32_t k, float* x,
                   int64_t ld, float* bias_out, float* lin_out, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(n >= 0 && m > 0 && k > 0 && KP >= k + 2, "bad shape");
  if (n == 0) return 0;
  RM_CHECK_ARG(recv && pos && x, "null pointer");
  RM_UNSUPPORTED(k % 4 == 0 && KP % 4 == 0 && ld % 4 == 0 && aligned16(recv) && aligned16(x),
                 "routed rows need k % 4 == 0 and 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
#define RM_UP(L)                                                                                                     \
  case L:                                                                                                            \
    unpack_rows_kernel<L><<<grid_for(n, 256 / L, 8), 256, 0, st>>>(recv, n, KP, pos, (uint32_t)m, k, x, ld, bias_out, \
                                                                   lin_out);                                         \
    break
  switch (lpr_for(k)) {
    RM_UP(1);
    RM_UP(2);
    RM_UP(4);
    RM_UP(8);
    RM_UP(16);
    RM_UP(32);
  }
#undef RM_UP
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_pack_grad_rows(const float* dx, const float* x, int64_t ld, const float* sum, const float* g_fm,
                      const float* g_lin, int64_t n, int32_t KP, const int32_t* pos, int32_t m, int32_t k, float* send,
                      void* stream) {
  using namespace rm;
  RM_CHECK_ARG(n >= 0 && m > 0 && k > 0 && (KP == k || KP >= k + 4), "bad shape");
  if (n == 0) return 0;
  RM_CHECK_ARG(send, "null pointer");
  RM_CHECK_ARG(!g_fm || (x && sum), "g_fm needs x and sum");
  RM_UNSUPPORTED(k % 4 == 0 && KP % 4 == 0 && ld % 4 == 0 && (!dx || aligned16(dx)) && (!x || aligned16(x)) &&
                     (!sum || aligned16(sum)) && aligned16(send),
                 "routed rows need k % 4 == 0 and 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
#define RM_PK(L)                                                                                                       \
  case L:                                                                                                              \
    pack_grad_rows_kernel<L><<<grid_for(n, 256 / L, 8), 256, 0, st>>>(dx, x, ld, sum, g_fm, g_lin, n, KP, pos,          \
                                                                      (uint32_t)m, k, send);                           \
    break
  switch (lpr_for(k)) {
    RM_PK(1);
    RM_PK(2);
    RM_PK(4);
    RM_PK(8);
    RM_PK(16);
    RM_PK(32);
  }
#undef RM_PK
  RM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
