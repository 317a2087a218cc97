// N1: optimizer step on K2's (unique rows, summed gradient rows) output and on
// dense parameters.  The reference builds a NEW optimizer inside every
// fit_on_batch (recman/tf/core/xDeepFM.py:116-126 -> create_optimizer,
// recman/tf/core/utils.py:201-213), so every step is a first step with zero
// slot variables; these kernels implement exactly that stateless update
// (oracle.fresh_optimizer_step) and touch only the rows that received gradient.
#include "optim.cuh"

namespace rm {

__global__ void __launch_bounds__(256) sparse_opt_kernel(float* __restrict__ table, int k, int64_t row_stride,
                                                         const int64_t* __restrict__ uniq_rows,
                                                         const float* __restrict__ rows,
                                                         const int32_t* __restrict__ n_unique, OptParams o) {
  const int64_t total = (int64_t)(*n_unique) * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = i / k;
    const int c = (int)(i - u * k);
    float* p = table + uniq_rows[u] * row_stride + c;
    *p = opt_update(*p, rows[i], o);
  }
}

__global__ void __launch_bounds__(256) sparse_opt_vec_kernel(float* __restrict__ table, int k4,
                                                             const int64_t* __restrict__ uniq_rows,
                                                             const float* __restrict__ rows,
                                                             const int32_t* __restrict__ n_unique, OptParams o) {
  const int64_t total = (int64_t)(*n_unique) * k4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = i / k4;
    const int c = (int)(i - u * k4);
    float* p = table + (uniq_rows[u] * (int64_t)k4 + c) * 4;
    float4 pv = ld4(p);
    const float4 g = ld4(rows + i * 4);
    pv.x = opt_update(pv.x, g.x, o);
    pv.y = opt_update(pv.y, g.y, o);
    pv.z = opt_update(pv.z, g.z, o);
    pv.w = opt_update(pv.w, g.w, o);
    st4(p, pv);
  }
}

__global__ void __launch_bounds__(256) dense_opt_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n,
                                                        OptParams o) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = opt_update(p[i], g[i], o);
}

// Multi-tensor variant: the replicated (dense) parameters are many small tensors - one launch updates up to
// DENSE_BATCH of them (descriptors travel as kernel parameters: no device table, CUDA-graph capturable).
constexpr int DENSE_BATCH = 96;
struct DenseBatch {
  float* p[DENSE_BATCH];
  const float* g[DENSE_BATCH];
  uint32_t n[DENSE_BATCH];
};

__global__ void __launch_bounds__(256) dense_opt_multi_kernel(const DenseBatch batch, OptParams o) {
  const int t = blockIdx.y;
  float* __restrict__ p = batch.p[t];
  const float* __restrict__ g = batch.g[t];
  const uint32_t n = batch.n[t];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = opt_update(p[i], g[i], o);
}

}  // namespace rm

extern "C" {

int rm_sparse_opt_step_strided(float* table, int32_t k, int64_t row_stride, const int64_t* uniq_rows, const float* rows,
                               const int32_t* n_unique, int64_t max_rows, int32_t opt, float lr, float l2,
                               void* stream) {
  using namespace rm;
  RM_CHECK_ARG(table && uniq_rows && rows && n_unique, "null pointer");
  RM_CHECK_ARG(k > 0 && max_rows >= 0 && row_stride >= k, "bad shape");
  OptParams o;
  int rc = make_params(opt, lr, l2, &o);
  if (rc) return rc;
  if (max_rows == 0) return 0;
  sparse_opt_kernel<<<grid_for(max_rows * k, 256, 8), 256, 0, (cudaStream_t)stream>>>(table, k, row_stride, uniq_rows,
                                                                                    rows, n_unique, o);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_sparse_opt_step(float* table, int32_t k, const int64_t* uniq_rows, const float* rows, const int32_t* n_unique,
                       int64_t max_rows, int32_t opt, float lr, float l2, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(table && uniq_rows && rows && n_unique, "null pointer");
  RM_CHECK_ARG(k > 0 && max_rows >= 0, "bad shape");
  OptParams o;
  int rc = make_params(opt, lr, l2, &o);
  if (rc) return rc;
  if (max_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (k % 4 == 0 && aligned16(table) && aligned16(rows)) {
    sparse_opt_vec_kernel<<<grid_for(max_rows * (k / 4), 256, 8), 256, 0, st>>>(table, k / 4, uniq_rows, rows, n_unique,
                                                                               o);
  } else {
    sparse_opt_kernel<<<grid_for(max_rows * k, 256, 8), 256, 0, st>>>(table, k, (int64_t)k, uniq_rows, rows, n_unique, o);
  }
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_dense_opt_step(float* p, const float* g, int64_t n, int32_t opt, float lr, float l2, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(p && g, "null pointer");
  RM_CHECK_ARG(n >= 0, "bad shape");
  OptParams o;
  int rc = make_params(opt, lr, l2, &o);
  if (rc) return rc;
  if (n == 0) return 0;
  dense_opt_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, n, o);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_dense_opt_step_multi(float* const* ps, const float* const* gs, const int64_t* ns, int32_t count, int32_t opt,
                            float lr, float l2, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(count >= 0 && (count == 0 || (ps && gs && ns)), "bad argument");
  OptParams o;
  int rc = make_params(opt, lr, l2, &o);
  if (rc) return rc;
  for (int32_t i = 0; i < count;) {  // consume tensors until a batch is full; empty tensors take no slot
    DenseBatch b;
    int nb = 0;
    int64_t max_n = 0;
    for (; i < count && nb < DENSE_BATCH; ++i) {
      RM_CHECK_ARG(ns[i] >= 0 && ns[i] < ((int64_t)1 << 32), "tensor too large for the multi-tensor update");
      if (ns[i] == 0) continue;
      RM_CHECK_ARG(ps[i] && gs[i], "null pointer");
      b.p[nb] = ps[i];
      b.g[nb] = gs[i];
      b.n[nb] = (uint32_t)ns[i];
      if (ns[i] > max_n) max_n = ns[i];
      ++nb;
    }
    if (nb == 0) continue;
    int gx = (int)ceil_div(max_n, 256 * 4);
    if (gx > 4 * RM_NUM_SMS) gx = 4 * RM_NUM_SMS;
    dense_opt_multi_kernel<<<dim3((unsigned)gx, (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(b, o);
    RM_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
