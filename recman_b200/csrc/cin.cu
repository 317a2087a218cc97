// K5: CIN layer entry points (recman/tf/core/layers.py:711-751) - dispatch between the CUDA-core fp32 path
// (cin_simt.cu) and the tcgen05 tensor-core path (cin_tc.cu).
#include "cin.cuh"

namespace rm {

// split-half + sum-pool of one layer's output (layers.py:738-751): the first n0 feature maps feed the next layer (a
// view, nothing to do), the others are summed over D.  One thread per (b, n >= n0); fixed order -> deterministic.
__global__ void __launch_bounds__(256) cin_pool_fwd_kernel(const float* __restrict__ out, int64_t B, int N, int D, int n0,
                                                           float* __restrict__ pooled) {
  const int NP = N - n0;
  const int64_t total = B * NP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / NP;
    const int n = n0 + (int)(i - b * NP);
    const float* src = out + (b * N + n) * D;
    float acc = 0.f;
    if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      for (int d = 0; d < D; d += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + d);
        acc += v.x; acc += v.y; acc += v.z; acc += v.w;
      }
    } else {
      for (int d = 0; d < D; ++d) acc += src[d];
    }
    pooled[i] = acc;
  }
}

// its backward: dout[b, n, d] = n < n0 ? d_next[b, n, d] (0 when absent) : d_pool[b, n - n0]   - one write pass.
// V = 4: one float4 per thread, consecutive threads -> consecutive 16-byte chunks (D % 4 == 0, aligned); V = 1: scalar.
template <int V>
__global__ void __launch_bounds__(256) cin_pool_bwd_kernel(const float* __restrict__ d_next, int64_t next_bstride,
                                                           const float* __restrict__ d_pool, int64_t B, int N, int D,
                                                           int n0, float* __restrict__ dout) {
  const int NP = N - n0;
  const int DV = D / V;
  const int64_t total = B * N * DV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t bn;
    int c;
    if (total < ((int64_t)1 << 31)) {
      const uint32_t iu = (uint32_t)i;
      bn = iu / (uint32_t)DV;
      c = (int)(iu - (uint32_t)bn * (uint32_t)DV);
    } else {
      bn = i / DV;
      c = (int)(i - bn * DV);
    }
    const uint32_t b = (uint32_t)bn / (uint32_t)N;  // B * N < 2^31 is checked by the caller
    const int n = (int)((uint32_t)bn - b * (uint32_t)N);
    if (V == 4) {
      float4 v;
      if (n < n0) {
        v = d_next ? *reinterpret_cast<const float4*>(d_next + (int64_t)b * next_bstride + (int64_t)n * D + 4 * c)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const float s = d_pool ? d_pool[(int64_t)b * NP + (n - n0)] : 0.f;
        v = make_float4(s, s, s, s);
      }
      reinterpret_cast<float4*>(dout)[i] = v;
    } else {
      float v;
      if (n < n0) v = d_next ? d_next[(int64_t)b * next_bstride + (int64_t)n * D + c] : 0.f;
      else v = d_pool ? d_pool[(int64_t)b * NP + (n - n0)] : 0.f;
      dout[i] = v;
    }
  }
}

}  // namespace rm

extern "C" {

int rm_cin_pool_fwd(const float* out, int64_t B, int32_t N, int32_t D, int32_t n0, float* pooled, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && N > 0 && D > 0 && n0 >= 0 && n0 < N, "bad shape");
  if (B == 0) return 0;
  RM_CHECK_ARG(out && pooled, "null pointer");
  cin_pool_fwd_kernel<<<grid_for(B * (int64_t)(N - n0), 256, 8), 256, 0, (cudaStream_t)stream>>>(out, B, N, D, n0, pooled);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_cin_pool_bwd(const float* d_next, int64_t next_bstride, const float* d_pool, int64_t B, int32_t N, int32_t D,
                    int32_t n0, float* dout, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && N > 0 && D > 0 && n0 >= 0 && n0 < N, "bad shape");
  if (B == 0) return 0;
  RM_CHECK_ARG(dout, "null pointer");
  RM_CHECK_ARG(!d_next || next_bstride >= (int64_t)n0 * D, "batch stride too small");
  RM_UNSUPPORTED(B * (int64_t)N < ((int64_t)1 << 31), "B * N must be < 2^31");
  const bool vec = (D & 3) == 0 && aligned16(dout) && (!d_next || (aligned16(d_next) && (next_bstride & 3) == 0));
  if (vec)
    cin_pool_bwd_kernel<4><<<grid_for(B * (int64_t)N * (D / 4), 256, 8), 256, 0, (cudaStream_t)stream>>>(
        d_next, next_bstride, d_pool, B, N, D, n0, dout);
  else
    cin_pool_bwd_kernel<1><<<grid_for(B * (int64_t)N * D, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        d_next, next_bstride, d_pool, B, N, D, n0, dout);
  RM_LAUNCH_CHECK();
  return 0;
}

size_t rm_cin_layer_workspace_bytes(int64_t B, int32_t m, int32_t H, int32_t D, int32_t N, int32_t precision) {
  if (precision != RM_CIN_FP32_SIMT && rm::cin_tc_supported(B, m, H, D, N))
    return rm::cin_tc_fwd_workspace(B, m, H, D, N, precision);
  return 0;
}

int rm_cin_layer_fwd(const float* x0, int64_t x0_bstride, const float* xk, int64_t xk_bstride, const float* W,
                     const float* bias, int64_t B, int32_t m, int32_t H, int32_t D, int32_t N, int32_t act,
                     int32_t precision, float* out, float* pre, void* workspace, size_t workspace_bytes,
                     void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && m > 0 && H > 0 && D > 0 && N > 0, "bad shape");
  RM_CHECK_ARG(act >= RM_ACT_IDENTITY && act <= RM_ACT_LEAKY_RELU, "unknown activation");
  RM_CHECK_ARG(precision >= RM_CIN_FP32_SIMT && precision <= RM_CIN_TF32, "unknown precision");
  if (B == 0) return 0;
  RM_CHECK_ARG(x0 && xk && W && bias && out, "null pointer");
  RM_CHECK_ARG(x0_bstride >= (int64_t)m * D && xk_bstride >= (int64_t)H * D, "batch stride too small");
  RM_UNSUPPORTED(B <= ((int64_t)1 << 24), "batch must be <= 2^24");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != RM_CIN_FP32_SIMT && cin_tc_supported(B, m, H, D, N))
    return cin_fwd_tc(x0, x0_bstride, xk, xk_bstride, W, bias, B, m, H, D, N, act, precision, out, pre, workspace,
                      workspace_bytes, st);
  return cin_fwd_simt(x0, x0_bstride, xk, xk_bstride, W, bias, B, m, H, D, N, act, out, pre, st);
}

size_t rm_cin_layer_bwd_workspace_bytes(int64_t B, int32_t m, int32_t H, int32_t D, int32_t N, int32_t precision) {
  if (B < 0 || m <= 0 || H <= 0 || D <= 0 || N <= 0) return 0;
  if (precision != RM_CIN_FP32_SIMT && rm::cin_tc_bwd_supported(B, m, H, D, N))
    return rm::cin_tc_bwd_workspace(B, m, H, D, N, precision);
  return rm::cin_bwd_layout(B, m, H, D, N, nullptr).total;
}

int rm_cin_layer_bwd(const float* x0, int64_t x0_bstride, const float* xk, int64_t xk_bstride, const float* W,
                     const float* pre, const float* dout, int64_t B, int32_t m, int32_t H, int32_t D, int32_t N,
                     int32_t act, int32_t precision, float* dW, float* dbias, float* dx0, float* dxk,
                     int64_t dxk_bstride, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && m > 0 && H > 0 && D > 0 && N > 0, "bad shape");
  RM_CHECK_ARG(act >= RM_ACT_IDENTITY && act <= RM_ACT_LEAKY_RELU, "unknown activation");
  RM_CHECK_ARG(dW && dbias, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    RM_CUDA(cudaMemsetAsync(dW, 0, (size_t)m * H * N * sizeof(float), st));
    RM_CUDA(cudaMemsetAsync(dbias, 0, (size_t)N * sizeof(float), st));
    return 0;
  }
  RM_CHECK_ARG(x0 && xk && W && pre && dout && dx0 && dxk && workspace, "null pointer");
  RM_CHECK_ARG(x0_bstride >= (int64_t)m * D && xk_bstride >= (int64_t)H * D && dxk_bstride >= (int64_t)H * D,
               "batch stride too small");
  if (precision != RM_CIN_FP32_SIMT && cin_tc_bwd_supported(B, m, H, D, N) && x0_bstride % 4 == 0 &&
      xk_bstride % 4 == 0 && aligned16(x0) && aligned16(xk))
    return cin_bwd_tc(x0, x0_bstride, xk, xk_bstride, W, pre, dout, B, m, H, D, N, act, precision, dW, dbias, dx0, dxk,
                      dxk_bstride, workspace, workspace_bytes, st);
  return cin_bwd_simt(x0, x0_bstride, xk, xk_bstride, W, pre, dout, B, m, H, D, N, act, dW, dbias, dx0, dxk,
                      dxk_bstride, workspace, workspace_bytes, st);
}

}  // extern "C"
