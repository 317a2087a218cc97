// K5: CIN layer entry points (recman/tf/core/layers.py:711-751) - dispatch between the CUDA-core fp32 path
// (cin_simt.cu) and the tcgen05 tensor-core path (cin_tc.cu).
#include "cin.cuh"

extern "C" {

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
