// Internal interface between cin.cu (C ABI + dispatch), cin_simt.cu and cin_tc.cu.
#pragma once
#include "common.cuh"

namespace rm {

struct CinBwdWs {
  float* dF;       // [B,N,D] dout * act'(pre)
  float* partial;  // [slabs, m*H, N] per-slab dW partials
  int slabs;
  int slab_samples;
  size_t total;
};

// cin_simt.cu
int cin_fwd_simt(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* bias,
                 int64_t B, int m, int H, int D, int N, int act, float* out, float* pre, cudaStream_t st);
CinBwdWs cin_bwd_layout(int64_t B, int m, int H, int D, int N, void* base);
int cin_bwd_simt(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* pre,
                 const float* dout, int64_t B, int m, int H, int D, int N, int act, float* dW, float* dbias,
                 float* dx0, float* dxk, int64_t dbsk, void* workspace, size_t workspace_bytes, cudaStream_t st);

// dF = dout * act'(pre) and dbias[n] = sum_{b,d} dF (deterministic); shared by the SIMT and tensor-core backward
// (scratch: >= N floats of workspace that nothing else uses until this returns, e.g. the dW partial buffer)
int cin_dF_dbias(const float* dout, const float* pre, int64_t B, int N, int D, int act, float* dF, float* dbias,
                 float* scratch, size_t scratch_floats, cudaStream_t st);

// cin_tc_bwd.cu
bool cin_tc_bwd_supported(int64_t B, int m, int H, int D, int N);
size_t cin_tc_bwd_workspace(int64_t B, int m, int H, int D, int N, int precision);
int cin_bwd_tc(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* pre,
               const float* dout, int64_t B, int m, int H, int D, int N, int act, int precision, float* dW,
               float* dbias, float* dx0, float* dxk, int64_t dbsk, void* workspace, size_t workspace_bytes,
               cudaStream_t st);

// cin_tc.cu
bool cin_tc_supported(int64_t B, int m, int H, int D, int N);
size_t cin_tc_fwd_workspace(int64_t B, int m, int H, int D, int N, int precision);
int cin_fwd_tc(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* bias,
               int64_t B, int m, int H, int D, int N, int act, int precision, float* out, float* pre, void* workspace,
               size_t workspace_bytes, cudaStream_t st);

}  // namespace rm
