// Fused DeepFM "head": everything between the first DNN layer's pre-activation and the loss, forward AND backward,
// in one kernel (+ one deterministic partial-sum reduction):
//
//   h1 = act(y1); y2 = h1 @ W2 + b2; h2 = act(y2); dnn = h2 @ w3 + b3        DNN.__call__ layers.py:589-609
//   logit = lin + w0 + fm + dnn                                             add_n of the towers, DeepFM.py:150-160
//   pred = sigmoid(logit)  (classification)                                 PredictionLayer layers.py:796-808
//   loss = mean( -(y log(p+eps) + (1-y) log(1-p+eps)) ), p = clip(pred, eps, 1-eps), eps = 1e-7
//                                                                            create_loss utils.py:192-198 (Keras BCE)
//          or mean((logit - y)^2) for regression
//   and the gradients TensorFlow's autodiff would produce: g = dL/dlogit (= the FM and first-order gradients),
//   g1 = dL/dy1, dW2, db2, dw3, db3, dw0, db1 - and, given the samples' dense features, the two products of the first
//   layer / first-order term that involve them: dW1_dense = dense^T g1 [n_dense, 32], dlin_dense = dense^T g [n_dense]
//   (what would otherwise be a reduce kernel, a skinny GEMM + split-K reduce and a GEMV per step).
//
// In the stock path these are ~100 tiny elementwise / reduce / GEMM launches on [B, 32] tensors - a third of a C5
// step after the embedding kernels were fused.  One thread owns one sample (its two 32-wide hidden vectors live in
// registers, W2 is broadcast from shared memory); batch reductions go CTA partial -> fixed-order final sum (no atomics:
// run-to-run identical).  Built for the reference's default tower hidden_units = (32, 32); other shapes keep the
// separate kernels.
#include "common.cuh"

namespace rm {

constexpr int HD_N = 32;          // both hidden widths
constexpr int HD_THREADS = 128;   // samples per CTA
constexpr int HD_PART = HD_N * HD_N + 3 * HD_N + 4;  // dW2 | db2 | dw3 | db1 | loss, g_sum, pad, pad

struct HeadParams {
  const float* y1;
  const float* fm;
  const float* lin;
  const float* w0;
  const float* W2;
  const float* b2;
  const float* w3;
  const float* b3;
  const float* labels;  // nullptr: forward only
  float* logit;
  float* pred;
  float* g1;
  float* g;
  float* partials;  // [grid, part_stride]: HD_PART + n_dense * (HD_N + 1) (dW1_dense rows | dlin_dense), padded
  const float* dense;  // [B, nd] or nullptr
  int64_t B;
  int act, task, nd, part_stride;
  float inv_B;
};

template <int ACT>
__device__ __forceinline__ float hd_actT(float v) {
  if (ACT == RM_ACT_RELU) return v > 0.f ? v : 0.f;
  if (ACT == RM_ACT_LEAKY_RELU) return v > 0.f ? v : 0.2f * v;
  return v;
}
// act'(y) read off h = act(y): the three activations keep the sign of their argument
template <int ACT>
__device__ __forceinline__ float hd_dactT(float h) {
  if (ACT == RM_ACT_RELU) return h > 0.f ? 1.f : 0.f;
  if (ACT == RM_ACT_LEAKY_RELU) return h > 0.f ? 1.f : 0.2f;
  return 1.f;
}

template <int ACT>
__global__ void __launch_bounds__(HD_THREADS) head_kernel(const HeadParams P) {
  __shared__ __align__(16) float sW2[HD_N * HD_N];
  __shared__ float sb2[HD_N], sw3[HD_N];
  __shared__ float sA[HD_THREADS][HD_N + 1];  // h1 of the CTA's samples, later g1
  __shared__ float sD[HD_THREADS][HD_N + 1];  // dy2 of the CTA's samples
  __shared__ float sred[HD_THREADS / 32][3 * HD_N + 2];  // per-warp column sums: db2 | dw3 | db1 | loss | g
  const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
  for (int i = tid; i < HD_N * HD_N; i += HD_THREADS) sW2[i] = P.W2[i];
  if (tid < HD_N) {
    sb2[tid] = P.b2[tid];
    sw3[tid] = P.w3[tid];
  }
  const int64_t b = (int64_t)blockIdx.x * HD_THREADS + tid;
  const bool live = b < P.B;
  // h1 = act(y1) goes to this thread's row of sA; the i-loops below stay rolled (h1[i] is read back from shared memory),
  // only the 32 outputs of a layer live in registers
#pragma unroll
  for (int i = 0; i < HD_N; i += 4) {
    const float4 v = live ? ld4(P.y1 + b * HD_N + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    sA[tid][i] = hd_actT<ACT>(v.x);
    sA[tid][i + 1] = hd_actT<ACT>(v.y);
    sA[tid][i + 2] = hd_actT<ACT>(v.z);
    sA[tid][i + 3] = hd_actT<ACT>(v.w);
  }
  __syncthreads();
  float h2[HD_N];
#pragma unroll
  for (int j = 0; j < HD_N; ++j) h2[j] = sb2[j];
#pragma unroll 2
  for (int i = 0; i < HD_N; ++i) {
    const float a = sA[tid][i];
    const float4* wrow = reinterpret_cast<const float4*>(sW2 + i * HD_N);
#pragma unroll
    for (int j4 = 0; j4 < HD_N / 4; ++j4) {
      const float4 w = wrow[j4];
      h2[4 * j4] = fmaf(a, w.x, h2[4 * j4]);
      h2[4 * j4 + 1] = fmaf(a, w.y, h2[4 * j4 + 1]);
      h2[4 * j4 + 2] = fmaf(a, w.z, h2[4 * j4 + 2]);
      h2[4 * j4 + 3] = fmaf(a, w.w, h2[4 * j4 + 3]);
    }
  }
  float dnn = P.b3[0];
#pragma unroll
  for (int j = 0; j < HD_N; ++j) {
    h2[j] = hd_actT<ACT>(h2[j]);
    dnn = fmaf(h2[j], sw3[j], dnn);
  }
  float logit = 0.f, pred = 0.f;
  if (live) {
    logit = ((P.lin[b] + P.w0[0]) + P.fm[b]) + dnn;
    pred = P.task == 0 ? 1.f / (1.f + expf(-logit)) : logit;
    if (P.logit) P.logit[b] = logit;
    if (P.pred) P.pred[b] = pred;
  }
  if (!P.labels) return;  // forward only (uniform: kernel argument)

  // ---- loss and dL/dlogit
  float loss_b = 0.f, g = 0.f;
  if (live) {
    const float y = P.labels[b];
    if (P.task == 0) {
      const float eps = 1e-7f;
      const float p = fminf(fmaxf(pred, eps), 1.f - eps);
      loss_b = -(y * logf(p + eps) + (1.f - y) * logf(1.f - p + eps));
      float dp = 0.f;
      if (pred >= eps && pred <= 1.f - eps) dp = -(y / (p + eps) - (1.f - y) / (1.f - p + eps));
      g = dp * pred * (1.f - pred) * P.inv_B;
    } else {
      const float d = logit - y;
      loss_b = d * d;
      g = 2.f * d * P.inv_B;
    }
    P.g[b] = g;
  }
  sD[tid][HD_N] = g;  // the pad column keeps dL/dlogit of the CTA's samples for the dense-feature products below
  float* part = P.partials + (int64_t)blockIdx.x * P.part_stride;
  // ---- second layer backward: dy2 (kept in h2's registers), db2 / dw3 column sums by warp shuffle
#pragma unroll
  for (int j = 0; j < HD_N; ++j) {
    const float gh = g * h2[j];  // dw3 term
    const float d = g * sw3[j] * hd_dactT<ACT>(h2[j]);
    h2[j] = d;
    sD[tid][j] = d;
    const float s2 = warp_sum(d), s3 = warp_sum(gh);
    if (lane == 0) {
      sred[wv][j] = s2;
      sred[wv][HD_N + j] = s3;
    }
  }
  {
    const float ls = warp_sum(loss_b), gs = warp_sum(g);
    if (lane == 0) {
      sred[wv][3 * HD_N] = ls;
      sred[wv][3 * HD_N + 1] = gs;
    }
  }
  __syncthreads();
  {
    // dW2[i][j] = sum_b h1[b][i] * dy2[b][j]: thread t owns i = t / 4, j = 8 (t % 4) .. +7, samples in order
    const int i = tid >> 2, j0 = (tid & 3) * 8;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int s = 0; s < HD_THREADS; ++s) {
      const float a = sA[s][i];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(a, sD[s][j0 + e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[i * HD_N + j0 + e] = acc[e];
  }
  __syncthreads();  // every thread is done with the other rows of sA: row tid may now be overwritten with g1
  // ---- first layer backward: g1[i] = act'(y1[i]) * sum_j dy2[j] * W2[i][j]
#pragma unroll 2
  for (int i = 0; i < HD_N; ++i) {
    const float4* wrow = reinterpret_cast<const float4*>(sW2 + i * HD_N);
    float acc = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < HD_N / 4; ++j4) {
      const float4 w = wrow[j4];
      acc = fmaf(h2[4 * j4], w.x, acc);
      acc = fmaf(h2[4 * j4 + 1], w.y, acc);
      acc = fmaf(h2[4 * j4 + 2], w.z, acc);
      acc = fmaf(h2[4 * j4 + 3], w.w, acc);
    }
    const float g1 = acc * hd_dactT<ACT>(sA[tid][i]);
    sA[tid][i] = g1;
    const float s1 = warp_sum(g1);
    if (lane == 0) sred[wv][2 * HD_N + i] = s1;
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < HD_N; i += 4)
      st4(P.g1 + b * HD_N + i, make_float4(sA[tid][i], sA[tid][i + 1], sA[tid][i + 2], sA[tid][i + 3]));
  }
  __syncthreads();
  if (tid < 3 * HD_N + 2) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < HD_THREADS / 32; ++w) acc += sred[w][tid];
    part[HD_N * HD_N + tid] = acc;  // db2 | dw3 | db1 | loss | g sum
  }
  if (tid == 0) {
    part[HD_N * HD_N + 3 * HD_N + 2] = 0.f;
    part[HD_N * HD_N + 3 * HD_N + 3] = 0.f;
  }
  if (P.dense) {
    // dW1_dense[j][n] = sum_s dense[s][j] * g1[s][n],  dlin_dense[j] = sum_s dense[s][j] * g[s]: four interleaved
    // partial sums over the CTA's samples, added in a fixed order
    const int64_t b0 = (int64_t)blockIdx.x * HD_THREADS;
    const int nlive = (int)(P.B - b0 < HD_THREADS ? P.B - b0 : HD_THREADS);
    const float* dn = P.dense + b0 * P.nd;
    const bool staged = P.nd <= HD_N;  // the CTA's dense rows fit the (dead) dy2 columns of sD
    __syncthreads();
    if (staged) {
      for (int idx = tid; idx < HD_THREADS * P.nd; idx += HD_THREADS) {
        const int sidx = idx / P.nd, j = idx - sidx * P.nd;
        sD[sidx][j] = sidx < nlive ? __ldg(dn + idx) : 0.f;
      }
      __syncthreads();
    }
    for (int idx = tid; idx < P.nd * (HD_N + 1); idx += HD_THREADS) {
      const int j = idx / (HD_N + 1), n = idx - j * (HD_N + 1);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (staged) {
#pragma unroll 2
        for (int sidx = 0; sidx < HD_THREADS; sidx += 4) {
          a0 = fmaf(sD[sidx][j], n < HD_N ? sA[sidx][n] : sD[sidx][HD_N], a0);
          a1 = fmaf(sD[sidx + 1][j], n < HD_N ? sA[sidx + 1][n] : sD[sidx + 1][HD_N], a1);
          a2 = fmaf(sD[sidx + 2][j], n < HD_N ? sA[sidx + 2][n] : sD[sidx + 2][HD_N], a2);
          a3 = fmaf(sD[sidx + 3][j], n < HD_N ? sA[sidx + 3][n] : sD[sidx + 3][HD_N], a3);
        }
      } else {
        for (int sidx = 0; sidx < nlive; ++sidx)
          a0 = fmaf(__ldg(dn + (int64_t)sidx * P.nd + j), n < HD_N ? sA[sidx][n] : sD[sidx][HD_N], a0);
      }
      part[HD_PART + idx] = (a0 + a1) + (a2 + a3);
    }
  }
}

// out[e] = sum over the CTAs' partials, fixed association (run-to-run identical): CTA = 32 columns x 8 row groups,
// group r adds rows r, r + 8, ... in order, then the 8 group sums are added in order.
__global__ void __launch_bounds__(256) head_reduce_kernel(const float* __restrict__ partials, int n_part, int stride,
                                                          int nd, float inv_B, float* __restrict__ dW2,
                                                          float* __restrict__ db2, float* __restrict__ dw3,
                                                          float* __restrict__ db1, float* __restrict__ loss,
                                                          float* __restrict__ dscal, float* __restrict__ dW1_dense,
                                                          float* __restrict__ dlin_dense) {
  __shared__ float sm[8][33];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + c;
  float acc = 0.f;
  if (e < stride)
    for (int p = r; p < n_part; p += 8) acc += partials[(int64_t)p * stride + e];
  sm[r][c] = acc;
  __syncthreads();
  if (r != 0 || e >= stride) return;
#pragma unroll
  for (int q = 1; q < 8; ++q) acc += sm[q][c];
  if (e < HD_N * HD_N) dW2[e] = acc;
  else if (e < HD_N * HD_N + HD_N) db2[e - HD_N * HD_N] = acc;
  else if (e < HD_N * HD_N + 2 * HD_N) dw3[e - HD_N * HD_N - HD_N] = acc;
  else if (e < HD_N * HD_N + 3 * HD_N) db1[e - HD_N * HD_N - 2 * HD_N] = acc;
  else if (e == HD_N * HD_N + 3 * HD_N) loss[0] = acc * inv_B;
  else if (e == HD_N * HD_N + 3 * HD_N + 1) dscal[0] = acc;  // db3 = dw0 = sum_b g_b
  else if (e >= HD_PART && e < HD_PART + nd * (HD_N + 1)) {
    const int idx = e - HD_PART, j = idx / (HD_N + 1), n = idx - j * (HD_N + 1);
    if (n < HD_N) dW1_dense[j * HD_N + n] = acc;
    else dlin_dense[j] = acc;
  }
}

static int head_part_stride(int nd) { return (HD_PART + nd * (HD_N + 1) + 3) & ~3; }

}  // namespace rm

extern "C" {

int rm_deepfm_head_supported(int32_t N1, int32_t N2) { return (N1 == rm::HD_N && N2 == rm::HD_N) ? 1 : 0; }

size_t rm_deepfm_head_workspace_bytes(int64_t B, int32_t n_dense) {
  if (B <= 0 || n_dense < 0) return 256;
  return (size_t)rm::ceil_div(B, rm::HD_THREADS) * rm::head_part_stride(n_dense) * sizeof(float);
}

int rm_deepfm_head(const float* y1, const float* fm, const float* lin, const float* w0, const float* W2,
                   const float* b2, const float* w3, const float* b3, const float* labels, const float* dense,
                   int32_t n_dense, int64_t B, int32_t N1, int32_t N2, int32_t act, int32_t task, float grad_scale,
                   float* logit, float* pred, float* loss, float* g1, float* g, float* dW2, float* db2, float* dw3,
                   float* dscal, float* db1, float* dW1_dense, float* dlin_dense, void* workspace,
                   size_t workspace_bytes, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(y1 && fm && lin && w0 && W2 && b2 && w3 && b3, "null pointer");
  RM_CHECK_ARG(B >= 0 && (task == 0 || task == 1) && n_dense >= 0, "bad argument");
  RM_UNSUPPORTED(rm_deepfm_head_supported(N1, N2), "the fused head is built for hidden_units = (32, 32)");
  RM_UNSUPPORTED(act == RM_ACT_IDENTITY || act == RM_ACT_RELU || act == RM_ACT_LEAKY_RELU, "activation kind");
  RM_UNSUPPORTED(aligned16(y1) && (!g1 || aligned16(g1)), "rows must be 16-byte aligned");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool with_dense = labels && dense && n_dense > 0;
  HeadParams P;
  P.y1 = y1; P.fm = fm; P.lin = lin; P.w0 = w0; P.W2 = W2; P.b2 = b2; P.w3 = w3; P.b3 = b3; P.labels = labels;
  P.logit = logit; P.pred = pred; P.g1 = g1; P.g = g; P.partials = (float*)workspace; P.B = B; P.act = act; P.task = task;
  P.dense = with_dense ? dense : nullptr;
  P.nd = with_dense ? n_dense : 0;
  P.part_stride = head_part_stride(P.nd);
  P.inv_B = grad_scale / (float)B;
  const int grid = (int)ceil_div(B, HD_THREADS);
  if (labels) {
    RM_CHECK_ARG(loss && g1 && g && dW2 && db2 && dw3 && dscal && db1 && workspace, "null output pointer");
    RM_CHECK_ARG(!with_dense || (dW1_dense && dlin_dense), "null output pointer (dense-feature gradients)");
    const size_t need = (size_t)grid * P.part_stride * sizeof(float);
    if (workspace_bytes < need) {
      set_error("rm_deepfm_head: workspace %zu < required %zu", workspace_bytes, need);
      return RM_E_WORKSPACE;
    }
  }
  if (act == RM_ACT_RELU) head_kernel<RM_ACT_RELU><<<grid, HD_THREADS, 0, st>>>(P);
  else if (act == RM_ACT_LEAKY_RELU) head_kernel<RM_ACT_LEAKY_RELU><<<grid, HD_THREADS, 0, st>>>(P);
  else head_kernel<RM_ACT_IDENTITY><<<grid, HD_THREADS, 0, st>>>(P);
  RM_LAUNCH_CHECK();
  if (labels) {
    head_reduce_kernel<<<(int)ceil_div(P.part_stride, 32), 256, 0, st>>>((const float*)workspace, grid, P.part_stride,
                                                                         P.nd, 1.f / (float)B, dW2, db2, dw3, db1, loss,
                                                                         dscal, dW1_dense, dlin_dense);
    RM_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
