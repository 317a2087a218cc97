// K4: DCN cross network, all L layers fused, forward and backward.
//
// Call site: recman/tf/core/DCN.py:135-137 (`CrossNet(cross_layer_num, l2)(dnn_input)`;
// the class is absent from the reference).  Arithmetic: arXiv 1708.05123 eq. 3,
//   x_{l+1} = x0 * (x_l . w_l) + b_l + x_l ,   logit = x_L . w_out + w0_out.
//
// Roofline: HBM.  Forward reads x once (4d B/sample) and writes 4 + 4L B; the
// unfused TF form would move 3*4*d*L B.  Backward reads x twice and writes dx
// (3*4*d B/sample).  One warp owns S (1 or 2) samples: x0 and x_l live in registers
// (DPL values per lane), the x_l . w_l dots are warp-shuffle reductions.
//
// Backward uses the closed form x_l = x0*(1 + sum_{j<l} s_j) + sum_{j<l} b_j with
// the saved dots s_j, so no activations are stored.  Parameter gradients are
// batch reductions: dw_l = sum_b x0[b]*coef[b,l] + cb_l*T_l etc. (derivation in
// DESIGN.md); they are reduced per fixed sample chunk, then in chunk order -
// deterministic, no atomics.
#include "common.cuh"

namespace rm {

constexpr int CROSS_MAXL = 15;  // L + 1 <= 16 accumulators in the reduction kernel

template <int DPL, int S>
__global__ void __launch_bounds__(256) cross_fwd_kernel(const float* __restrict__ x, int64_t ld,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ w_out,
                                                        const float* __restrict__ w0_out, int64_t B, int d, int L,
                                                        float* __restrict__ logit, float* __restrict__ dots) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float w0 = w0_out ? w0_out[0] : 0.f;
  for (int64_t b0 = warp * S; b0 < B; b0 += n_warps * S) {
    float x0[S][DPL], xl[S][DPL];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const bool live = b0 + s < B;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        const int c = lane + 32 * i;
        x0[s][i] = (live && c < d) ? x[(b0 + s) * ld + c] : 0.f;
        xl[s][i] = x0[s][i];
      }
    }
    for (int l = 0; l < L; ++l) {
      float dot[S];
#pragma unroll
      for (int s = 0; s < S; ++s) dot[s] = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        const int c = lane + 32 * i;
        const float wv = c < d ? __ldg(w + (int64_t)l * d + c) : 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) dot[s] += xl[s][i] * wv;
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        dot[s] = warp_sum(dot[s]);
        if (lane == 0 && b0 + s < B && dots) dots[(b0 + s) * L + l] = dot[s];
      }
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        const int c = lane + 32 * i;
        const float bv = c < d ? __ldg(bias + (int64_t)l * d + c) : 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) xl[s][i] = (x0[s][i] * dot[s] + bv) + xl[s][i];
      }
    }
    float acc[S];
#pragma unroll
    for (int s = 0; s < S; ++s) acc[s] = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const int c = lane + 32 * i;
      const float wv = c < d ? __ldg(w_out + c) : 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s] += xl[s][i] * wv;
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float r = warp_sum(acc[s]);
      if (lane == 0 && b0 + s < B) logit[b0 + s] = r + w0;
    }
  }
}

// Per-sample backward: writes dx and the per-sample coefficients
//   coef[b, l] = (1 + cs_l) * t_l  (l < L),  coef[b, L] = g * (1 + cs_L)
//   tg  [b, l] = t_l               (l < L),  tg  [b, L] = g
template <int DPL>
__global__ void __launch_bounds__(256) cross_bwd_sample_kernel(const float* __restrict__ x, int64_t ld,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ w_out,
                                                               const float* __restrict__ dots,
                                                               const float* __restrict__ gout, int64_t B, int d, int L,
                                                               float* __restrict__ dx, int64_t d_ld, int accumulate,
                                                               float* __restrict__ coef, float* __restrict__ tg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int L1 = L + 1;
  for (int64_t b = warp; b < B; b += n_warps) {
    float x0[DPL], delta[DPL], dx0[DPL];
    const float g = gout[b];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const int c = lane + 32 * i;
      x0[i] = c < d ? x[b * ld + c] : 0.f;
      delta[i] = c < d ? g * __ldg(w_out + c) : 0.f;
      dx0[i] = 0.f;
    }
    float cs = 0.f;  // sum of all dots
    for (int l = 0; l < L; ++l) cs += dots[b * L + l];
    if (lane == 0) {
      coef[b * L1 + L] = g * (1.f + cs);
      tg[b * L1 + L] = g;
    }
    for (int l = L - 1; l >= 0; --l) {
      const float sl = dots[b * L + l];
      cs -= sl;  // now cs = sum_{j<l} s_j
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) t += x0[i] * delta[i];
      t = warp_sum(t);
      if (lane == 0) {
        coef[b * L1 + l] = (1.f + cs) * t;
        tg[b * L1 + l] = t;
      }
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        const int c = lane + 32 * i;
        dx0[i] += delta[i] * sl;
        delta[i] += (c < d ? __ldg(w + (int64_t)l * d + c) : 0.f) * t;
      }
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const int c = lane + 32 * i;
      if (c < d) {
        const float r = dx0[i] + delta[i];
        float* o = dx + b * d_ld + c;
        *o = accumulate ? *o + r : r;
      }
    }
  }
}

// Chunked batch reduction: block `blk` owns samples [blk*chunk, (blk+1)*chunk).
//   P[blk, l, c] = sum_b coef[b,l] * x[b,c]      (l <= L)
//   Tp[blk, l]   = sum_b tg[b,l]
__global__ void __launch_bounds__(256) cross_bwd_reduce_kernel(const float* __restrict__ x, int64_t ld,
                                                               const float* __restrict__ coef,
                                                               const float* __restrict__ tg, int64_t B, int d, int L,
                                                               int64_t chunk, float* __restrict__ P,
                                                               float* __restrict__ Tp) {
  const int L1 = L + 1;
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = lo + chunk < B ? lo + chunk : B;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc[CROSS_MAXL + 1];
#pragma unroll
    for (int l = 0; l <= CROSS_MAXL; ++l) acc[l] = 0.f;
    for (int64_t b = lo; b < hi; ++b) {
      const float xv = x[b * ld + c];
#pragma unroll
      for (int l = 0; l <= CROSS_MAXL; ++l)
        if (l < L1) acc[l] += coef[b * L1 + l] * xv;
    }
#pragma unroll
    for (int l = 0; l <= CROSS_MAXL; ++l)
      if (l < L1) P[((int64_t)blockIdx.x * L1 + l) * d + c] = acc[l];
  }
  if (threadIdx.x < L1) {
    float acc = 0.f;
    for (int64_t b = lo; b < hi; ++b) acc += tg[b * L1 + threadIdx.x];
    Tp[(int64_t)blockIdx.x * L1 + threadIdx.x] = acc;
  }
}

// Ordered final pass over the chunk partials + closed-form terms: one thread per (layer l, column c), so the nblk-long
// ordered sums run in parallel over (L+1)*d threads instead of serially per column.
__global__ void __launch_bounds__(256) cross_bwd_final_kernel(const float* __restrict__ P, const float* __restrict__ Tp,
                                                              int nblk, const float* __restrict__ w,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ w_out, int d, int L,
                                                              float* __restrict__ dw, float* __restrict__ db,
                                                              float* __restrict__ dw_out, float* __restrict__ dw0_out) {
  const int L1 = L + 1;
  __shared__ float T[CROSS_MAXL + 1];
  if (threadIdx.x < L1) {
    float acc = 0.f;
    for (int k = 0; k < nblk; ++k) acc += Tp[(int64_t)k * L1 + threadIdx.x];
    T[threadIdx.x] = acc;
  }
  __syncthreads();
  const float G = T[L];
  const int l = blockIdx.y;
  if (blockIdx.x == 0 && l == 0 && threadIdx.x == 0 && dw0_out) dw0_out[0] = G;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float p = 0.f;
  for (int k = 0; k < nblk; ++k) p += P[((int64_t)k * L1 + l) * d + c];
  float cb = 0.f;  // cb_l = sum_{j<l} b_j
  for (int j = 0; j < l; ++j) cb += __ldg(bias + (int64_t)j * d + c);
  if (l < L) {
    if (dw) dw[(int64_t)l * d + c] = p + cb * T[l];
    if (db) {  // db_l = w_out*G + sum_{j>l} w_j*T_j, summed from j = L-1 downwards
      float tail = __ldg(w_out + c) * G;
      for (int j = L - 1; j > l; --j) tail += __ldg(w + (int64_t)j * d + c) * T[j];
      db[(int64_t)l * d + c] = tail;
    }
  } else if (dw_out) {
    dw_out[c] = p + cb * G;
  }
}

struct CrossWs {
  float* coef;
  float* tg;
  float* P;
  float* Tp;
  int nblk;
  int64_t chunk;
  size_t total;
};

static CrossWs cross_layout(int64_t B, int d, int L, void* base) {
  CrossWs w;
  const int L1 = L + 1;
  int64_t nblk = ceil_div(B, 16);  // small chunks: the chunk kernel is a serial walk per column, the final pass is parallel
  if (nblk > 2 * RM_NUM_SMS) nblk = 2 * RM_NUM_SMS;
  if (nblk < 1) nblk = 1;
  w.chunk = ceil_div(B, nblk);
  if (w.chunk < 1) w.chunk = 1;
  w.nblk = (int)ceil_div(B > 0 ? B : 1, w.chunk);
  char* p = (char*)base;
  size_t off = 0;
  w.coef = (float*)(p + off); off += align_up((size_t)B * L1 * 4, 256);
  w.tg = (float*)(p + off); off += align_up((size_t)B * L1 * 4, 256);
  w.P = (float*)(p + off); off += align_up((size_t)w.nblk * L1 * d * 4, 256);
  w.Tp = (float*)(p + off); off += align_up((size_t)w.nblk * L1 * 4, 256);
  w.total = off;
  return w;
}

static int pick_dpl(int d) {
  const int need = (d + 31) / 32;
  if (need <= 4) return 4;
  if (need <= 8) return 8;
  if (need <= 16) return 16;
  if (need <= 32) return 32;
  if (need <= 64) return 64;
  return 0;
}

}  // namespace rm

extern "C" {

int rm_cross_fwd(const float* x, int64_t ld, const float* w, const float* b, const float* w_out, const float* w0_out,
                 int64_t B, int32_t d, int32_t L, float* logit, float* dots, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(x && w_out && logit && (L == 0 || (w && b)), "null pointer");
  RM_CHECK_ARG(B >= 0 && d > 0 && L >= 0 && ld >= d, "bad shape");
  const int dpl = pick_dpl(d);
  RM_UNSUPPORTED(dpl != 0, "d must be <= 2048");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define RM_CF(D, S)                                                                                          \
  case D:                                                                                                    \
    cross_fwd_kernel<D, S><<<grid_for(B, 8 * S, 4), 256, 0, st>>>(x, ld, w, b, w_out, w0_out, B, d, L, logit, \
                                                                  dots);                                     \
    break
  switch (dpl) {
    RM_CF(4, 2);
    RM_CF(8, 2);
    RM_CF(16, 2);
    RM_CF(32, 1);
    RM_CF(64, 1);
  }
#undef RM_CF
  RM_LAUNCH_CHECK();
  return 0;
}

size_t rm_cross_bwd_workspace_bytes(int64_t B, int32_t d, int32_t L) {
  if (B < 0 || d <= 0 || L < 0) return 0;
  return rm::cross_layout(B, d, L, nullptr).total;
}

int rm_cross_bwd(const float* x, int64_t ld, const float* w, const float* b, const float* w_out, const float* dots,
                 const float* gout, int64_t B, int32_t d, int32_t L, float* dx, int64_t d_ld, int32_t accumulate,
                 float* dw, float* db, float* dw_out, float* dw0_out, void* workspace, size_t workspace_bytes,
                 void* stream) {
  using namespace rm;
  RM_CHECK_ARG(x && w_out && dots && gout && dx && workspace && (L == 0 || (w && b)), "null pointer");
  RM_CHECK_ARG(B >= 0 && d > 0 && L >= 0 && ld >= d && d_ld >= d, "bad shape");
  RM_UNSUPPORTED(L <= CROSS_MAXL, "at most 15 cross layers");
  const int dpl = pick_dpl(d);
  RM_UNSUPPORTED(dpl != 0, "d must be <= 2048");
  CrossWs ws = cross_layout(B, d, L, workspace);
  if (workspace_bytes < ws.total) {
    set_error("rm_cross_bwd: workspace %zu < required %zu", workspace_bytes, ws.total);
    return RM_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (B > 0) {
    const int grid = grid_for(B, 8, 4);
#define RM_CB(D)                                                                                                  \
  case D:                                                                                                         \
    cross_bwd_sample_kernel<D><<<grid, 256, 0, st>>>(x, ld, w, w_out, dots, gout, B, d, L, dx, d_ld, accumulate,  \
                                                     ws.coef, ws.tg);                                             \
    break
    switch (dpl) {
      RM_CB(4);
      RM_CB(8);
      RM_CB(16);
      RM_CB(32);
      RM_CB(64);
    }
#undef RM_CB
    RM_LAUNCH_CHECK();
    cross_bwd_reduce_kernel<<<ws.nblk, 256, 0, st>>>(x, ld, ws.coef, ws.tg, B, d, L, ws.chunk, ws.P, ws.Tp);
    RM_LAUNCH_CHECK();
  }
  const int nblk = B > 0 ? ws.nblk : 0;
  cross_bwd_final_kernel<<<dim3((unsigned)ceil_div(d, 256), (unsigned)(L + 1)), 256, 0, st>>>(
      ws.P, ws.Tp, nblk, w, b, w_out, d, L, dw, db, dw_out, dw0_out);
  RM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
