// Backward of a narrow dense layer  y[B,N] = x[B,d] W[d,N] + b  with N <= 64 (DeepFM's reference default is
// deep_hidden_units = (32, 32), recman/tf/core/DeepFM.py:35; DNN layer recman/tf/core/layers.py:576-609):
//
//   rm_linear_bwd_input   dx[B, d_ld] = g[B,N] W^T          (K = N is tiny: one pass, W^T and the g tile live in smem)
//   rm_linear_bwd_weight  dW[d, N]    = x^T g               (reduction over the batch: fixed slabs, ordered final sum)
//
// Both are "tall-skinny" products in which the batch is the only large dimension: the input-gradient is bound by
// writing dx (4*d_ld B per sample), the weight-gradient by reading x once; a library SGEMM tiles them for square
// problems (K = 32 leaves its pipeline empty, the 65536-long reduction gets split-K with atomics or a serial tail).
// Exact fp32 FMA arithmetic on the CUDA cores (no TF32): the MLP stays inside the 1e-5 parity bound.  The forward
// product and wide layers (N > 64) stay on cuBLAS.
#include "common.cuh"

namespace rm {

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// dx tile 128 rows x 128 columns per step, 8 x 8 outputs per thread, reduction over n < N from shared memory.
// Columns d <= c < d_ld (the 16-byte row padding of the row buffer) come out as exact zeros (their W rows are zero).
// ---------------------------------------------------------------------------------------------------------------
constexpr int DX_BM = 128, DX_BN = 128, DX_LD = 132;  // smem row stride (floats): 16-byte aligned, conflict-free

// A CTA keeps its g tile and walks DX_CT column tiles; the W^T slice of the next column tile is fetched into registers
// while the current one is multiplied (the transposing shared-memory store cannot be done by cp.async).
constexpr int DX_CT = 4;  // column tiles per CTA

template <int NT>
__global__ void __launch_bounds__(256, 2) linear_dx_kernel(const float* __restrict__ g, int64_t B, int N,
                                                           const float* __restrict__ W, int d, float* __restrict__ dx,
                                                           int64_t d_ld, const float* __restrict__ fx, int64_t fx_ld,
                                                           const float* __restrict__ fS, const float* __restrict__ fg,
                                                           int fk) {
  constexpr int LDV = (NT / 4) * 128 / 256;  // float4 loads per thread and tile
  extern __shared__ float smem_dx[];
  float* Gs = smem_dx;                // [NT][DX_LD]: Gs[n][r] = g[r0 + r, n]
  float* Ws0 = smem_dx + NT * DX_LD;  // [2][NT][DX_LD]: Ws[n][c] = W[c0 + c, n]
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * DX_BM;
  const int n_ct = (int)((d_ld + DX_BN - 1) / DX_BN);
  const int ct_lo = blockIdx.y * DX_CT;
  const int ct_hi = ct_lo + DX_CT < n_ct ? ct_lo + DX_CT : n_ct;
  if (ct_lo >= ct_hi) return;

  auto load_w = [&](int ct, float4 (&wv)[LDV]) {
#pragma unroll
    for (int i = 0; i < LDV; ++i) {
      const int idx = tid + i * 256;
      const int r = idx & 127, n4 = idx >> 7;
      const int c = ct * DX_BN + r;
      wv[i] = (4 * n4 < N && c < d) ? ld4(W + (int64_t)c * N + 4 * n4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_t = [&](float* dst, const float4 (&v)[LDV]) {  // transposed store: dst[n][r]
#pragma unroll
    for (int i = 0; i < LDV; ++i) {
      const int idx = tid + i * 256;
      const int r = idx & 127, n4 = idx >> 7;
      dst[(4 * n4 + 0) * DX_LD + r] = v[i].x; dst[(4 * n4 + 1) * DX_LD + r] = v[i].y;
      dst[(4 * n4 + 2) * DX_LD + r] = v[i].z; dst[(4 * n4 + 3) * DX_LD + r] = v[i].w;
    }
  };
  {
    float4 gv[LDV], wv[LDV];
#pragma unroll
    for (int i = 0; i < LDV; ++i) {
      const int idx = tid + i * 256;
      const int r = idx & 127, n4 = idx >> 7;
      gv[i] = (4 * n4 < N && r0 + r < B) ? ld4(g + (r0 + r) * N + 4 * n4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    load_w(ct_lo, wv);
    store_t(Gs, gv);
    store_t(Ws0, wv);
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  for (int ct = ct_lo; ct < ct_hi; ++ct) {
    const float* Ws = Ws0 + ((ct - ct_lo) & 1) * NT * DX_LD;
    float4 wnext[LDV];
    const bool more = ct + 1 < ct_hi;
    if (more) load_w(ct + 1, wnext);
    if (fx) {  // the epilogue's x values: pull them into L1 while the products are formed (no registers held)
      const int c0p = ct * DX_BN;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = r0 + ty * 8 + i;
        if (row < B) {
          if (c0p + tx * 4 < d) prefetch_l1(fx + row * fx_ld + c0p + tx * 4);
          if (c0p + 64 + tx * 4 < d) prefetch_l1(fx + row * fx_ld + c0p + 64 + tx * 4);
        }
      }
    }
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
      const float4 a0 = ld4(Gs + n * DX_LD + ty * 8), a1 = ld4(Gs + n * DX_LD + ty * 8 + 4);
      const float4 b0 = ld4(Ws + n * DX_LD + tx * 4), b1 = ld4(Ws + n * DX_LD + 64 + tx * 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    const int c0 = ct * DX_BN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = r0 + ty * 8 + i;
      if (row < B) {
        float* o = dx + row * d_ld;
        const int ca = c0 + tx * 4, cb = c0 + 64 + tx * 4;
        float4 va = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        float4 vb = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        if (fx) {  // FM backward fused in: out = dx + g_fm[row] * (S[row, c % k] - x[row, c])  (layers.py:457-478)
          const float gf = fg[row];
          if (ca < d) {
            const float4 xv = ld4(fx + row * fx_ld + ca), sv = ld4(fS + row * fk + (ca % fk));
            va.x += gf * (sv.x - xv.x); va.y += gf * (sv.y - xv.y); va.z += gf * (sv.z - xv.z); va.w += gf * (sv.w - xv.w);
          }
          if (cb < d) {
            const float4 xv = ld4(fx + row * fx_ld + cb), sv = ld4(fS + row * fk + (cb % fk));
            vb.x += gf * (sv.x - xv.x); vb.y += gf * (sv.y - xv.y); vb.z += gf * (sv.z - xv.z); vb.w += gf * (sv.w - xv.w);
          }
        }
        if (ca < d_ld) st4(o + ca, va);
        if (cb < d_ld) st4(o + cb, vb);
      }
    }
    if (more) {
      store_t(Ws0 + (((ct - ct_lo) & 1) ^ 1) * NT * DX_LD, wnext);
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dW: CTA (feature tile of 128, batch slab) -> partial[slab][f][n]; chunks of 32 samples double-buffered with cp.async.
// 128 threads, thread: 8 features x (NT / 8) outputs.  The slabs are summed in slab order by linear_dw_final_kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DW_BF = 128, DW_BK = 32;

template <int NT>
__global__ void __launch_bounds__(128) linear_dw_kernel(const float* __restrict__ x, int64_t ld,
                                                        const float* __restrict__ g, int64_t B, int K, int N,
                                                        int64_t slab_rows, float* __restrict__ partial, int Kpad) {
  constexpr int NV = NT / 32;  // float4 output groups per thread along n
  extern __shared__ float smem_dw[];
  float* Xs = smem_dw;                          // [2][DW_BK][DW_BF]
  float* Gs = smem_dw + 2 * DW_BK * DW_BF;      // [2][DW_BK][NT]
  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * DW_BF;
  const int64_t b_lo = (int64_t)blockIdx.y * slab_rows;
  const int64_t b_hi = b_lo + slab_rows < B ? b_lo + slab_rows : B;
  const int n_chunks = (int)((b_hi - b_lo + DW_BK - 1) / DW_BK);
  const uint32_t xs_u32 = (uint32_t)__cvta_generic_to_shared(Xs), gs_u32 = (uint32_t)__cvta_generic_to_shared(Gs);

  auto issue = [&](int ch) {
    const int buf = ch & 1;
    const int64_t b0 = b_lo + (int64_t)ch * DW_BK;
#pragma unroll
    for (int i = 0; i < DW_BK * DW_BF / 4 / 128; ++i) {  // 8 float4 of the x chunk per thread
      const int idx = tid + i * 128;
      const int r = idx >> 5, c4 = idx & 31;
      const int64_t b = b0 + r;
      const int f = f0 + 4 * c4;
      const bool ok = b < b_hi && f < ld;  // ld % 4 == 0: a float4 never straddles the row end
      cp_async16_zfill(xs_u32 + (uint32_t)((buf * DW_BK + r) * DW_BF + 4 * c4) * 4u, ok ? x + b * ld + f : x, ok);
    }
    for (int idx = tid; idx < DW_BK * NT / 4; idx += 128) {
      const int r = idx / (NT / 4), c4 = idx - r * (NT / 4);
      const int64_t b = b0 + r;
      const bool ok = b < b_hi && 4 * c4 < N;
      cp_async16_zfill(gs_u32 + (uint32_t)((buf * DW_BK + r) * NT + 4 * c4) * 4u, ok ? g + b * N + 4 * c4 : g, ok);
    }
    cp_async_commit_group();
  };

  const int tf = tid >> 3, tn = tid & 7;  // features tf*8 .. +7, outputs tn*4 .. +3 (+32 for NT = 64)
  float acc[8][4 * NV];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * NV; ++j) acc[i][j] = 0.f;
  if (n_chunks > 0) issue(0);
  for (int ch = 0; ch < n_chunks; ++ch) {
    if (ch + 1 < n_chunks) issue(ch + 1);
    else cp_async_commit_group();
    cp_async_wait_group<1>();
    __syncthreads();
    const float* xb = Xs + (ch & 1) * DW_BK * DW_BF + tf * 8;
    const float* gb = Gs + (ch & 1) * DW_BK * NT + tn * 4;
#pragma unroll 4
    for (int r = 0; r < DW_BK; ++r) {
      const float4 x0 = ld4(xb + r * DW_BF), x1 = ld4(xb + r * DW_BF + 4);
      const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 gv = ld4(gb + r * NT + 32 * v);
        const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][4 * v + j] = fmaf(xa[i], ga[j], acc[i][4 * v + j]);
      }
    }
    __syncthreads();  // the buffer is refilled by the next iteration's issue
  }
  float* out = partial + ((int64_t)blockIdx.y * Kpad + f0 + tf * 8) * NT;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int v = 0; v < NV; ++v)
      st4(out + i * NT + 32 * v + tn * 4,
          make_float4(acc[i][4 * v], acc[i][4 * v + 1], acc[i][4 * v + 2], acc[i][4 * v + 3]));
}

__global__ void __launch_bounds__(256) linear_dw_final_kernel(const float* __restrict__ partial, int slabs, int Kpad,
                                                              int NT, int K, int N, float* __restrict__ dW) {
  const int64_t total = (int64_t)K * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int64_t f = i / N;
    float acc = 0.f;
    for (int s = 0; s < slabs; ++s) acc += partial[((int64_t)s * Kpad + f) * NT + n];
    dW[i] = acc;
  }
}

struct DwLayout {
  int NT, n_ftiles, Kpad, slabs;
  int64_t slab_rows;
  size_t total;
};

static DwLayout dw_layout(int64_t B, int K, int N) {
  DwLayout L;
  L.NT = N <= 32 ? 32 : 64;
  L.n_ftiles = (int)ceil_div(K, DW_BF);
  L.Kpad = L.n_ftiles * DW_BF;
  int64_t slabs = (5 * RM_NUM_SMS) / L.n_ftiles;  // about five 128-thread CTAs per SM (40 KB smem each) in one wave
  const int64_t max_slabs = ceil_div(B > 0 ? B : 1, 256);
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  L.slab_rows = ceil_div(ceil_div(B > 0 ? B : 1, slabs), DW_BK) * DW_BK;
  L.slabs = (int)ceil_div(B > 0 ? B : 1, L.slab_rows);
  L.total = align_up((size_t)L.slabs * L.Kpad * L.NT * sizeof(float), 256);
  return L;
}

}  // namespace rm

extern "C" {

static int linear_dx_launch(const float* g, int64_t B, int32_t N, const float* W, int32_t d, float* dx, int64_t d_ld,
                            const float* fx, int64_t fx_ld, const float* fS, const float* fg, int32_t fk,
                            void* stream) {
  using namespace rm;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(B, DX_BM), (unsigned)ceil_div(ceil_div(d_ld, DX_BN), DX_CT));
  if (N <= 32) {
    const size_t smem = (size_t)3 * 32 * DX_LD * sizeof(float);
    RM_SMEM_ATTR_ONCE(smem, linear_dx_kernel<32>);
    linear_dx_kernel<32><<<grid, 256, smem, st>>>(g, B, N, W, d, dx, d_ld, fx, fx_ld, fS, fg, fk);
  } else {
    const size_t smem = (size_t)3 * 64 * DX_LD * sizeof(float);
    RM_SMEM_ATTR_ONCE(smem, linear_dx_kernel<64>);
    linear_dx_kernel<64><<<grid, 256, smem, st>>>(g, B, N, W, d, dx, d_ld, fx, fx_ld, fS, fg, fk);
  }
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_linear_bwd_input(const float* g, int64_t B, int32_t N, const float* W, int32_t d, float* dx, int64_t d_ld,
                        void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && N > 0 && d > 0 && d_ld >= d, "bad shape");
  if (B == 0) return 0;
  RM_CHECK_ARG(g && W && dx, "null pointer");
  RM_UNSUPPORTED(N <= 64 && N % 4 == 0 && d_ld % 4 == 0 && aligned16(g) && aligned16(W) && aligned16(dx),
                 "narrow-layer input gradient needs N <= 64, N % 4 == 0 and 16-byte aligned rows");
  return linear_dx_launch(g, B, N, W, d, dx, d_ld, nullptr, 0, nullptr, nullptr, 4, stream);
}

int rm_linear_bwd_input_fm(const float* g, int64_t B, int32_t N, const float* W, int32_t m, int32_t k, const float* x,
                           int64_t x_ld, const float* sum, const float* g_fm, float* out, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && N > 0 && m > 0 && k > 0 && x_ld >= (int64_t)m * k, "bad shape");
  if (B == 0) return 0;
  RM_CHECK_ARG(g && W && x && sum && g_fm && out, "null pointer");
  RM_UNSUPPORTED(N <= 64 && N % 4 == 0 && k % 4 == 0 && x_ld % 4 == 0 && aligned16(g) && aligned16(W) &&
                     aligned16(x) && aligned16(sum) && aligned16(out),
                 "fused FM backward needs N <= 64, N % 4 == 0, k % 4 == 0 and 16-byte aligned rows");
  const int32_t d = m * k;
  return linear_dx_launch(g, B, N, W, d, out, d, x, x_ld, sum, g_fm, k, stream);
}

size_t rm_linear_bwd_weight_workspace_bytes(int64_t B, int32_t K, int32_t N) {
  if (B < 0 || K <= 0 || N <= 0 || N > 64) return 256;
  return rm::dw_layout(B, K, N).total;
}

int rm_linear_bwd_weight(const float* x, int64_t ld, const float* g, int64_t B, int32_t K, int32_t N, float* dW,
                         void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && K > 0 && N > 0 && ld >= K, "bad shape");
  RM_CHECK_ARG(dW && workspace && (B == 0 || (x && g)), "null pointer");
  RM_UNSUPPORTED(N <= 64 && N % 4 == 0 && ld % 4 == 0 && (B == 0 || (aligned16(x) && aligned16(g))) &&
                     aligned16(workspace),
                 "narrow-layer weight gradient needs N <= 64, N % 4 == 0 and 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    RM_CUDA(cudaMemsetAsync(dW, 0, (size_t)K * N * sizeof(float), st));
    return 0;
  }
  const DwLayout L = dw_layout(B, K, N);
  if (workspace_bytes < L.total) {
    set_error("rm_linear_bwd_weight: workspace %zu < required %zu", workspace_bytes, L.total);
    return RM_E_WORKSPACE;
  }
  float* partial = (float*)workspace;
  dim3 grid((unsigned)L.n_ftiles, (unsigned)L.slabs);
  const size_t smem = (size_t)2 * DW_BK * (DW_BF + L.NT) * sizeof(float);
  if (L.NT == 32) {
    linear_dw_kernel<32><<<grid, 128, smem, st>>>(x, ld, g, B, K, N, L.slab_rows, partial, L.Kpad);
  } else {
    RM_SMEM_ATTR_ONCE(smem, linear_dw_kernel<64>);
    linear_dw_kernel<64><<<grid, 128, smem, st>>>(x, ld, g, B, K, N, L.slab_rows, partial, L.Kpad);
  }
  RM_LAUNCH_CHECK();
  linear_dw_final_kernel<<<grid_for((int64_t)K * N, 256, 8), 256, 0, st>>>(partial, L.slabs, L.Kpad, L.NT, K, N, dW);
  RM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
