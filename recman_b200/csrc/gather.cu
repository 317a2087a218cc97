// K1: multi-table embedding gather (+ sqrtn pooled variant) and the fused
// gather + FM + first-order front end of DeepFM.
//
// Replaces tf.nn.embedding_lookup / embedding_lookup_sparse + tf.concat at
// recman/tf/core/layers.py:117-128, :144-169, :238-261 and, fused, FMLayer
// (:457-478) and LinearLayer (:330-347).
//
// Roofline: HBM.  Algorithmic bytes per (sample, field): 8 (id) + 4k (row read)
// + 4k (row write).  Rows are read with 128-bit ld.global.nc.L1::no_allocate,
// LPR = k/4 lanes per row, U independent rows in flight per lane.
#include "common.cuh"

namespace rm {

// ---------------------------------------------------------------------------
// vector path: k % 4 == 0, 16-byte aligned pointers, out_stride % 4 == 0
// ---------------------------------------------------------------------------
template <int LPR, int U>
__global__ void __launch_bounds__(256) gather_vec_kernel(const float* __restrict__ table,
                                                         const int64_t* __restrict__ offs,
                                                         const int64_t* __restrict__ ids, uint32_t N, uint32_t m,
                                                         int k, float* __restrict__ out, int64_t out_stride,
                                                         int32_t* status) {
  constexpr int GPW = 32 / LPR;  // row groups per warp
  constexpr int RPW = GPW * U;   // rows per warp per iteration
  const int lane = threadIdx.x & 31;
  const int lir = lane % LPR;  // lane in row
  const int giw = lane / LPR;  // group in warp
  const int k4 = k >> 2;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;

  for (uint64_t base = (uint64_t)warp * RPW; base < N; base += (uint64_t)n_warps * RPW) {
    const float* src[U];
    float* dst[U];
    bool live[U], ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t p64 = base + (uint64_t)u * GPW + giw;  // a warp's u-th access covers GPW consecutive rows
      live[u] = p64 < N;
      ok[u] = false;
      src[u] = table;
      dst[u] = out;
      if (live[u]) {
        const uint32_t p = (uint32_t)p64;
        const uint32_t b = p / m, f = p - b * m;
        const int64_t id = ids[p];
        const int64_t lo = offs[f], hi = offs[f + 1];
        ok[u] = (id >= 0) && (id < hi - lo);
        src[u] = table + (lo + (ok[u] ? id : 0)) * (int64_t)k;
        dst[u] = out + (int64_t)b * out_stride + (int64_t)f * k;
        if (!ok[u] && lir == 0 && status) atomicOr(status, 1);
      }
    }
    for (int c = lir; c < k4; c += LPR) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[u]) v[u] = ldg_stream4(src[u] + 4 * c);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (live[u]) st4(dst[u] + 4 * c, v[u]);
    }
  }
}

// scalar path: any k (k == 1 is the bias / linear-weight lookup), any alignment
__global__ void __launch_bounds__(256) gather_scalar_kernel(const float* __restrict__ table,
                                                            const int64_t* __restrict__ offs,
                                                            const int64_t* __restrict__ ids, uint64_t total, uint32_t m,
                                                            uint32_t k, float* __restrict__ out, int64_t out_stride,
                                                            int32_t* status) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = i / k;
    const uint32_t c = (uint32_t)(i - p * k);
    const uint64_t b = p / m;
    const uint32_t f = (uint32_t)(p - b * m);
    const int64_t id = ids[p];
    const int64_t lo = offs[f], hi = offs[f + 1];
    const bool ok = (id >= 0) && (id < hi - lo);
    float v = 0.f;
    if (ok)
      v = ldg_stream1(table + (lo + id) * (int64_t)k + c);
    else if (status && c == 0)
      atomicOr(status, 1);
    out[b * out_stride + (int64_t)f * k + c] = v;
  }
}

// sqrtn pooled gather: one warp per sample, lanes over the k columns, values in CSR order
__global__ void __launch_bounds__(256) gather_pooled_kernel(const float* __restrict__ table, int64_t row_offset,
                                                            int64_t table_rows, const int64_t* __restrict__ values,
                                                            const int64_t* __restrict__ offsets, int64_t B, int k,
                                                            float* __restrict__ out, int64_t out_stride,
                                                            int32_t* status) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += n_warps) {
    const int64_t lo = offsets[b], hi = offsets[b + 1];
    const int64_t n = hi - lo;
    for (int c = lane; c < k; c += 32) {
      float acc = 0.f;
      for (int64_t j = lo; j < hi; ++j) {
        const int64_t id = values[j];
        if (id >= 0 && id < table_rows)
          acc += ldg_stream1(table + (row_offset + id) * (int64_t)k + c);
        else if (status)
          atomicOr(status, 1);
      }
      // oracle: acc / sqrt(n).  Keep a true division so the result is bit-identical.
      out[b * out_stride + c] = n > 0 ? acc / sqrtf((float)n) : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------
// fused gather + FM + first-order: one row group (LPR lanes) per sample
// ---------------------------------------------------------------------------
template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) gather_fm_kernel(
    const float* __restrict__ table, const float* __restrict__ bias_table, const float* __restrict__ lin_table,
    const int64_t* __restrict__ offs, const int64_t* __restrict__ ids, const float* __restrict__ dense,
    const float* __restrict__ lin_dense, int n_dense, int64_t B, int m, int k, float* __restrict__ x, int64_t ld,
    float* __restrict__ fm_out, float* __restrict__ lin_out, float* __restrict__ sum_out, int32_t* status) {
  const int lir = threadIdx.x % LPR;
  const int k4 = k >> 2;
  const bool col_ok = lir < k4;  // LPR = pow2ceil(k/4): upper lanes idle when k/4 is not a power of two
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  // every lane of a warp must run the same number of iterations (shuffles below)
  const int64_t iters = (B + n_groups - 1) / n_groups;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t b = group + it * n_groups;
    const bool live = b < B;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 Q = make_float4(0.f, 0.f, 0.f, 0.f);
    float bias_acc = 0.f, lin_acc = 0.f;
    if (live) {
      const int64_t* my_ids = ids + b * m;
      float* xrow = x + b * ld;
      for (int f0 = 0; f0 < m; f0 += U) {
        int64_t row[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int f = f0 + u;
          ok[u] = false;
          row[u] = 0;
          if (f < m) {
            const int64_t id = my_ids[f];
            const int64_t lo = offs[f], hi = offs[f + 1];
            ok[u] = (id >= 0) && (id < hi - lo);
            row[u] = lo + (ok[u] ? id : 0);
            if (!ok[u] && lir == 0 && status) atomicOr(status, 1);
          }
        }
        float4 v[U];
        float bv[U], lv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          bv[u] = 0.f;
          lv[u] = 0.f;
          if (ok[u]) {
            if (col_ok) v[u] = ldg_stream4(table + row[u] * (int64_t)k + 4 * lir);
            if (lir == (u % LPR)) {  // spread the k=1 lookups over the lanes of the group
              if (bias_table) bv[u] = ldg_stream1(bias_table + row[u]);
              if (lin_table) lv[u] = ldg_stream1(lin_table + row[u]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int f = f0 + u;
          if (f < m) {
            if (col_ok) st4(xrow + (int64_t)f * k + 4 * lir, v[u]);
            S.x += v[u].x; S.y += v[u].y; S.z += v[u].z; S.w += v[u].w;
            Q.x += v[u].x * v[u].x; Q.y += v[u].y * v[u].y; Q.z += v[u].z * v[u].z; Q.w += v[u].w * v[u].w;
            bias_acc += bv[u];
            lin_acc += lv[u];
          }
        }
      }
      // dense tail of the DNN input row + its first-order contribution
      for (int j = lir; j < n_dense; j += LPR) {
        const float dv = dense[b * n_dense + j];
        xrow[(int64_t)m * k + j] = dv;
        if (lin_dense) lin_acc += dv * lin_dense[j];
      }
      if (sum_out && col_ok) st4(sum_out + b * k + 4 * lir, S);
    }
    float second = 0.5f * (S.x * S.x - Q.x) + 0.5f * (S.y * S.y - Q.y) + 0.5f * (S.z * S.z - Q.z) +
                   0.5f * (S.w * S.w - Q.w);
    second = group_sum<LPR>(second);
    bias_acc = group_sum<LPR>(bias_acc);
    lin_acc = group_sum<LPR>(lin_acc);
    if (live && lir == 0) {
      if (fm_out) fm_out[b] = bias_acc + second;
      if (lin_out) lin_out[b] = lin_acc;
    }
  }
}

template <int LPR>
static int launch_gather_vec(const float* table, const int64_t* offs, const int64_t* ids, int64_t N, int m, int k,
                             float* out, int64_t out_stride, int32_t* status, cudaStream_t st) {
  constexpr int U = 4;
  constexpr int rows_per_cta = 8 * (32 / LPR) * U;
  const int grid = grid_for(N, rows_per_cta, 8);
  gather_vec_kernel<LPR, U><<<grid, 256, 0, st>>>(table, offs, ids, (uint32_t)N, (uint32_t)m, k, out, out_stride,
                                                  status);
  RM_LAUNCH_CHECK();
  return 0;
}

template <int LPR>
static int launch_gather_fm(const float* table, const float* bias_table, const float* lin_table, const int64_t* offs,
                            const int64_t* ids, const float* dense, const float* lin_dense, int n_dense, int64_t B,
                            int m, int k, float* x, int64_t ld, float* fm_out, float* lin_out, float* sum_out,
                            int32_t* status, cudaStream_t st) {
  const int groups_per_cta = 256 / LPR;
  const int grid = grid_for(B, groups_per_cta, 8);
#define RM_GFM(UU, MB)                                                                                              \
  gather_fm_kernel<LPR, UU, MB><<<grid, 256, 0, st>>>(table, bias_table, lin_table, offs, ids, dense, lin_dense,     \
                                                      n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status)
  RM_GFM(4, 4);  // 4 rows in flight per lane, 4 CTAs / SM: the measured best of (8,2) (4,3) (2,6) (4,4) on B200
#undef RM_GFM
  RM_LAUNCH_CHECK();
  return 0;
}

static inline int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace rm

extern "C" {

int rm_gather_fwd(const float* table, const int64_t* table_offsets, const int64_t* ids, int64_t B, int32_t m,
                  int32_t k, float* out, int64_t out_stride, int32_t* status, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(B >= 0 && m > 0 && k > 0, "bad shape");
  RM_CHECK_ARG(out_stride >= (int64_t)m * k, "out_stride smaller than m*k");
  const int64_t N = B * m;
  RM_UNSUPPORTED(N < (int64_t)1 << 31, "B*m must be < 2^31");
  if (N == 0) return 0;  // empty batch: nothing to do, pointers may be null
  RM_CHECK_ARG(table && table_offsets && ids && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (k % 4 == 0) && (out_stride % 4 == 0) && aligned16(table) && aligned16(out);
  if (vec) {
    int lpr = pow2ceil(k / 4);
    if (lpr > 32) lpr = 32;
    switch (lpr) {
      case 1: return launch_gather_vec<1>(table, table_offsets, ids, N, m, k, out, out_stride, status, st);
      case 2: return launch_gather_vec<2>(table, table_offsets, ids, N, m, k, out, out_stride, status, st);
      case 4: return launch_gather_vec<4>(table, table_offsets, ids, N, m, k, out, out_stride, status, st);
      case 8: return launch_gather_vec<8>(table, table_offsets, ids, N, m, k, out, out_stride, status, st);
      case 16: return launch_gather_vec<16>(table, table_offsets, ids, N, m, k, out, out_stride, status, st);
      default: return launch_gather_vec<32>(table, table_offsets, ids, N, m, k, out, out_stride, status, st);
    }
  }
  const uint64_t total = (uint64_t)N * (uint64_t)k;
  const int grid = grid_for((int64_t)total, 256, 8);
  gather_scalar_kernel<<<grid, 256, 0, st>>>(table, table_offsets, ids, total, (uint32_t)m, (uint32_t)k, out,
                                             out_stride, status);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_gather_pooled_fwd(const float* table, int64_t row_offset, int64_t table_rows, const int64_t* values,
                         const int64_t* offsets, int64_t B, int32_t k, float* out, int64_t out_stride,
                         int32_t* status, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(table && offsets && out, "null pointer");
  RM_CHECK_ARG(B >= 0 && k > 0 && row_offset >= 0 && table_rows > 0, "bad shape");
  RM_CHECK_ARG(out_stride >= k, "out_stride smaller than k");
  if (B == 0) return 0;
  const int grid = grid_for(B, 8, 8);
  gather_pooled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table, row_offset, table_rows, values, offsets, B, k,
                                                               out, out_stride, status);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_gather_fm_fwd(const float* table, const float* bias_table, const float* lin_table,
                     const int64_t* table_offsets, const int64_t* ids, const float* dense, const float* lin_dense,
                     int32_t n_dense, int64_t B, int32_t m, int32_t k, float* x, int64_t ld, float* fm_out,
                     float* lin_out, float* sum_out, int32_t* status, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(table && table_offsets && ids && x, "null pointer");
  RM_CHECK_ARG(B >= 0 && m > 0 && k > 0 && n_dense >= 0, "bad shape");
  RM_CHECK_ARG(n_dense == 0 || dense, "dense pointer missing");
  RM_CHECK_ARG(ld >= (int64_t)m * k + n_dense, "ld smaller than m*k+n_dense");
  RM_UNSUPPORTED(k % 4 == 0 && k <= 128, "fused front end needs k % 4 == 0 and k <= 128");
  RM_UNSUPPORTED(ld % 4 == 0 && aligned16(table) && aligned16(x) && (!sum_out || aligned16(sum_out)),
                 "fused front end needs 16-byte aligned rows (ld % 4 == 0)");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (pow2ceil(k / 4)) {
    case 1: return launch_gather_fm<1>(table, bias_table, lin_table, table_offsets, ids, dense, lin_dense, n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status, st);
    case 2: return launch_gather_fm<2>(table, bias_table, lin_table, table_offsets, ids, dense, lin_dense, n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status, st);
    case 4: return launch_gather_fm<4>(table, bias_table, lin_table, table_offsets, ids, dense, lin_dense, n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status, st);
    case 8: return launch_gather_fm<8>(table, bias_table, lin_table, table_offsets, ids, dense, lin_dense, n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status, st);
    case 16: return launch_gather_fm<16>(table, bias_table, lin_table, table_offsets, ids, dense, lin_dense, n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status, st);
    default: return launch_gather_fm<32>(table, bias_table, lin_table, table_offsets, ids, dense, lin_dense, n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status, st);
  }
}

}  // extern "C"
