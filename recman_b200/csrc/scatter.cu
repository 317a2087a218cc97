// K2: deterministic sparse embedding-gradient scatter-add.
//
// Replaces TensorFlow's autodiff of tf.nn.embedding_lookup
// (recman/tf/core/layers.py:117-128): an IndexedSlices gradient whose duplicate
// rows are summed.  Stock frameworks do this with unsorted fp32 atomics
// (non-deterministic); here it is sort -> run-length encode -> segmented sum
// in ascending-position order, no atomics:
//
//   plan   (ids only; can overlap the forward on a side stream)
//     key[p] = table_offsets[p % m] + ids[p], stable LSD radix sort of (key, p)
//     (cub::DeviceRadixSort restricted to the significant key bits), heads of
//     runs selected with cub::DeviceSelect -> seg_start, uniq_rows, n_unique.
//   reduce (gradient rows)
//     one row group (k/4 lanes, 128-bit loads) per unique row walks its
//     segment in order; U rows in flight.
//
// Roofline: HBM.  Algorithmic bytes: 8 + 4k read per id, 4k + 8 written per
// unique row; the sort traffic is overhead, not algorithmic.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"

namespace rm {

__global__ void __launch_bounds__(256) make_keys_kernel(const int64_t* __restrict__ ids,
                                                        const int64_t* __restrict__ offs, uint32_t N, uint32_t m,
                                                        uint32_t* __restrict__ keys, int32_t* __restrict__ pos) {
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
    int64_t key = ids[p];
    if (offs) key += offs[p % m];
    keys[p] = (uint32_t)key;
    pos[p] = (int32_t)p;
  }
}

struct HeadOfRun {
  const uint32_t* keys;
  __host__ __device__ __forceinline__ bool operator()(const int32_t& i) const {
    return i == 0 || keys[i] != keys[i - 1];
  }
};

__global__ void __launch_bounds__(256) finish_plan_kernel(const uint32_t* __restrict__ sorted_keys,
                                                          int32_t* __restrict__ seg_start,
                                                          const int32_t* __restrict__ n_unique, int32_t N,
                                                          int64_t* __restrict__ uniq_rows) {
  const int32_t U = *n_unique;
  for (int32_t u = blockIdx.x * blockDim.x + threadIdx.x; u <= U; u += gridDim.x * blockDim.x) {
    if (u == U)
      seg_start[U] = N;
    else
      uniq_rows[u] = (int64_t)sorted_keys[seg_start[u]];
  }
}

struct PlanWorkspace {
  uint32_t* keys_in;
  uint32_t* keys_out;
  int32_t* pos_in;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static PlanWorkspace plan_layout(int64_t N, void* base) {
  PlanWorkspace w;
  size_t sort_bytes = 0, select_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)N, 0, 32);
  thrust::counting_iterator<int32_t> counting(0);
  HeadOfRun pred{nullptr};
  cub::DeviceSelect::If(nullptr, select_bytes, counting, (int32_t*)nullptr, (int32_t*)nullptr, (int)N, pred);
  w.cub_bytes = sort_bytes > select_bytes ? sort_bytes : select_bytes;
  size_t off = 0;
  char* b = (char*)base;
  const size_t arr = align_up((size_t)N * 4, 256);
  w.keys_in = (uint32_t*)(b + off); off += arr;
  w.keys_out = (uint32_t*)(b + off); off += arr;
  w.pos_in = (int32_t*)(b + off); off += arr;
  w.cub_temp = (void*)(b + off); off += align_up(w.cub_bytes, 256);
  w.total = off;
  return w;
}

// ---------------------------------------------------------------------------
// segmented reduce
// ---------------------------------------------------------------------------
// FUSED == false : row(p) = grad[(p/m)*ld + (p%m)*k + :]
// FUSED == true  : row(p) = dx[..] + g_fm[b]*(S[b,:] - x[..]); also k=1 sums of g_fm / g_lin
//
// A row group owns SEG segments at a time and first issues the loads of all their FIRST elements (with uniform ids
// almost every segment is a singleton, so this is where the memory-level parallelism comes from: the dependent chain
// seg_start -> sorted_pos -> rows is walked for SEG segments concurrently); the remaining elements of longer
// segments follow, U at a time.  Every segment is still summed in ascending position order.
template <bool FUSED>
struct RowSrc {
  const float* grad;
  const float* x;
  const float* sum;
  const float* g_fm;
  const float* g_lin;
  int64_t ld;
  uint32_t m;
  int k;
  __device__ __forceinline__ void load(uint32_t p, int c, float4& v, float& gf, float& gl) const {
    const uint32_t b = p / m, f = p - b * m;
    const int64_t o = (int64_t)b * ld + (int64_t)f * k + 4 * c;
    v = grad ? ld4(grad + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    gf = 0.f;
    gl = 0.f;
    if (FUSED) {
      if (g_fm) {
        gf = g_fm[b];
        const float4 xv = ld4(x + o);
        const float4 sv = ld4(sum + (int64_t)b * k + 4 * c);
        v.x += gf * (sv.x - xv.x);
        v.y += gf * (sv.y - xv.y);
        v.z += gf * (sv.z - xv.z);
        v.w += gf * (sv.w - xv.w);
      }
      if (g_lin) gl = g_lin[b];
    }
  }
};

template <int LPR, int SEG, int U, bool FUSED>
__global__ void __launch_bounds__(256) segment_reduce_kernel(
    const float* __restrict__ grad, const float* __restrict__ x, int64_t ld, const float* __restrict__ sum,
    const float* __restrict__ g_fm, const float* __restrict__ g_lin, uint32_t m, int k,
    const int32_t* __restrict__ sorted_pos, const int32_t* __restrict__ seg_start,
    const int32_t* __restrict__ n_unique, float* __restrict__ out_rows, float* __restrict__ out_bias,
    float* __restrict__ out_lin) {
  const int lir = threadIdx.x % LPR;
  const int k4 = k >> 2;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  const int32_t NU = *n_unique;
  const RowSrc<FUSED> src{grad, x, sum, g_fm, g_lin, ld, m, k};
  for (int64_t u0 = group * SEG; u0 < NU; u0 += n_groups * SEG) {
    int32_t s[SEG], e[SEG];
#pragma unroll
    for (int t = 0; t < SEG; ++t) {
      const bool live = u0 + t < NU;
      s[t] = live ? seg_start[u0 + t] : 0;
      e[t] = live ? seg_start[u0 + t + 1] : 0;
    }
    for (int c = lir; c < k4; c += LPR) {  // lane 0 always owns column chunk 0, so it also carries the k=1 sums
      uint32_t p0[SEG];
#pragma unroll
      for (int t = 0; t < SEG; ++t) p0[t] = s[t] < e[t] ? (uint32_t)sorted_pos[s[t]] : 0u;
      float4 acc[SEG];
      float ba[SEG], la[SEG];
#pragma unroll
      for (int t = 0; t < SEG; ++t) {
        acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        ba[t] = 0.f;
        la[t] = 0.f;
        if (s[t] < e[t]) src.load(p0[t], c, acc[t], ba[t], la[t]);
      }
#pragma unroll
      for (int t = 0; t < SEG; ++t) {
        for (int32_t j0 = s[t] + 1; j0 < e[t]; j0 += U) {  // longer segments: U rows in flight, added in order
          float4 v[U];
          float gf[U], gl[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            gf[i] = 0.f;
            gl[i] = 0.f;
            if (j0 + i < e[t]) src.load((uint32_t)sorted_pos[j0 + i], c, v[i], gf[i], gl[i]);
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            if (j0 + i < e[t]) {
              acc[t].x += v[i].x; acc[t].y += v[i].y; acc[t].z += v[i].z; acc[t].w += v[i].w;
              ba[t] += gf[i];
              la[t] += gl[i];
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < SEG; ++t) {
        if (u0 + t < NU) {
          if (out_rows) st4(out_rows + (u0 + t) * k + 4 * c, acc[t]);
          if (FUSED && c == 0) {
            if (out_bias) out_bias[u0 + t] = ba[t];
            if (out_lin) out_lin[u0 + t] = la[t];
          }
        }
      }
    }
  }
}

// scalar fallback (k % 4 != 0 or misaligned): one thread per (unique row, column)
__global__ void __launch_bounds__(256) segment_reduce_scalar_kernel(const float* __restrict__ grad, int64_t ld,
                                                                    uint32_t m, uint32_t k,
                                                                    const int32_t* __restrict__ sorted_pos,
                                                                    const int32_t* __restrict__ seg_start,
                                                                    const int32_t* __restrict__ n_unique,
                                                                    float* __restrict__ out_rows) {
  const int64_t total = (int64_t)(*n_unique) * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = i / k;
    const uint32_t c = (uint32_t)(i - u * k);
    float acc = 0.f;
    for (int32_t j = seg_start[u]; j < seg_start[u + 1]; ++j) {
      const uint32_t p = (uint32_t)sorted_pos[j];
      const uint32_t b = p / m, f = p - b * m;
      acc += grad[(int64_t)b * ld + (int64_t)f * k + c];
    }
    out_rows[i] = acc;
  }
}

static inline int pow2ceil__(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int LPR, bool FUSED>
static int launch_segment_reduce(const float* grad, const float* x, int64_t ld, const float* sum, const float* g_fm,
                                 const float* g_lin, int m, int k, int64_t N, const int32_t* sorted_pos,
                                 const int32_t* seg_start, const int32_t* n_unique, float* out_rows, float* out_bias,
                                 float* out_lin, cudaStream_t st) {
  const int variant = tune_variant("RM_TUNE_SEGRED", FUSED ? 2 : 4);  // measured best on B200 (profiles/r1_kbench.json)
#define RM_SRK(SEG)                                                                                                  \
  segment_reduce_kernel<LPR, SEG, 4, FUSED><<<grid_for(N, (256 / LPR) * SEG, 8), 256, 0, st>>>(                       \
      grad, x, ld, sum, g_fm, g_lin, (uint32_t)m, k, sorted_pos, seg_start, n_unique, out_rows, out_bias, out_lin)
  // n_unique <= N lives on the device: the grid is sized for the worst case
  if (variant == 1) RM_SRK(1);
  else if (variant == 2) RM_SRK(2);
  else RM_SRK(4);
#undef RM_SRK
  RM_LAUNCH_CHECK();
  return 0;
}

template <bool FUSED>
static int dispatch_segment_reduce(const float* grad, const float* x, int64_t ld, const float* sum, const float* g_fm,
                                   const float* g_lin, int m, int k, int64_t N, const int32_t* sorted_pos,
                                   const int32_t* seg_start, const int32_t* n_unique, float* out_rows,
                                   float* out_bias, float* out_lin, cudaStream_t st) {
  int lpr = pow2ceil__(k / 4);
  if (lpr > 32) lpr = 32;
#define RM_SR(L)                                                                                                    \
  case L:                                                                                                           \
    return launch_segment_reduce<L, FUSED>(grad, x, ld, sum, g_fm, g_lin, m, k, N, sorted_pos, seg_start, n_unique, \
                                           out_rows, out_bias, out_lin, st)
  switch (lpr) {
    RM_SR(1);
    RM_SR(2);
    RM_SR(4);
    RM_SR(8);
    RM_SR(16);
    default:
      return launch_segment_reduce<32, FUSED>(grad, x, ld, sum, g_fm, g_lin, m, k, N, sorted_pos, seg_start, n_unique,
                                              out_rows, out_bias, out_lin, st);
  }
#undef RM_SR
}

}  // namespace rm

extern "C" {

size_t rm_segment_plan_workspace_bytes(int64_t N) {
  if (N <= 0) return 256;
  return rm::plan_layout(N, nullptr).total;
}

int rm_segment_plan(const int64_t* ids, const int64_t* table_offsets, int64_t N, int32_t m, int64_t total_rows,
                    void* workspace, size_t workspace_bytes, int32_t* sorted_pos, int32_t* seg_start,
                    int64_t* uniq_rows, int32_t* n_unique, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(seg_start && n_unique, "null pointer");
  RM_CHECK_ARG(N >= 0 && m > 0 && total_rows > 0, "bad shape");
  RM_CHECK_ARG(N == 0 || (ids && sorted_pos && uniq_rows), "null pointer");
  RM_UNSUPPORTED(N < ((int64_t)1 << 31) - 1, "N must be < 2^31 - 1");
  RM_UNSUPPORTED(total_rows <= ((int64_t)1 << 32), "total_rows must be <= 2^32 (32-bit sort keys)");
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    RM_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(int32_t), st));
    RM_CUDA(cudaMemsetAsync(seg_start, 0, sizeof(int32_t), st));
    return 0;
  }
  RM_CHECK_ARG(workspace, "null workspace");
  PlanWorkspace w = plan_layout(N, workspace);
  if (workspace_bytes < w.total) {
    set_error("rm_segment_plan: workspace %zu < required %zu", workspace_bytes, w.total);
    return RM_E_WORKSPACE;
  }
  make_keys_kernel<<<grid_for(N, 256, 8), 256, 0, st>>>(ids, table_offsets, (uint32_t)N, (uint32_t)m, w.keys_in,
                                                        w.pos_in);
  RM_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 32 && ((int64_t)1 << end_bit) < total_rows) ++end_bit;
  size_t bytes = w.cub_bytes;
  RM_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, bytes, (const uint32_t*)w.keys_in, w.keys_out,
                                          (const int32_t*)w.pos_in, sorted_pos, (int)N, 0, end_bit, st));
  count_launch();
  thrust::counting_iterator<int32_t> counting(0);
  HeadOfRun pred{w.keys_out};
  bytes = w.cub_bytes;
  RM_CUDA(cub::DeviceSelect::If(w.cub_temp, bytes, counting, seg_start, n_unique, (int)N, pred, st));
  count_launch();
  finish_plan_kernel<<<grid_for(N + 1, 256, 8), 256, 0, st>>>(w.keys_out, seg_start, n_unique, (int32_t)N, uniq_rows);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_segment_reduce(const float* grad, int64_t ld, int32_t m, int32_t k, int64_t N, const int32_t* sorted_pos,
                      const int32_t* seg_start, const int32_t* n_unique, float* out_rows, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(grad && sorted_pos && seg_start && n_unique && out_rows, "null pointer");
  RM_CHECK_ARG(N >= 0 && m > 0 && k > 0 && ld >= (int64_t)m * k, "bad shape");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (k % 4 == 0) && (ld % 4 == 0) && aligned16(grad) && aligned16(out_rows);
  if (vec)
    return dispatch_segment_reduce<false>(grad, nullptr, ld, nullptr, nullptr, nullptr, m, k, N, sorted_pos, seg_start,
                                          n_unique, out_rows, nullptr, nullptr, st);
  segment_reduce_scalar_kernel<<<grid_for(N * k, 256, 8), 256, 0, st>>>(grad, ld, (uint32_t)m, (uint32_t)k, sorted_pos,
                                                                        seg_start, n_unique, out_rows);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_emb_fm_bwd(const float* dx, const float* x, int64_t ld, const float* sum, const float* g_fm, const float* g_lin,
                  int32_t m, int32_t k, int64_t N, const int32_t* sorted_pos, const int32_t* seg_start,
                  const int32_t* n_unique, float* out_rows, float* out_bias, float* out_lin, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(sorted_pos && seg_start && n_unique, "null pointer");
  RM_CHECK_ARG(N >= 0 && m > 0 && k > 0 && ld >= (int64_t)m * k, "bad shape");
  RM_CHECK_ARG(!g_fm || (x && sum), "g_fm needs x and sum");
  RM_UNSUPPORTED((k % 4 == 0) && (ld % 4 == 0) && (!dx || aligned16(dx)) && (!x || aligned16(x)) &&
                     (!sum || aligned16(sum)) && (!out_rows || aligned16(out_rows)),
                 "fused embedding backward needs k % 4 == 0 and 16-byte aligned rows");
  if (N == 0) return 0;
  return dispatch_segment_reduce<true>(dx, x, ld, sum, g_fm, g_lin, m, k, N, sorted_pos, seg_start, n_unique, out_rows,
                                       out_bias, out_lin, (cudaStream_t)stream);
}

}  // extern "C"
