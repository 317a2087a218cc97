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
#include "optim.cuh"

namespace rm {

// key = global table row of position p.  An id outside its table (the forward gather zero-fills such a row and raises the
// status flag) must never become a key inside another field's row range: it gets the sentinel key `total_rows`, sorts
// behind every real row and its segment is dropped from the plan (drop_sentinel_kernel).
__global__ void __launch_bounds__(256) make_keys_kernel(const int64_t* __restrict__ ids,
                                                        const int64_t* __restrict__ offs, uint32_t N, uint32_t m,
                                                        int64_t total_rows, uint32_t* __restrict__ keys,
                                                        int32_t* __restrict__ pos, int32_t* status) {
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
    int64_t key = ids[p];
    bool ok;
    if (offs) {
      const uint32_t f = p % m;
      const int64_t lo = offs[f], hi = offs[f + 1];
      ok = key >= 0 && key < hi - lo;
      key += lo;
    } else {
      ok = key >= 0 && key < total_rows;
    }
    if (!ok && status) atomicOr(status, 1);
    keys[p] = ok ? (uint32_t)key : (uint32_t)total_rows;
    pos[p] = (int32_t)p;
  }
}

// the (at most one) trailing segment of sentinel keys is not a table row: remove it from the unique-row count
__global__ void drop_sentinel_kernel(const uint32_t* __restrict__ sorted_keys, const int32_t* __restrict__ seg_start,
                                     int32_t* __restrict__ n_unique, int64_t total_rows) {
  const int32_t U = *n_unique;
  if (U > 0 && (int64_t)sorted_keys[seg_start[U - 1]] >= total_rows) *n_unique = U - 1;
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
  static constexpr bool kScalars = FUSED;
  static constexpr int kTailU = FUSED ? 2 : 4;  // rows in flight while walking the tail of a longer segment
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


// Row source for row-sharded tables: position gp = src_rank * rows_per_rank + local position; the row lives in rank
// src_rank's gradient buffer G[rows_per_rank, KP] = [g_0..g_{k-1} | g_bias | g_lin | 0 | 0], read over NVLink (peer
// memory) for src_rank != this rank.
#define RM_MAX_PEERS 8
struct PeerRowSrc {
  static constexpr bool kScalars = true;
  static constexpr int kTailU = 2;
  const float* G[RM_MAX_PEERS];
  const float* gscal;  // nullable: [W * samples_per_rank, 2] = (g_bias, g_lin) of EVERY rank's samples, local memory
  uint32_t rows_per_rank, m;
  int KP, k;
  __device__ __forceinline__ void load(uint32_t gp, int c, float4& v, float& gf, float& gl) const {
    const uint32_t r = gp / rows_per_rank, lp = gp - r * rows_per_rank;
    const float* row = G[r] + (int64_t)lp * KP;
    v = ldg_stream4(row + 4 * c);
    gf = 0.f;
    gl = 0.f;
    if (c == 0) {
      if (gscal) {  // the two k=1 gradients are per-sample values: all-gathered once, read from local L2 here
        const float2 t = *reinterpret_cast<const float2*>(gscal + 2 * (int64_t)(gp / m));
        gf = t.x;
        gl = t.y;
      } else {
        const float4 t = ldg_stream4(row + k);
        gf = t.x;
        gl = t.y;
      }
    }
  }
};

// Optional fused optimizer (N1): instead of (or in addition to) storing the summed rows, apply the stateless
// first-step update to table[uniq_rows[u]] in the same pass - the summed gradient never round-trips through HBM.
struct UpdateSink {
  float* table;       // nullptr = disabled
  float* bias_table;  // k = 1 tables in the same row numbering (nullable)
  float* lin_table;
  const int64_t* uniq_rows;
  OptParams o;
};

// Long segments (skewed ids: a hot row can collect thousands of positions of one batch).  Walking such a segment
// serially inside one row group would take milliseconds, so segments longer than SEG_LONG positions are not summed by
// the main kernel: it appends them to a device work list, a second kernel sums every SEG_CHUNK consecutive positions
// (chunk boundaries are fixed relative to the segment start), and a third adds the chunk partials in chunk order and
// finishes the row.  The result depends only on the sorted positions - deterministic run to run and independent of the
// world size; atomics are used only to hand out slots of the work list, never on data.  oracle.segment_sum_sorted
// restates the same association (sequential up to SEG_LONG, chunked above), which keeps the unfused kernel bit-exact.
constexpr int SEG_LONG = 16;
constexpr int SEG_CHUNK = 32;
struct LongWs {
  int32_t* counters;    // [0] long segments, [1] chunk slots handed out (nullptr = long path disabled)
  int32_t* long_u;      // [max_long]   unique-row index of the long segment
  int32_t* long_base;   // [max_long]   its first chunk slot
  int32_t* chunk_long;  // [max_chunks] long-list index of the chunk slot
  float* part_rows;     // [max_chunks, k]
  float* part_scal;     // [max_chunks, 2]
  int32_t max_long, max_chunks;
};

static LongWs long_layout(int64_t N, int k, void* base, size_t* total) {
  LongWs w;
  w.max_long = (int32_t)(N / (SEG_LONG + 1) + 1);
  w.max_chunks = (int32_t)(N / SEG_CHUNK + w.max_long + 1);
  char* b = (char*)base;
  size_t off = 0;
  w.counters = (int32_t*)(b + off); off += 256;
  w.long_u = (int32_t*)(b + off); off += align_up((size_t)w.max_long * 4, 256);
  w.long_base = (int32_t*)(b + off); off += align_up((size_t)w.max_long * 4, 256);
  w.chunk_long = (int32_t*)(b + off); off += align_up((size_t)w.max_chunks * 4, 256);
  w.part_rows = (float*)(b + off); off += align_up((size_t)w.max_chunks * k * 4, 256);
  w.part_scal = (float*)(b + off); off += align_up((size_t)w.max_chunks * 2 * 4, 256);
  *total = off;
  return w;
}

// store / update of one finished row: shared by the main kernel and the long-segment combine kernel
template <bool FUSED>
__device__ __forceinline__ void finish_row(int64_t u, int c, int k, const float4 acc, float ba, float la,
                                           float* __restrict__ out_rows, float* __restrict__ out_bias,
                                           float* __restrict__ out_lin, const UpdateSink& sink, int64_t urow,
                                           const float4 tv, float bo, float lo) {
  if (out_rows) st4(out_rows + u * k + 4 * c, acc);
  if (FUSED && c == 0) {
    if (out_bias) out_bias[u] = ba;
    if (out_lin) out_lin[u] = la;
  }
  if (sink.table) {
    float4 pv = tv;
    pv.x = opt_update(pv.x, acc.x, sink.o);
    pv.y = opt_update(pv.y, acc.y, sink.o);
    pv.z = opt_update(pv.z, acc.z, sink.o);
    pv.w = opt_update(pv.w, acc.w, sink.o);
    st4(sink.table + urow * k + 4 * c, pv);
    if (FUSED && c == 0) {
      if (sink.bias_table) sink.bias_table[urow] = opt_update(bo, ba, sink.o);
      if (sink.lin_table) sink.lin_table[urow] = opt_update(lo, la, sink.o);
    }
  }
}

// Work distribution: a warp owns chunks of 32 consecutive unique rows.  Lane l loads the index chain of row
// chunk*32 + l (seg_start -> sorted_pos head, uniq_rows) with coalesced loads, software-pipelined two chunks deep
// (bounds of chunk i+2 and heads of chunk i+1 are requested while the rows of chunk i are in flight), and the row
// groups (LPR lanes each) pick their segments' indices up by shuffle - so the three dependent latencies of the index
// chain are off the critical path and cost one load instruction per 32 segments.  A group has UB segments' rows in
// flight at a time; at step i the groups of a warp work on consecutive unique rows (coalesced output).
template <int LPR, int UBT, int U, int MINB, class Src>
__global__ void __launch_bounds__(256, MINB) segment_reduce_kernel(
    const Src src, int k, const int32_t* __restrict__ sorted_pos, const int32_t* __restrict__ seg_start,
    const int32_t* __restrict__ n_unique, float* __restrict__ out_rows, float* __restrict__ out_bias,
    float* __restrict__ out_lin, const UpdateSink sink, const LongWs lw) {
  constexpr bool FUSED = Src::kScalars;
  constexpr int GPW = 32 / LPR;              // row groups per warp
  constexpr int UB = UBT < LPR ? UBT : LPR;  // segments in flight per group (a group owns LPR segments per chunk)
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int lir = lane % LPR, giw = lane / LPR;
  const int c = lir;  // column chunk of this lane (k <= 128: exactly one; lane 0 owns chunk 0 and the k=1 sums)
  const bool col = c < (k >> 2);
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t stride = (((int64_t)gridDim.x * blockDim.x) >> 5) * 32;
  const int32_t NU = *n_unique;
  int64_t ub = warp * 32;
  if (ub >= NU) return;  // warp-uniform
  const bool upd = sink.table != nullptr;

  int32_t sA = 0, sB = 0, nA = 0, nB = 0;
  uint32_t hp = 0;
  int64_t ur = 0;
  {
    const int64_t u = ub + lane;
    if (u < NU) {
      sA = seg_start[u];
      sB = seg_start[u + 1];
      hp = (uint32_t)sorted_pos[sA];
      if (upd) ur = sink.uniq_rows[u];
    }
    const int64_t un = u + stride;
    if (un < NU) {
      nA = seg_start[un];
      nB = seg_start[un + 1];
    }
  }
  for (; ub < NU; ub += stride) {
    // ---- prefetch: heads of the next chunk (its bounds arrived during the previous chunk), bounds of the one after
    uint32_t nhp = 0;
    int64_t nur = 0;
    int32_t n2A = 0, n2B = 0;
    {
      const int64_t un = ub + stride + lane;
      if (un < NU) {
        nhp = (uint32_t)sorted_pos[nA];
        if (upd) nur = sink.uniq_rows[un];
      }
      const int64_t u2 = un + stride;
      if (u2 < NU) {
        n2A = seg_start[u2];
        n2B = seg_start[u2 + 1];
      }
    }
#pragma unroll 1
    for (int t0 = 0; t0 < LPR; t0 += UB) {
      int32_t s[UB], e[UB];
      uint32_t p0[UB];
      int64_t urow[UB];
      bool live[UB];
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        const int sl = giw + GPW * (t0 + j);  // lane that holds this segment's indices
        s[j] = __shfl_sync(FULL, sA, sl);
        e[j] = __shfl_sync(FULL, sB, sl);
        p0[j] = __shfl_sync(FULL, hp, sl);
        urow[j] = __shfl_sync(FULL, ur, sl);
        live[j] = col && (ub + sl < NU);
        if (lw.counters && e[j] - s[j] > SEG_LONG && ub + sl < NU) {  // long segment: hand it to the chunked path
          if (lir == 0) {
            const int nc = (e[j] - s[j] + SEG_CHUNK - 1) / SEG_CHUNK;
            const int li = atomicAdd(lw.counters, 1);
            const int base = atomicAdd(lw.counters + 1, nc);
            lw.long_u[li] = (int32_t)(ub + sl);
            lw.long_base[li] = base;
            for (int t = 0; t < nc; ++t) lw.chunk_long[base + t] = li;
          }
          live[j] = false;
        }
      }
      float4 acc[UB], tv[UB];
      float ba[UB], la[UB], bo[UB], lo[UB];
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        tv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        ba[j] = la[j] = bo[j] = lo[j] = 0.f;
        if (live[j]) {
          src.load(p0[j], c, acc[j], ba[j], la[j]);
          if (upd) {  // the parameter row to update is loaded beside its first gradient row
            tv[j] = ld4(sink.table + urow[j] * k + 4 * c);
            if (FUSED && c == 0) {
              if (sink.bias_table) bo[j] = sink.bias_table[urow[j]];
              if (sink.lin_table) lo[j] = sink.lin_table[urow[j]];
            }
          }
        }
      }
      // longer segments: U rows in flight, added in ascending position order
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        if (live[j]) {
          for (int32_t j0 = s[j] + 1; j0 < e[j]; j0 += U) {
            float4 v[U];
            float gf[U], gl[U];
#pragma unroll
            for (int i = 0; i < U; ++i) {
              v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              gf[i] = 0.f;
              gl[i] = 0.f;
              if (j0 + i < e[j]) src.load((uint32_t)sorted_pos[j0 + i], c, v[i], gf[i], gl[i]);
            }
#pragma unroll
            for (int i = 0; i < U; ++i) {
              if (j0 + i < e[j]) {
                acc[j].x += v[i].x; acc[j].y += v[i].y; acc[j].z += v[i].z; acc[j].w += v[i].w;
                ba[j] += gf[i];
                la[j] += gl[i];
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < UB; ++j) {
        if (live[j]) {
          const int64_t u = ub + giw + GPW * (t0 + j);
          finish_row<FUSED>(u, c, k, acc[j], ba[j], la[j], out_rows, out_bias, out_lin, sink, urow[j], tv[j], bo[j],
                            lo[j]);
        }
      }
    }
    sA = nA; sB = nB; hp = nhp; ur = nur;
    nA = n2A; nB = n2B;
  }
}

// long segments, pass 2: one row group per chunk slot sums its SEG_CHUNK positions in ascending order
template <int LPR, class Src>
__global__ void __launch_bounds__(256) segment_chunk_kernel(const Src src, int k,
                                                            const int32_t* __restrict__ sorted_pos,
                                                            const int32_t* __restrict__ seg_start, const LongWs lw) {
  constexpr int U = 4;
  const int lir = threadIdx.x % LPR;
  const int c = lir;
  const bool col = c < (k >> 2);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  const int32_t n_chunks = lw.counters[1];
  for (int64_t slot = group; slot < n_chunks; slot += n_groups) {
    const int li = lw.chunk_long[slot];
    const int u = lw.long_u[li];
    const int jc = (int)slot - lw.long_base[li];
    const int32_t s = seg_start[u] + jc * SEG_CHUNK;
    const int32_t e_seg = seg_start[u + 1];
    const int32_t e = s + SEG_CHUNK < e_seg ? s + SEG_CHUNK : e_seg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float ba = 0.f, la = 0.f;
    if (col) {
      for (int32_t j0 = s; j0 < e; j0 += U) {
        float4 v[U];
        float gf[U], gl[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          gf[i] = 0.f;
          gl[i] = 0.f;
          if (j0 + i < e) src.load((uint32_t)sorted_pos[j0 + i], c, v[i], gf[i], gl[i]);
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
          if (j0 + i < e) {
            acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w;
            ba += gf[i];
            la += gl[i];
          }
        }
      }
      st4(lw.part_rows + slot * k + 4 * c, acc);
      if (c == 0) *reinterpret_cast<float2*>(lw.part_scal + 2 * slot) = make_float2(ba, la);
    }
  }
}

// long segments, pass 3: one row group per long segment adds its chunk partials in chunk order and finishes the row
template <int LPR, class Src>
__global__ void __launch_bounds__(256) segment_combine_kernel(int k, const int32_t* __restrict__ seg_start,
                                                              float* __restrict__ out_rows,
                                                              float* __restrict__ out_bias,
                                                              float* __restrict__ out_lin, const UpdateSink sink,
                                                              const LongWs lw) {
  constexpr bool FUSED = Src::kScalars;
  const int lir = threadIdx.x % LPR;
  const int c = lir;
  const bool col = c < (k >> 2);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  const int32_t n_long = lw.counters[0];
  for (int64_t li = group; li < n_long; li += n_groups) {
    const int u = lw.long_u[li];
    const int base = lw.long_base[li];
    const int nc = (seg_start[u + 1] - seg_start[u] + SEG_CHUNK - 1) / SEG_CHUNK;
    if (!col) continue;
    int64_t urow = 0;
    float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
    float bo = 0.f, lo = 0.f;
    if (sink.table) {
      urow = sink.uniq_rows[u];
      tv = ld4(sink.table + urow * k + 4 * c);
      if (FUSED && c == 0) {
        if (sink.bias_table) bo = sink.bias_table[urow];
        if (sink.lin_table) lo = sink.lin_table[urow];
      }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float ba = 0.f, la = 0.f;
    for (int t = 0; t < nc; ++t) {
      const float4 v = ld4(lw.part_rows + (int64_t)(base + t) * k + 4 * c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      if (FUSED && c == 0) {
        const float2 sc = *reinterpret_cast<const float2*>(lw.part_scal + 2 * (int64_t)(base + t));
        ba += sc.x;
        la += sc.y;
      }
    }
    finish_row<FUSED>(u, c, k, acc, ba, la, out_rows, out_bias, out_lin, sink, urow, tv, bo, lo);
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
    const int32_t s = seg_start[u], e = seg_start[u + 1];
    auto at = [&](int32_t j) {
      const uint32_t p = (uint32_t)sorted_pos[j];
      const uint32_t b = p / m, f = p - b * m;
      return grad[(int64_t)b * ld + (int64_t)f * k + c];
    };
    float acc = 0.f;
    if (e - s <= SEG_LONG) {
      for (int32_t j = s; j < e; ++j) acc += at(j);
    } else {  // same association as the chunked long-segment path of the vector kernels
      for (int32_t c0 = s; c0 < e; c0 += SEG_CHUNK) {
        float part = 0.f;
        const int32_t ce = c0 + SEG_CHUNK < e ? c0 + SEG_CHUNK : e;
        for (int32_t j = c0; j < ce; ++j) part += at(j);
        acc += part;
      }
    }
    out_rows[i] = acc;
  }
}

static inline int pow2ceil__(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int LPR, class Src>
static int launch_segment_reduce(const Src& src, int k, int64_t N, const int32_t* sorted_pos, const int32_t* seg_start,
                                 const int32_t* n_unique, float* out_rows, float* out_bias, float* out_lin,
                                 const UpdateSink& sink, int dflt_variant, const LongWs& lw, cudaStream_t st) {
  const int variant = dflt_variant;  // per entry point: the measured best on B200 (profiles/r1_call19_kbench_c5shape_10Mrows.json)
  // n_unique <= N lives on the device: the grid is sized for the worst case (a CTA iteration = 8 warps x 32 rows)
  const int grid = grid_for(N, 256, 8);
  if (lw.counters) RM_CUDA(cudaMemsetAsync(lw.counters, 0, 2 * sizeof(int32_t), st));
#define RM_SRK(UB, MINB)                                                                                         \
  segment_reduce_kernel<LPR, UB, Src::kTailU, MINB, Src><<<grid, 256, 0, st>>>(src, k, sorted_pos, seg_start, n_unique, \
                                                                     out_rows, out_bias, out_lin, sink, lw)
  if (variant == 1) RM_SRK(1, 4);
  else if (variant == 2) RM_SRK(2, 4);
  else if (variant == 3) RM_SRK(2, 3);
  else RM_SRK(4, 2);
#undef RM_SRK
  RM_LAUNCH_CHECK();
  if (lw.counters) {
    // worst-case grids (the counts live on the device); both kernels exit at once when nothing was queued
    segment_chunk_kernel<LPR, Src><<<grid_for(lw.max_chunks, 256 / LPR, 8), 256, 0, st>>>(src, k, sorted_pos, seg_start, lw);
    RM_LAUNCH_CHECK();
    segment_combine_kernel<LPR, Src><<<grid_for(lw.max_long, 256 / LPR, 8), 256, 0, st>>>(k, seg_start, out_rows, out_bias,
                                                                                      out_lin, sink, lw);
    RM_LAUNCH_CHECK();
  }
  return 0;
}

template <class Src>
static int dispatch_segment_reduce(const Src& src, int k, int64_t N, const int32_t* sorted_pos,
                                   const int32_t* seg_start, const int32_t* n_unique, float* out_rows,
                                   float* out_bias, float* out_lin, const UpdateSink& sink, int dflt_variant,
                                   const LongWs& lw, cudaStream_t st) {
  int lpr = pow2ceil__(k / 4);
  if (lpr > 32) lpr = 32;
#define RM_SR(L)                                                                                              \
  case L:                                                                                                     \
    return launch_segment_reduce<L, Src>(src, k, N, sorted_pos, seg_start, n_unique, out_rows, out_bias,       \
                                         out_lin, sink, dflt_variant, lw, st)
  switch (lpr) {
    RM_SR(1);
    RM_SR(2);
    RM_SR(4);
    RM_SR(8);
    RM_SR(16);
    default:
      return launch_segment_reduce<32, Src>(src, k, N, sorted_pos, seg_start, n_unique, out_rows, out_bias, out_lin,
                                            sink, dflt_variant, lw, st);
  }
#undef RM_SR
}

// ---------------------------------------------------------------------------
// owner-side plan for row-sharded tables (row r of a table lives on rank r mod W at local row r div W, W = 2^s):
// from the ids of ALL ranks (gids [W*b, m], rank-major) select the entries this rank owns, in ascending global
// position, key them by owner-local row and sort.  The owned count lives on the device (no host sync); the sort runs
// over a fixed capacity N_cap with sentinel keys behind the live entries.
// ---------------------------------------------------------------------------
struct OwnedBy {
  const int64_t* gids;
  const int64_t* feat_sizes;
  uint32_t m;
  int W, wshift, rank;
  __host__ __device__ __forceinline__ bool operator()(const int32_t& i) const {
    const int64_t id = gids[i];
    if (!(id >= 0 && id < feat_sizes[(uint32_t)i % m])) return false;
    int owner;
    int64_t lr;
    shard_of(id, W, wshift, owner, lr);
    return owner == rank;
  }
};

__global__ void __launch_bounds__(256) shard_keys_kernel(const int64_t* __restrict__ gids,
                                                         const int64_t* __restrict__ local_offs, uint32_t m, int W,
                                                         int wshift, const int32_t* __restrict__ own_gpos,
                                                         const int32_t* __restrict__ n_own, int32_t N_cap,
                                                         uint32_t sentinel, uint32_t* __restrict__ keys,
                                                         int32_t* __restrict__ pos, int32_t* status) {
  const int32_t n_all = *n_own;
  const int32_t n = n_all < N_cap ? n_all : N_cap;
  if (n_all > N_cap && blockIdx.x == 0 && threadIdx.x == 0 && status) atomicOr(status, 4);  // capacity exceeded
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N_cap; i += gridDim.x * blockDim.x) {
    if (i < n) {
      const int32_t gp = own_gpos[i];
      int owner;
      int64_t lr;
      shard_of(gids[gp], W, wshift, owner, lr);
      keys[i] = (uint32_t)(local_offs[(uint32_t)gp % m] + lr);
      pos[i] = gp;
    } else {
      keys[i] = sentinel;
      pos[i] = 0;
    }
  }
}

struct HeadOfLiveRun {
  const uint32_t* keys;
  const int32_t* n_own;
  int32_t N_cap;
  __device__ __forceinline__ bool operator()(const int32_t& i) const {
    const int32_t n_all = *n_own;
    const int32_t n = n_all < N_cap ? n_all : N_cap;
    return i < n && (i == 0 || keys[i] != keys[i - 1]);
  }
};

__global__ void __launch_bounds__(256) finish_shard_plan_kernel(const uint32_t* __restrict__ sorted_keys,
                                                                int32_t* __restrict__ seg_start,
                                                                const int32_t* __restrict__ n_unique,
                                                                const int32_t* __restrict__ n_own, int32_t N_cap,
                                                                int64_t* __restrict__ uniq_rows) {
  const int32_t U = *n_unique;
  const int32_t n_all = *n_own;
  const int32_t n = n_all < N_cap ? n_all : N_cap;
  for (int32_t u = blockIdx.x * blockDim.x + threadIdx.x; u <= U; u += gridDim.x * blockDim.x) {
    if (u == U)
      seg_start[U] = n;
    else
      uniq_rows[u] = (int64_t)sorted_keys[seg_start[u]];
  }
}

struct ShardPlanWs {
  int32_t* own_gpos;
  uint32_t* keys_in;
  uint32_t* keys_out;
  int32_t* pos_in;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static ShardPlanWs shard_plan_layout(int64_t Ntot, int64_t N_cap, void* base) {
  ShardPlanWs w;
  size_t sort_bytes = 0, sel1 = 0, sel2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)N_cap, 0, 32);
  thrust::counting_iterator<int32_t> counting(0);
  OwnedBy own{nullptr, nullptr, 1, 1, 0, 0};
  cub::DeviceSelect::If(nullptr, sel1, counting, (int32_t*)nullptr, (int32_t*)nullptr, (int)Ntot, own);
  HeadOfLiveRun head{nullptr, nullptr, 0};
  cub::DeviceSelect::If(nullptr, sel2, counting, (int32_t*)nullptr, (int32_t*)nullptr, (int)N_cap, head);
  w.cub_bytes = sort_bytes;
  if (sel1 > w.cub_bytes) w.cub_bytes = sel1;
  if (sel2 > w.cub_bytes) w.cub_bytes = sel2;
  size_t off = 0;
  char* b = (char*)base;
  w.own_gpos = (int32_t*)(b + off); off += align_up((size_t)Ntot * 4, 256);
  const size_t arr = align_up((size_t)N_cap * 4, 256);
  w.keys_in = (uint32_t*)(b + off); off += arr;
  w.keys_out = (uint32_t*)(b + off); off += arr;
  w.pos_in = (int32_t*)(b + off); off += arr;
  w.cub_temp = (void*)(b + off); off += align_up(w.cub_bytes, 256);
  w.total = off;
  return w;
}

}  // namespace rm

extern "C" {

// long-segment workspace of the reduce entry points (NULL workspace = long path disabled: every segment is walked
// serially, fine for near-uniform ids)
static int long_ws_from(void* workspace, size_t workspace_bytes, int64_t N, int k, const char* who, rm::LongWs* lw) {
  using namespace rm;
  *lw = LongWs{};
  if (!workspace) return 0;
  size_t total = 0;
  *lw = long_layout(N, k, workspace, &total);
  if (workspace_bytes < total) {
    set_error("%s: workspace %zu < required %zu", who, workspace_bytes, total);
    return RM_E_WORKSPACE;
  }
  if (!aligned16(workspace)) {
    set_error("%s: workspace must be 16-byte aligned", who);
    return RM_E_INVALID;
  }
  return 0;
}

size_t rm_segment_reduce_workspace_bytes(int64_t N, int32_t k) {
  if (N <= 0 || k <= 0) return 256;
  size_t total = 0;
  rm::long_layout(N, k, nullptr, &total);
  return total;
}

size_t rm_segment_plan_workspace_bytes(int64_t N) {
  if (N <= 0) return 256;
  return rm::plan_layout(N, nullptr).total;
}

int rm_segment_plan(const int64_t* ids, const int64_t* table_offsets, int64_t N, int32_t m, int64_t total_rows,
                    void* workspace, size_t workspace_bytes, int32_t* sorted_pos, int32_t* seg_start,
                    int64_t* uniq_rows, int32_t* n_unique, int32_t* status, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(seg_start && n_unique, "null pointer");
  RM_CHECK_ARG(N >= 0 && m > 0 && total_rows > 0, "bad shape");
  RM_CHECK_ARG(N == 0 || (ids && sorted_pos && uniq_rows), "null pointer");
  RM_UNSUPPORTED(N < ((int64_t)1 << 31) - 1, "N must be < 2^31 - 1");
  RM_UNSUPPORTED(total_rows < ((int64_t)1 << 32) - 1, "total_rows must be < 2^32 - 1 (32-bit sort keys + sentinel)");
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
  make_keys_kernel<<<grid_for(N, 256, 8), 256, 0, st>>>(ids, table_offsets, (uint32_t)N, (uint32_t)m, total_rows,
                                                        w.keys_in, w.pos_in, status);
  RM_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 32 && ((int64_t)1 << end_bit) <= total_rows) ++end_bit;  // the sentinel key total_rows sorts too
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
  drop_sentinel_kernel<<<1, 1, 0, st>>>(w.keys_out, seg_start, n_unique, total_rows);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_segment_reduce(const float* grad, int64_t ld, int32_t m, int32_t k, int64_t N, const int32_t* sorted_pos,
                      const int32_t* seg_start, const int32_t* n_unique, float* out_rows, void* workspace,
                      size_t workspace_bytes, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(grad && sorted_pos && seg_start && n_unique && out_rows, "null pointer");
  RM_CHECK_ARG(N >= 0 && m > 0 && k > 0 && ld >= (int64_t)m * k, "bad shape");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (k % 4 == 0) && (k <= 128) && (ld % 4 == 0) && aligned16(grad) && aligned16(out_rows);
  if (vec) {
    const RowSrc<false> src{grad, nullptr, nullptr, nullptr, nullptr, ld, (uint32_t)m, k};
    LongWs lw;
    const int rc = long_ws_from(workspace, workspace_bytes, N, k, "rm_segment_reduce", &lw);
    if (rc) return rc;
    return dispatch_segment_reduce(src, k, N, sorted_pos, seg_start, n_unique, out_rows, nullptr, nullptr, UpdateSink{}, 2,
                                   lw, st);
  }
  segment_reduce_scalar_kernel<<<grid_for(N * k, 256, 8), 256, 0, st>>>(grad, ld, (uint32_t)m, (uint32_t)k, sorted_pos,
                                                                        seg_start, n_unique, out_rows);
  RM_LAUNCH_CHECK();
  return 0;
}

static int emb_fm_bwd_impl(const float* dx, const float* x, int64_t ld, const float* sum, const float* g_fm,
                           const float* g_lin, int32_t m, int32_t k, int64_t N, const int32_t* sorted_pos,
                           const int32_t* seg_start, const int32_t* n_unique, float* out_rows, float* out_bias,
                           float* out_lin, const rm::UpdateSink& sink, void* workspace, size_t workspace_bytes,
                           void* stream) {
  using namespace rm;
  RM_CHECK_ARG(sorted_pos && seg_start && n_unique, "null pointer");
  RM_CHECK_ARG(N >= 0 && m > 0 && k > 0 && ld >= (int64_t)m * k, "bad shape");
  RM_CHECK_ARG(!g_fm || (x && sum), "g_fm needs x and sum");
  RM_UNSUPPORTED((k % 4 == 0) && (k <= 128) && (ld % 4 == 0) && (!dx || aligned16(dx)) && (!x || aligned16(x)) &&
                     (!sum || aligned16(sum)) && (!out_rows || aligned16(out_rows)) &&
                     (!sink.table || aligned16(sink.table)),
                 "fused embedding backward needs k % 4 == 0, k <= 128 and 16-byte aligned rows");
  if (N == 0) return 0;
  const RowSrc<true> src{dx, x, sum, g_fm, g_lin, ld, (uint32_t)m, k};
  LongWs lw;
  const int rc = long_ws_from(workspace, workspace_bytes, N, k, "rm_emb_fm_bwd", &lw);
  if (rc) return rc;
  return dispatch_segment_reduce(src, k, N, sorted_pos, seg_start, n_unique, out_rows, out_bias, out_lin, sink, 1, lw,
                                 (cudaStream_t)stream);
}

int rm_emb_fm_bwd(const float* dx, const float* x, int64_t ld, const float* sum, const float* g_fm, const float* g_lin,
                  int32_t m, int32_t k, int64_t N, const int32_t* sorted_pos, const int32_t* seg_start,
                  const int32_t* n_unique, float* out_rows, float* out_bias, float* out_lin, void* workspace,
                  size_t workspace_bytes, void* stream) {
  return emb_fm_bwd_impl(dx, x, ld, sum, g_fm, g_lin, m, k, N, sorted_pos, seg_start, n_unique, out_rows, out_bias,
                         out_lin, rm::UpdateSink{}, workspace, workspace_bytes, stream);
}

int rm_emb_fm_bwd_update(const float* dx, const float* x, int64_t ld, const float* sum, const float* g_fm,
                         const float* g_lin, int32_t m, int32_t k, int64_t N, const int32_t* sorted_pos,
                         const int32_t* seg_start, const int64_t* uniq_rows, const int32_t* n_unique, float* table,
                         float* bias_table, float* lin_table, int32_t opt, float lr, float l2, void* workspace,
                         size_t workspace_bytes, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(table && uniq_rows, "null pointer");
  UpdateSink sink{table, g_fm ? bias_table : nullptr, g_lin ? lin_table : nullptr, uniq_rows, OptParams{}};
  const int rc = make_params(opt, lr, l2, &sink.o);
  if (rc) return rc;
  return emb_fm_bwd_impl(dx, x, ld, sum, g_fm, g_lin, m, k, N, sorted_pos, seg_start, n_unique, nullptr, nullptr,
                         nullptr, sink, workspace, workspace_bytes, stream);
}

size_t rm_shard_plan_workspace_bytes(int64_t Ntot, int64_t N_cap) {
  if (Ntot <= 0 || N_cap <= 0) return 256;
  return rm::shard_plan_layout(Ntot, N_cap, nullptr).total;
}

int rm_shard_plan(const int64_t* gids, int64_t Ntot, int32_t m, int32_t W, int32_t rank, const int64_t* feat_sizes,
                  const int64_t* local_offsets, int64_t total_local, int64_t N_cap, void* workspace,
                  size_t workspace_bytes, int32_t* sorted_gpos, int32_t* seg_start, int64_t* uniq_rows,
                  int32_t* n_unique, int32_t* n_own, int32_t* status, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(gids && feat_sizes && local_offsets && workspace && sorted_gpos && seg_start && uniq_rows && n_unique &&
                   n_own, "null pointer");
  RM_CHECK_ARG(Ntot > 0 && m > 0 && N_cap > 0 && total_local > 0 && rank >= 0 && rank < W, "bad shape");
  RM_UNSUPPORTED(W >= 1 && W <= RM_MAX_PEERS && true, "world size must be <= 8");
  RM_UNSUPPORTED(Ntot < ((int64_t)1 << 31) - 1 && N_cap <= Ntot, "W*B*m must be < 2^31 - 1 and N_cap <= W*B*m");
  RM_UNSUPPORTED(total_local < ((int64_t)1 << 31), "local rows must be < 2^31 (32-bit sort keys + sentinel)");
  cudaStream_t st = (cudaStream_t)stream;
  ShardPlanWs w = shard_plan_layout(Ntot, N_cap, workspace);
  if (workspace_bytes < w.total) {
    set_error("rm_shard_plan: workspace %zu < required %zu", workspace_bytes, w.total);
    return RM_E_WORKSPACE;
  }
  const int wshift = world_shift(W);
  thrust::counting_iterator<int32_t> counting(0);
  OwnedBy own{gids, feat_sizes, (uint32_t)m, (int)W, wshift, (int)rank};
  size_t bytes = w.cub_bytes;
  RM_CUDA(cub::DeviceSelect::If(w.cub_temp, bytes, counting, w.own_gpos, n_own, (int)Ntot, own, st));
  count_launch();
  int end_bit = 1;
  while (end_bit < 31 && ((int64_t)1 << end_bit) < total_local) ++end_bit;
  const uint32_t sentinel = 1u << end_bit;
  shard_keys_kernel<<<grid_for(N_cap, 256, 8), 256, 0, st>>>(gids, local_offsets, (uint32_t)m, (int)W, wshift, w.own_gpos, n_own,
                                                            (int32_t)N_cap, sentinel, w.keys_in, w.pos_in, status);
  RM_LAUNCH_CHECK();
  bytes = w.cub_bytes;
  RM_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, bytes, (const uint32_t*)w.keys_in, w.keys_out,
                                          (const int32_t*)w.pos_in, sorted_gpos, (int)N_cap, 0, end_bit + 1, st));
  count_launch();
  HeadOfLiveRun head{w.keys_out, n_own, (int32_t)N_cap};
  bytes = w.cub_bytes;
  RM_CUDA(cub::DeviceSelect::If(w.cub_temp, bytes, counting, seg_start, n_unique, (int)N_cap, head, st));
  count_launch();
  finish_shard_plan_kernel<<<grid_for(N_cap + 1, 256, 8), 256, 0, st>>>(w.keys_out, seg_start, n_unique, n_own,
                                                                       (int32_t)N_cap, uniq_rows);
  RM_LAUNCH_CHECK();
  return 0;
}

static int segment_reduce_p2p_impl(const float* const* G, const float* gscal, int32_t m, int32_t W,
                                   int64_t rows_per_rank, int32_t KP, int32_t k, int64_t N_cap, const int32_t* sorted_gpos, const int32_t* seg_start,
                                   const int32_t* n_unique, float* out_rows, float* out_bias, float* out_lin,
                                   const rm::UpdateSink& sink, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  using namespace rm;
  RM_CHECK_ARG(G && sorted_gpos && seg_start && n_unique, "null pointer");
  RM_CHECK_ARG(W >= 1 && W <= RM_MAX_PEERS && rows_per_rank > 0 && k > 0 && N_cap >= 0, "bad shape");
  RM_CHECK_ARG(gscal ? (KP >= k && m > 0 && rows_per_rank % m == 0) : KP >= k + 4,
               "rows need their two k=1 columns (KP >= k + 4) unless the per-sample scalars are passed");
  RM_UNSUPPORTED(k % 4 == 0 && k <= 128 && KP % 4 == 0 && (!out_rows || aligned16(out_rows)) &&
                     (!sink.table || aligned16(sink.table)),
                 "peer segment reduce needs k % 4 == 0 and 16-byte aligned rows");
  RM_UNSUPPORTED(rows_per_rank * W < ((int64_t)1 << 31), "W * rows_per_rank must be < 2^31");
  if (N_cap == 0) return 0;
  PeerRowSrc src;
  for (int r = 0; r < RM_MAX_PEERS; ++r) {
    src.G[r] = r < W ? G[r] : nullptr;
    RM_CHECK_ARG(r >= W || (G[r] && aligned16(G[r])), "null / misaligned peer buffer");
  }
  src.rows_per_rank = (uint32_t)rows_per_rank;
  src.gscal = gscal;
  src.m = (uint32_t)(m > 0 ? m : 1);
  src.KP = KP;
  src.k = k;
  LongWs lw;
  const int rcw = long_ws_from(workspace, workspace_bytes, N_cap, k, "rm_segment_reduce_p2p", &lw);
  if (rcw) return rcw;
  return dispatch_segment_reduce(src, k, N_cap, sorted_gpos, seg_start, n_unique, out_rows, out_bias, out_lin, sink, 1,
                                 lw, (cudaStream_t)stream);
}

int rm_segment_reduce_p2p(const float* const* G, const float* gscal, int32_t m, int32_t W, int64_t rows_per_rank,
                          int32_t KP, int32_t k, int64_t N_cap, const int32_t* sorted_gpos, const int32_t* seg_start,
                          const int32_t* n_unique, float* out_rows, float* out_bias, float* out_lin, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return segment_reduce_p2p_impl(G, gscal, m, W, rows_per_rank, KP, k, N_cap, sorted_gpos, seg_start, n_unique,
                                 out_rows, out_bias, out_lin, rm::UpdateSink{}, workspace, workspace_bytes, stream);
}

int rm_segment_reduce_p2p_update(const float* const* G, const float* gscal, int32_t m, int32_t W,
                                 int64_t rows_per_rank, int32_t KP, int32_t k, int64_t N_cap, const int32_t* sorted_gpos, const int32_t* seg_start,
                                 const int64_t* uniq_rows, const int32_t* n_unique, float* table, float* bias_table,
                                 float* lin_table, int32_t opt, float lr, float l2, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(table && uniq_rows, "null pointer");
  UpdateSink sink{table, bias_table, lin_table, uniq_rows, OptParams{}};
  const int rc = make_params(opt, lr, l2, &sink.o);
  if (rc) return rc;
  return segment_reduce_p2p_impl(G, gscal, m, W, rows_per_rank, KP, k, N_cap, sorted_gpos, seg_start, n_unique,
                                 nullptr, nullptr, nullptr, sink, workspace, workspace_bytes, stream);
}

}  // extern "C"
