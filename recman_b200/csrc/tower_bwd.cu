// Fused DeepFM "tower" backward + optimizer, one pass over the touched embedding rows in SORTED order:
//
//   dx    = g1 @ W1^T                         backward of the first DNN matmul, recman/tf/core/layers.py:589-609
//   grad  = dx + g_fm * (S - x)               backward of FMLayer, layers.py:457-478
//   row   = sum of grad over the positions    autodiff of tf.nn.embedding_lookup (IndexedSlices, duplicates summed),
//           that looked the row up            layers.py:117-128 - deterministic: ascending-position order, no atomics
//   table[row] <- fresh-optimizer step        xDeepFM.py:116-126 (a new optimizer every batch)
//   dW1  += x^T @ g1                          weight gradient of the first DNN matmul
//
// A tile is 128 consecutive sorted positions of ONE field.  Its table rows x (256 B each, HBM) and its g1 rows
// (128 B, L2) are gathered into swizzled shared-memory tiles (cp.async for the rows); the tiles are the K-major A operand of
// GEMM 1 (dx tile = g1 tile @ W1_f^T, M = 128 positions) and the MN-major A / B operands of GEMM 2
// (dW1_f += x tile^T @ g1 tile, K = 128 positions) - nothing is transposed, nothing is written back.  Both GEMMs run
// on tcgen05 in 3xTF32 (operand = trunc + exact remainder; W1 pre-split round-to-nearest).  The epilogue owns one
// position per thread pair: dx row from TMEM, FM term, segment sum over equal rows (carried across tiles), optimizer
// update, 128-byte stores of the new row.  A table row is read once and written once per step; dx, the row buffer and
// the summed gradients never exist in HBM.
//
// Work units: a field's sorted range is cut every `unit` positions, the cut moved forward to the next segment head
// (tower_bounds_kernel), so a segment never straddles two units; unit u writes its dW1 slab [64, 32] and
// tower_dw_reduce_kernel adds the slabs in unit order (deterministic, independent of the grid).  The TMEM
// accumulation of GEMM 2 truncates, so it is drained into fp32 registers every 4 tiles (512 positions).
//
// Hot rows (skewed ids): the kernel is compiled in two variants.  In the hot-row variant a run of more than BK_WALK
// positions on one row is summed by the leader's whole warp (one column per lane, same ascending order: bit-identical
// to the leader's own walk, ~15x fewer cycles per position).  The plan leaves a device-side flag behind the unit cuts
// (tower_hot_flag_kernel: some row holds > 32 positions); both variants are launched and the one the flag does not
// select returns at once, or the caller names the variant (rm_tower_bwd_update `variant`).
//
// Roofline: HBM.  Algorithmic bytes: per position 8 (key, position) + 4k (row read), per unique row 4k + 8 + 8
// written / read-modify-written.  Measured (C5): 0.71 ms = 0.20 of the HBM peak - bound by the per-tile latency chain
// at 3 gather stages, not by bandwidth.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "optim.cuh"
#include "tower_common.cuh"

namespace rm {

constexpr int BK_K = 64;    // embedding size
constexpr int BK_N1 = 32;   // first hidden layer width
constexpr int BK_TILE = 128;
constexpr int BK_NS = 3;    // gather stages
constexpr int BK_DEPTH = 2; // tiles in flight per producer thread
constexpr int BK_PRODUCERS = 128;
// Epilogue variants that were measured and lost at the C5 shape (kept switchable for other shapes):
//   BK_FASTPATH: tiles of singleton segments release their gather stage before the update math and store rows straight
//                from registers - 0.80 ms against 0.72 ms with the staged, coalesced write-back;
//   BK_PREFETCH: next tile's S rows requested one tile ahead - the 32 extra live registers spill (0.88 ms);
//   BK_NH = 4  : four threads per position - per-position bookkeeping replicated, 0.78 .. 1.06 ms;
//   (not kept) a register-free `prefetch.global.L1` of the next tile's S rows: 0.714 vs 0.705 ms - the ~28 KB of L1 left
//   beside 227 KB of shared memory do not hold a tile's 32 KB of S rows.
constexpr bool BK_FASTPATH = false;
#ifndef RM_BK_PREFETCH
#define RM_BK_PREFETCH 0
#endif
#ifndef RM_BK_NH
#define RM_BK_NH 2
#endif
constexpr bool BK_PREFETCH = RM_BK_PREFETCH != 0;  // request the next tile's per-sample operands one tile ahead
constexpr int BK_NH = RM_BK_NH;              // epilogue threads per position (each owns 64 / BK_NH columns of the row)
constexpr int BK_EPI = 128 * BK_NH;
constexpr int BK_THREADS = 128 + 32 + BK_EPI;  // 4 producer warps, 1 MMA warp, 4 * BK_NH epilogue warps
constexpr int BK_DRAIN = 4;  // tiles per TMEM accumulation group of GEMM 2
constexpr int BK_WALK = 6;   // positions a segment leader adds by itself before the warp takes the run over
constexpr uint32_t BK_XT = 32768, BK_GT = 16384, BK_STAGE = BK_XT;  // a gather stage holds the x tile only
constexpr uint32_t BK_META = 1088;  // key[130] (prev, 128, next) + b[128], padded

// ------------------------------------------------------------------------------------------------ plan
__global__ void __launch_bounds__(256) tower_keys_kernel(const int64_t* __restrict__ ids,
                                                         const int64_t* __restrict__ offs, uint32_t N, uint32_t m,
                                                         uint32_t sentinel, uint32_t* __restrict__ keys,
                                                         int32_t* __restrict__ pos, int32_t* status) {
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
    const uint32_t f = p % m;
    const int64_t id = ids[p];
    const int64_t lo = offs[f], hi = offs[f + 1];
    const bool ok = id >= 0 && id < hi - lo;
    if (!ok && status) atomicOr(status, 1);
    keys[p] = ok ? (uint32_t)(lo + id) : sentinel;  // ids outside their table sort behind every row and are dropped
    pos[p] = (int32_t)p;
  }
}

__device__ __forceinline__ int32_t lower_bound_u32(const uint32_t* a, int32_t lo, int32_t hi, uint32_t v) {
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// field_bounds[f] = first sorted position of field f (field_bounds[m] = number of valid positions);
// unit_bounds[f*(upf+1) + i] = the i-th cut of field f moved forward to the next segment head.
// *flag = 1 when some table row collects more than BK_HOT_RUN positions of the batch (skewed ids: "hot rows"), found by
// comparing sorted keys BK_HOT_RUN apart; the backward then runs its variant with warp-cooperative run summation.
// Both variants add a run's positions in the same ascending order: the flag selects speed, never results.
constexpr int BK_HOT_RUN = 32;
// launched behind the bounds kernel (which zeroes the flag): n_valid = field_bounds[m], the entries before the sentinels
__global__ void __launch_bounds__(256) tower_hot_flag_kernel(const uint32_t* __restrict__ keys,
                                                             const int32_t* __restrict__ field_bounds, int m,
                                                             int32_t* __restrict__ flag) {
  const int32_t n_valid = field_bounds[m];
  bool hot = false;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * BK_HOT_RUN; p + BK_HOT_RUN < n_valid;
       p += (int64_t)gridDim.x * blockDim.x * BK_HOT_RUN)
    hot = hot || keys[p] == keys[p + BK_HOT_RUN];
  if (__any_sync(0xffffffffu, hot) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void __launch_bounds__(1024) tower_bounds_kernel(const uint32_t* __restrict__ keys, int32_t N,
                                                            const int64_t* __restrict__ offs, int m, int unit, int upf,
                                                            int32_t* __restrict__ field_bounds,
                                                            int32_t* __restrict__ unit_bounds) {
  extern __shared__ int32_t fb[];
  for (int f = threadIdx.x; f <= m; f += blockDim.x) {
    fb[f] = lower_bound_u32(keys, 0, N, (uint32_t)offs[f]);
    field_bounds[f] = fb[f];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < m * (upf + 1); idx += blockDim.x) {
    const int f = idx / (upf + 1), i = idx - f * (upf + 1);
    const int32_t lo = fb[f], hi = fb[f + 1];
    int64_t p64 = (int64_t)lo + (int64_t)i * unit;
    int32_t p = (p64 < hi && i < upf) ? (int32_t)p64 : hi;  // the last cut is the field's end, whatever its length
    if (p > lo && p < hi && keys[p] == keys[p - 1]) {
      const uint32_t kv = keys[p];
      p = kv == 0xFFFFFFFFu ? hi : lower_bound_u32(keys, p, hi, kv + 1u);
    }
    unit_bounds[idx] = p;
  }
  if (threadIdx.x == 0) unit_bounds[m * (upf + 1)] = 0;  // hot-row flag: set by tower_hot_flag_kernel
}

// the same with owner-local offsets [m] (no closing entry) and the total passed separately (row-sharded tables)
__global__ void __launch_bounds__(1024) tower_bounds_local_kernel(const uint32_t* __restrict__ keys, int32_t N,
                                                                  const int64_t* __restrict__ offs_m, uint32_t total, int m,
                                                                  int unit, int upf, int32_t* __restrict__ field_bounds,
                                                                  int32_t* __restrict__ unit_bounds) {
  extern __shared__ int32_t fb[];
  for (int f = threadIdx.x; f <= m; f += blockDim.x) {
    fb[f] = lower_bound_u32(keys, 0, N, f < m ? (uint32_t)offs_m[f] : total);
    field_bounds[f] = fb[f];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < m * (upf + 1); idx += blockDim.x) {
    const int f = idx / (upf + 1), i = idx - f * (upf + 1);
    const int32_t lo = fb[f], hi = fb[f + 1];
    int64_t p64 = (int64_t)lo + (int64_t)i * unit;
    int32_t p = (p64 < hi && i < upf) ? (int32_t)p64 : hi;
    if (p > lo && p < hi && keys[p] == keys[p - 1]) p = lower_bound_u32(keys, p, hi, keys[p] + 1u);
    unit_bounds[idx] = p;
  }
  if (threadIdx.x == 0) unit_bounds[m * (upf + 1)] = 0;  // hot-row flag: set by tower_hot_flag_kernel
}

// W1 field images for GEMM 1: B operand = W1_f (rows c < 64, K = n < 32), K-major.  Per field [hi|lo][64][128 B].
__global__ void __launch_bounds__(256) tower_pack_w1_kernel(const float* __restrict__ W1, int m,
                                                            uint32_t* __restrict__ out) {
  const int64_t total = (int64_t)m * 2 * BK_K * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i & 7);
    int64_t t = i >> 3;
    const int c = (int)(t % BK_K); t /= BK_K;
    const int part = (int)(t & 1);
    const int f = (int)(t >> 1);
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float w = W1[((int64_t)f * BK_K + c) * BK_N1 + 4 * ch + e];
      const uint32_t hi = f32_to_tf32(w);
      v[e] = part ? f32_to_tf32(w - __uint_as_float(hi)) : hi;
    }
    uint32_t* dst = out + (((int64_t)f * 2 + part) * BK_K + c) * 32 + ((ch ^ (c & 7)) << 2);
    *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void __launch_bounds__(256) tower_dw_reduce_kernel(const float* __restrict__ slabs, int m, int upf,
                                                              float* __restrict__ dW1) {
  const int total = m * BK_K * BK_N1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int f = i / (BK_K * BK_N1), e = i - f * (BK_K * BK_N1);
    float acc = 0.f;
    for (int u = 0; u < upf; ++u) acc += slabs[((int64_t)f * upf + u) * (BK_K * BK_N1) + e];
    dW1[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ main kernel
struct TowerBwdParams {
  float* table;           // [rows, 64], gathered; updated in place when do_update
  float* scal;            // [rows, 2] (bias, linear weight), updated in place (nullable)
  const uint32_t* keys;   // sorted keys
  const int32_t* spos;    // sorted positions p = b*m + f
  const int32_t* ub;      // unit bounds [m][upf+1]
  const float* g1;        // [B, 32] d loss / d y1
  const float* S;         // [B, 64] field sums
  const float* g_fm;      // [B]
  const float* g_lin;     // [B] (nullable)
  const uint32_t* wpack;  // W1 field images
  float* slabs;           // [m*upf, 64, 32]
  float* out_rows;        // optional [N, 64]: summed gradient row at the sorted position that closes its segment
  float* out_scal;        // optional [N, 2]
  int32_t* status;
  OptParams o;
  int m, upf, n_units, do_update;
  int force;  // run whatever the plan's hot-row flag says (the caller launched this variant only)
};

__device__ __forceinline__ void atomicOr_shared(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct TileIter {
  int u, u_end, upf, stride;  // units u, u + stride, ... < u_end (strided: empty units spread evenly over the CTAs)
  const int32_t* ub;
  int32_t rs, re, p0;  // unit range, current tile start
  int f, t, nt;
  __device__ __forceinline__ void load_unit() {
    while (u < u_end) {
      f = u / upf;
      const int i = u - f * upf;
      rs = ub[f * (upf + 1) + i];
      re = ub[f * (upf + 1) + i + 1];
      nt = (re - rs + BK_TILE - 1) / BK_TILE;
      t = 0;
      p0 = rs;
      if (nt > 0) return;
      u += stride;
    }
  }
  __device__ __forceinline__ bool valid() const { return u < u_end; }
  __device__ __forceinline__ int cnt() const { return re - p0 < BK_TILE ? re - p0 : BK_TILE; }
  __device__ __forceinline__ bool last_in_unit() const { return t == nt - 1; }
  __device__ __forceinline__ void next() {
    ++t;
    p0 += BK_TILE;
    if (t >= nt) {
      u += stride;
      load_unit();
    }
  }
};

// Long runs of one table row inside a tile (hot rows: thousands of positions of one batch on one row under skewed ids)
// are summed by the whole warp: lane l owns column 32 h + l of the row (one 128-byte shared-memory row per position:
// conflict-free), positions added in ascending order - bit-identical to the leader thread's walk, ~15x fewer cycles per
// position.  `lm`: lanes whose position leads such a run and has already added the first BK_WALK members into its slot of
// the x tile; on return that slot holds the sum over the run's part inside this tile, jj the first position behind it.
__device__ __noinline__ void bk_sum_long_runs(uint32_t lm, uint32_t xs, uint32_t ms, uint32_t sc_base, int cnt, int h,
                                              int q, int lane, uint32_t key, int& jj, float& asf, float& asl) {
  while (lm) {
    const int src = __ffs(lm) - 1;
    lm &= lm - 1;
    const int j0 = 32 * q + src;
    const uint32_t key0 = __shfl_sync(0xffffffffu, key, src);
    const int jb = __shfl_sync(0xffffffffu, jj, src);  // first position not yet added
    __syncwarp();
    int je = cnt;  // end of the run inside this tile
    for (int base = jb; base < cnt; base += 32) {
      const int pj = base + lane;
      const uint32_t kk = pj < cnt ? lds32(ms + 4u * (pj + 1)) : TW_NONE;
      const uint32_t mism = __ballot_sync(0xffffffffu, kk != key0);
      if (mism) {
        je = base + __ffs(mism) - 1;
        break;
      }
    }
    const uint32_t colb = xs + (uint32_t)h * 16384u + (uint32_t)(lane & 3) * 4u;
    const uint32_t cq = (uint32_t)lane >> 2;
    float acc = __uint_as_float(lds32(colb + (uint32_t)j0 * 128u + sw32b_chunk(cq, (uint32_t)j0)));
    const float sf0 = __shfl_sync(0xffffffffu, asf, src), sl0 = __shfl_sync(0xffffffffu, asl, src);
    float sacc = lane == 0 ? sf0 : sl0;  // lane 0: bias-gradient sum, lane 1: weight-gradient sum (h == 0 only)
#pragma unroll 4
    for (int p2 = jb; p2 < je; ++p2) {
      acc += __uint_as_float(lds32(colb + (uint32_t)p2 * 128u + sw32b_chunk(cq, (uint32_t)p2)));
      if (h == 0 && lane < 2) sacc += __uint_as_float(lds32(sc_base + 8u * p2 + 4u * lane));
    }
    sts32(colb + (uint32_t)j0 * 128u + sw32b_chunk(cq, (uint32_t)j0), __float_as_uint(acc));
    const float sf = __shfl_sync(0xffffffffu, sacc, 0), sl = __shfl_sync(0xffffffffu, sacc, 1);
    __syncwarp();
    if (lane == src) {
      if (h == 0) {
        asf = sf;
        asl = sl;
      }
      jj = je;
    }
  }
}

// COOP: the variant for batches with hot rows (flag behind the unit bounds, tower_hot_flag).  Both variants are launched;
// the one the flag does not select returns at once.  (Compiled into one kernel the cooperative block costs the
// uniform-id path 10 % although it never runs there.)
template <bool COOP>
__global__ void __launch_bounds__(BK_THREADS, 1) tower_bwd_kernel(const TowerBwdParams P) {
  if (!P.force && (P.ub[P.m * (P.upf + 1)] != 0) != COOP) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto x_hi = [&](int s) { return base + (uint32_t)s * BK_STAGE; };
  const uint32_t x_lo = base + BK_NS * BK_STAGE;
  // g1 rows of the tile, four images: K-major SWIZZLE_128B (A of GEMM 1) and MN-major SWIZZLE_128B_BASE32B (B of
  // GEMM 2), each as truncated value ("hi" = the raw word) and exact remainder ("lo")
  const uint32_t g1_hi = x_lo + BK_XT, g1_lo = g1_hi + BK_GT, g2_hi = g1_lo + BK_GT, g2_lo = g2_hi + BK_GT;
  const uint32_t w_img = g2_lo + BK_GT;            // hi 8 KB | lo 8 KB
  const uint32_t meta_base = w_img + 16384u;        // 4 slots
  const uint32_t sc_base0 = meta_base + 4u * BK_META;  // [2][128] float2 (tile parity)
  const uint32_t carry_base = sc_base0 + 2048u;      // [2][64] floats
  const uint32_t carry_sc = carry_base + 512u;      // [2] float2
  const uint32_t scal_st = carry_sc + 16u;          // [BK_NS][128] float2: old (bias, weight) of the tile's rows
  const uint32_t wr_base = scal_st + BK_NS * 1024u; // [2][128] u32: table row to store (TW_NONE = none), tile parity
  const uint32_t bar_base = wr_base + 1024u;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (3 + s); };
  auto d1_full = [&](int b) { return bar_base + 8u * (6 + b); };
  auto d1_empty = [&](int b) { return bar_base + 8u * (8 + b); };
  auto d2_full = [&](int b) { return bar_base + 8u * (10 + b); };
  auto d2_empty = [&](int b) { return bar_base + 8u * (12 + b); };
  const uint32_t lo_free = bar_base + 8u * 14;
  const uint32_t wfull = bar_base + 8u * 15;
  const uint32_t tmem_slot = bar_base + 8u * 16;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < BK_NS; ++s) {
      mbar_init(full(s), BK_PRODUCERS);
      mbar_init(empty(s), BK_EPI);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(d1_full(b), 1);
      mbar_init(d1_empty(b), BK_EPI);
      mbar_init(d2_full(b), 1);
      mbar_init(d2_empty(b), BK_EPI);
    }
    mbar_init(lo_free, 1);
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 4) tmem_alloc_cols(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);
  bool ok = true;
  // (Unit assignment: strided over the CTAs.  Contiguous per-CTA ranges with equal TILE counts were measured under
  // Zipf ids and lost, 1.28 vs 1.07 ms: the tiles of a field's hot head cost 2-3x a singleton tile and a contiguous
  // range hands them all to the same CTAs, while the stride mixes them.)
  // (L2 eviction hints were measured and dropped: evict_first on the row stream costs 2-3 % at every world size, with
  // or without evict_last on the per-sample operands)

  TileIter it;
  it.upf = P.upf;
  it.ub = P.ub;
  it.u = (int)blockIdx.x;
  it.u_end = P.n_units;
  it.stride = (int)gridDim.x;
  const int u_begin = it.u;

  if (warp < 4) {
    // ================================================================== producers: gather + lo tiles
    TileIter it_issue = it, it_cons = it;
    it_issue.load_unit();
    it_cons.load_unit();
    const int cx = tid & 15, rgx = tid >> 4;  // x rows rgx + 8i, chunk cx
    const int cg = tid & 7, rgg = tid >> 3;   // g rows rgg + 16i, chunk cg
    // x tile: MN-major operand of GEMM 2 -> SWIZZLE_128B_BASE32B (two 32-column blocks of [128 rows x 128 B])
    const uint32_t xoff = (uint32_t)(cx >> 3) * 16384u + (uint32_t)rgx * 128u + sw32b_chunk((uint32_t)(cx & 7), (uint32_t)rgx);
    const uint32_t g1off = (uint32_t)rgg * 128u + ((uint32_t)(cg ^ (rgg & 7)) << 4);
    const uint32_t g2off = (uint32_t)rgg * 128u + sw32b_chunk((uint32_t)cg, (uint32_t)rgg);
    uint32_t mkey = TW_NONE, mprev = TW_NONE, mnext = TW_NONE;
    int32_t mb = 0;
    auto load_meta = [&](const TileIter& ti) {
      const int cnt = ti.cnt();
      mkey = TW_NONE;
      mb = 0;
      if (tid < cnt) {
        mkey = P.keys[ti.p0 + tid];
        mb = P.spos[ti.p0 + tid];  // raw position; divided by m when the meta is stored, one tile later, so that the
      }                             // load's latency (keys / positions stream from HBM) is not exposed here
      if (tid == 0) mprev = ti.p0 > ti.rs ? P.keys[ti.p0 - 1] : TW_NONE;
      if (tid == 1) mnext = ti.p0 + cnt < ti.re ? P.keys[ti.p0 + cnt] : TW_NONE;
    };
    int X = 0;  // tiles issued
    auto issue = [&]() {
      const int s = X % BK_NS;
      const uint32_t ms = meta_base + (uint32_t)(X & 3) * BK_META;
      ok = ok && mbar_wait(empty(s), (((uint32_t)(X / BK_NS)) & 1u) ^ 1u);
      const int cnt = it_issue.cnt();
      const uint32_t mkey_cur = mkey;
      // key slots: [0] previous key, [1..cnt] the tile, [cnt+1] next key, the rest TW_NONE
      sts32(ms + 4u * (tid < cnt ? tid + 1 : tid + 2), mkey);
      sts32(ms + 4u * (130 + tid), (uint32_t)(mb / P.m));
      if (tid == 0) sts32(ms, mprev);
      if (tid == 1) sts32(ms + 4u * (cnt + 1), mnext);
      if (tid == 2) sts32(ms + 4u * 258u, 0u);
      it_issue.next();
      if (it_issue.valid()) load_meta(it_issue);  // in flight while this tile's copies are issued
      named_bar_sync(2, BK_PRODUCERS);
      {
        // does any row of the tile appear twice (inside it, or across its borders)?  Tiles of singletons - nearly all of
        // them under uniform ids - take the epilogue's fast path: no exchange, the gather stage is released early.
        const bool dup = tid < cnt && (mkey_cur == lds32(ms + 4u * tid) || mkey_cur == lds32(ms + 4u * (tid + 2)));
        if (__any_sync(0xffffffffu, dup) && (tid & 31) == 0) atomicOr_shared(ms + 4u * 258u, 1u);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = rgx + 8 * i;
        const bool live = r < cnt;
        const uint32_t key = live ? lds32(ms + 4u * (r + 1)) : 0u;
        cp_async16(x_hi(s) + xoff + (uint32_t)i * 1024u, P.table + (int64_t)key * BK_K + 4 * cx, live ? 16u : 0u);
      }
      if (P.scal) {  // the row's (bias, weight) pair rides with the gather instead of stalling the epilogue
        const bool live = tid < cnt;
        cp_async8(scal_st + (uint32_t)s * 1024u + 8u * tid, P.scal + 2 * (int64_t)(live ? mkey_cur : 0u), live ? 8u : 0u);
      }
      ++X;
    };
    float4 gn[8];
    auto load_g = [&](const TileIter& ti, int Yt) {
      const int cnt = ti.cnt();
      const uint32_t mst = meta_base + (uint32_t)(Yt & 3) * BK_META;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rgg + 16 * i;
        gn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < cnt) {
          const int32_t b = (int32_t)lds32(mst + 4u * (130 + r));
          gn[i] = __ldg(reinterpret_cast<const float4*>(P.g1 + (int64_t)b * BK_N1 + 4 * cg));
        }
      }
    };
    if (it_issue.valid()) load_meta(it_issue);
    for (int d = 0; d < BK_DEPTH; ++d) {
      if (it_issue.valid()) issue();
      cp_async_commit();
    }
    for (int Y = 0; it_cons.valid(); ++Y) {
      cp_async_wait<BK_DEPTH - 1>();  // all but the newest BK_DEPTH-1 groups: tile Y has landed (this thread's chunks)
      const int s = Y % BK_NS;
      // this thread's chunks of the tile's g1 rows (L2-resident): requested one tile ahead (gn), used now (gv)
      float4 gv[8];
      if (Y == 0) load_g(it_cons, 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) gv[i] = gn[i];
      {
        TileIter nx = it_cons;
        nx.next();
        if (nx.valid()) load_g(nx, Y + 1);  // its meta was stored when the tile was issued (>= 1 iteration ago)
      }
      if (Y > 0) ok = ok && mbar_wait(lo_free, ((uint32_t)(Y - 1)) & 1u);
      // exact remainders of the truncated operands
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 v = lds128(x_hi(s) + xoff + (uint32_t)i * 1024u);
        sts128(x_lo + xoff + (uint32_t)i * 1024u, trunc_lo4(v));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 lo = trunc_lo4(gv[i]);
        sts128(g1_hi + g1off + (uint32_t)i * 2048u, gv[i]);
        sts128(g1_lo + g1off + (uint32_t)i * 2048u, lo);
        sts128(g2_hi + g2off + (uint32_t)i * 2048u, gv[i]);
        sts128(g2_lo + g2off + (uint32_t)i * 2048u, lo);
      }
      fence_proxy_async_smem();
      mbar_arrive(full(s));
      it_cons.next();
      if (it_issue.valid()) issue();  // waits for the epilogue of tile Y-1 only after tile Y has been handed over
      cp_async_commit();
    }
  } else if (warp == 4) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_tf32_major(128, BK_K, 0, 0);
      const uint32_t idesc2w = umma_idesc_tf32_major(128, 2 * BK_N1, 1, 1);  // x_hi^T x [g_hi | g_lo]
      const uint32_t idesc2 = umma_idesc_tf32_major(128, BK_N1, 1, 1);       // x_lo^T x g_hi
      it.load_unit();
      int Y = 0, G = 0, tin = 0, cur_f = -1, wloads = 0;
      while (it.valid()) {
        const int s = Y % BK_NS, db = Y & 1, gb = G & 1;
        if (it.f != cur_f) {  // new field: its W1 image (all earlier MMAs have finished reading the old one)
          if (Y > 0) ok = ok && mbar_wait(lo_free, ((uint32_t)(Y - 1)) & 1u);
          mbar_arrive_expect_tx(wfull, 16384u);
          bulk_g2s(w_img, P.wpack + (size_t)it.f * 4096, 16384u, wfull);
          ok = ok && mbar_wait(wfull, (uint32_t)wloads & 1u);
          ++wloads;
          cur_f = it.f;
        }
        ok = ok && mbar_wait(full(s), ((uint32_t)(Y / BK_NS)) & 1u);
        ok = ok && mbar_wait(d1_empty(db), (((uint32_t)(Y >> 1)) & 1u) ^ 1u);
        if (tin == 0) ok = ok && mbar_wait(d2_empty(gb), (((uint32_t)(G >> 1)) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d1 = tmem_base + (uint32_t)(64 * db);
        const uint32_t d2 = tmem_base + 128u + (uint32_t)(64 * gb);
        // GEMM 2: dW1_f[c, n] += sum_pos x[pos, c] * g1[pos, n]   (rows 64..127 of the accumulator are unused)
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t ko = (uint32_t)ks * 1024u;
          const uint64_t a_hi = umma_desc_make(x_hi(s) + ko, 16384u, 512u, 1u);
          const uint64_t a_lo = umma_desc_make(x_lo + ko, 16384u, 512u, 1u);
          // g2_lo sits 16 KB behind g2_hi: with the 32-column N atoms 16 KB apart, [g_hi | g_lo] is one N = 64 operand
          const uint64_t b_g = umma_desc_make(g2_hi + ko, 16384u, 512u, 1u);
          umma_tf32(d2, a_hi, b_g, idesc2w, (tin > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(d2, a_lo, b_g, idesc2, 1u);
        }
        // GEMM 1: dx[pos, c] = sum_n g1[pos, n] * W1_f[c, n]
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ko = (uint32_t)ks * 32u;
          const uint64_t a_hi = umma_desc(g1_hi + ko), a_lo = umma_desc(g1_lo + ko);
          const uint64_t b_hi = umma_desc(w_img + ko), b_lo = umma_desc(w_img + 8192u + ko);
          umma_tf32(d1, a_hi, b_hi, idesc1, ks > 0 ? 1u : 0u);
          umma_tf32(d1, a_lo, b_hi, idesc1, 1u);
          umma_tf32(d1, a_hi, b_lo, idesc1, 1u);
        }
        umma_commit(d1_full(db));
        umma_commit(lo_free);
        ++tin;
        if (tin == BK_DRAIN || it.last_in_unit()) {
          umma_commit(d2_full(gb));
          ++G;
          tin = 0;
        }
        ++Y;
        it.next();
      }
    }
  } else {
    // ================================================================== epilogue: BK_NH threads per position
    // thread (j, h): position j = TMEM lane, 16-byte chunks h*CPT .. h*CPT+CPT-1 of the 64-float row
    constexpr int CPT = 16 / BK_NH;   // chunks per thread
    constexpr int DWC = BK_N1 / BK_NH;  // dW1 columns per thread
    const int q = warp & 3, h = (warp - 5) >> 2;  // TMEM lane quadrant of this warp, column group
    const int j = 32 * q + lane;
    // all addressing that does not depend on the tile is computed once
    uint32_t xo[CPT];  // this thread's chunks of row j inside a gather stage
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const uint32_t cq = (uint32_t)(h * CPT + c);
      xo[c] = (cq >> 3) * 16384u + (uint32_t)j * 128u + sw32b_chunk(cq & 7u, (uint32_t)j);
    }
    const int et = tid - 160;  // index inside the epilogue group
    const int qc = et & 15, r0 = et >> 4;  // write-back: chunk qc of rows r0 + (BK_EPI/16) i
    const uint32_t so0 = (uint32_t)(qc >> 3) * 16384u + (uint32_t)r0 * 128u + sw32b_chunk((uint32_t)(qc & 7), (uint32_t)r0);
    const uint32_t tm_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    float dwacc[DWC];
#pragma unroll
    for (int i = 0; i < DWC; ++i) dwacc[i] = 0.f;
    // units without positions still own a slab: zero it
    for (int u = u_begin; u < it.u_end; u += it.stride) {
      const int f = u / P.upf, i = u - f * P.upf;
      if (P.ub[f * (P.upf + 1) + i + 1] <= P.ub[f * (P.upf + 1) + i] && j < BK_K) {
        float* dst = P.slabs + ((int64_t)u * BK_K + j) * BK_N1 + DWC * h;
#pragma unroll
        for (int i4 = 0; i4 < DWC / 4; ++i4) st4(dst + 4 * i4, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    it.load_unit();
    TileIter it_next = it;
    if (it_next.valid()) it_next.next();
    // per-sample operands of the tile (S row part, g_fm, g_lin): loaded one tile ahead when that tile has landed
    float4 Sv[CPT];
    float gf = 0.f, gl = 0.f;
    bool have = false;
    auto load_sample = [&](const TileIter& ti, int Yt) {
      const uint32_t mst = meta_base + (uint32_t)(Yt & 3) * BK_META;
#pragma unroll
      for (int c = 0; c < CPT; ++c) Sv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      gf = 0.f;
      gl = 0.f;
      if (j < ti.cnt()) {
        const int32_t b = (int32_t)lds32(mst + 4u * (130 + j));
        const float4* Sp = reinterpret_cast<const float4*>(P.S + (int64_t)b * BK_K + 4 * CPT * h);
#pragma unroll
        for (int c = 0; c < CPT; ++c) Sv[c] = __ldg(Sp + c);
        gf = __ldg(P.g_fm + b);
        if (h == 0 && P.g_lin) gl = __ldg(P.g_lin + b);
      }
    };
    int Y = 0, G = 0, tin = 0;
    while (it.valid()) {
      const int s = Y % BK_NS, db = Y & 1, gb = G & 1;
      const uint32_t ms = meta_base + (uint32_t)(Y & 3) * BK_META;
      const uint32_t sc_base = sc_base0 + (uint32_t)(Y & 1) * 1024u;
      const uint32_t xs = x_hi(s);
      const int cnt = it.cnt();
      ok = ok && mbar_wait(full(s), ((uint32_t)(Y / BK_NS)) & 1u);
      const bool valid = j < cnt;
      const uint32_t key = lds32(ms + 4u * (j + 1));
      const uint32_t kprev = lds32(ms + 4u * j);
      const uint32_t knext = valid ? lds32(ms + 4u * (j + 2)) : TW_NONE;
      const bool is_head = valid && key != kprev;
      const bool is_tail = valid && key != knext;
      if (!have) load_sample(it, Y);
      have = false;
      ok = ok && mbar_wait(d1_full(db), ((uint32_t)(Y >> 1)) & 1u);
      tc_fence_after();
      uint32_t dr[4 * CPT];
      tmem_ld_cols<4 * CPT>(tm_lane + (uint32_t)(64 * db + 4 * CPT * h), dr);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(d1_empty(db));
      float4 xr[CPT], gr[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        xr[c] = lds128(xs + xo[c]);
        gr[c].x = __uint_as_float(dr[4 * c + 0]) + gf * (Sv[c].x - xr[c].x);
        gr[c].y = __uint_as_float(dr[4 * c + 1]) + gf * (Sv[c].y - xr[c].y);
        gr[c].z = __uint_as_float(dr[4 * c + 2]) + gf * (Sv[c].z - xr[c].z);
        gr[c].w = __uint_as_float(dr[4 * c + 3]) + gf * (Sv[c].w - xr[c].w);
      }
      if (BK_FASTPATH && lds32(ms + 4u * 258u) == 0u) {
        // ---- fast path: every position of the tile is its own segment.  All this tile still needs sits in registers,
        // so the gather stage goes back to the producers before the update math and the stores.
        float2 sold = make_float2(0.f, 0.f);
        if (h == 0 && valid && P.scal) sold = lds64f(scal_st + (uint32_t)s * 1024u + 8u * j);
        mbar_arrive(empty(s));
        if (valid) {
          const int64_t opos = (int64_t)it.p0 + j;
          if (P.out_rows) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) st4(P.out_rows + opos * BK_K + 4 * (h * CPT + c), gr[c]);
          }
          if (P.do_update) {
            float* trow = P.table + (int64_t)key * BK_K + 4 * CPT * h;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              float4 nv;
              nv.x = opt_update(xr[c].x, gr[c].x, P.o);
              nv.y = opt_update(xr[c].y, gr[c].y, P.o);
              nv.z = opt_update(xr[c].z, gr[c].z, P.o);
              nv.w = opt_update(xr[c].w, gr[c].w, P.o);
              st4(trow + 4 * c, nv);
            }
          }
          if (h == 0) {
            if (P.out_scal) *reinterpret_cast<float2*>(P.out_scal + 2 * opos) = make_float2(gf, gl);
            if (P.scal && P.do_update) {
              float2 nv;
              nv.x = opt_update(sold.x, gf, P.o);
              nv.y = P.g_lin ? opt_update(sold.y, gl, P.o) : sold.y;
              *reinterpret_cast<float2*>(P.scal + 2 * (int64_t)key) = nv;
            }
          }
        }
      } else {
      float asf = gf, asl = gl;
      const bool single = is_head && is_tail;
      if (valid && !single) {  // members of longer segments exchange their rows through the (now dead) x tile
#pragma unroll
        for (int c = 0; c < CPT; ++c) sts128(xs + xo[c], gr[c]);
        if (h == 0) {
          sts32(sc_base + 8u * j, __float_as_uint(gf));
          sts32(sc_base + 8u * j + 4u, __float_as_uint(gl));
        }
      }
      // next tile's per-sample operands: requested now (if that tile has landed), consumed one iteration later
      if (BK_PREFETCH && it_next.valid()) {
        const int sn = (Y + 1) % BK_NS;
        uint32_t done;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(full(sn)), "r"(((uint32_t)((Y + 1) / BK_NS)) & 1u)
            : "memory");
        if (__all_sync(0xffffffffu, done != 0)) {
          load_sample(it_next, Y + 1);
          have = true;
        }
      }
      named_bar_sync(3, BK_EPI);
      uint32_t wkey = TW_NONE;  // table row this position finished (its new values sit in the x tile)
      if constexpr (!COOP) {
      // ---- common variant (no hot rows in the batch): every segment leader walks its run itself
      if (valid && (is_head || j == 0)) {  // leader of a segment (or of its part inside this tile)
        if (!is_head) {  // continues from the previous tile: the carried partial sum comes first (position order)
          const uint32_t cb = carry_base + (uint32_t)((Y + 1) & 1) * 256u + (uint32_t)(h * CPT) * 16u;
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const float4 cv = lds128(cb + 16u * c);
            gr[c].x = cv.x + gr[c].x; gr[c].y = cv.y + gr[c].y; gr[c].z = cv.z + gr[c].z; gr[c].w = cv.w + gr[c].w;
          }
          if (h == 0) {
            const float2 cs = lds64f(carry_sc + (uint32_t)((Y + 1) & 1) * 8u);
            asf = cs.x + asf;
            asl = cs.y + asl;
          }
        }
        int jj = j + 1;
        if (!is_tail) {
          while (jj < cnt && lds32(ms + 4u * (jj + 1)) == key) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              const uint32_t cq = (uint32_t)(h * CPT + c);
              const float4 v = lds128(xs + (cq >> 3) * 16384u + (uint32_t)jj * 128u + sw32b_chunk(cq & 7u, (uint32_t)jj));
              gr[c].x += v.x; gr[c].y += v.y; gr[c].z += v.z; gr[c].w += v.w;
            }
            if (h == 0) {
              const float2 sv = lds64f(sc_base + 8u * jj);
              asf += sv.x;
              asl += sv.y;
            }
            ++jj;
          }
        }
        const bool closed = lds32(ms + 4u * (jj + 1)) != key;
        if (closed) {
          const int64_t opos = (int64_t)it.p0 + jj - 1;  // sorted position that closes the segment
          if (P.out_rows) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) st4(P.out_rows + opos * BK_K + 4 * (h * CPT + c), gr[c]);
          }
          if (P.do_update) {  // new row -> this thread's slot of the (dead) x tile; stored 256 B at a time below
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              float4 nv;
              nv.x = opt_update(xr[c].x, gr[c].x, P.o);
              nv.y = opt_update(xr[c].y, gr[c].y, P.o);
              nv.z = opt_update(xr[c].z, gr[c].z, P.o);
              nv.w = opt_update(xr[c].w, gr[c].w, P.o);
              sts128(xs + xo[c], nv);
            }
            wkey = key;
          }
          if (h == 0) {
            if (P.out_scal) *reinterpret_cast<float2*>(P.out_scal + 2 * opos) = make_float2(asf, asl);
            if (P.scal && P.do_update) {
              const float2 sold = lds64f(scal_st + (uint32_t)s * 1024u + 8u * j);
              float2 nv;
              nv.x = opt_update(sold.x, asf, P.o);
              nv.y = P.g_lin ? opt_update(sold.y, asl, P.o) : sold.y;
              *reinterpret_cast<float2*>(P.scal + 2 * (int64_t)key) = nv;
            }
          }
        } else {
          const uint32_t cb = carry_base + (uint32_t)(Y & 1) * 256u + (uint32_t)(h * CPT) * 16u;
#pragma unroll
          for (int c = 0; c < CPT; ++c) sts128(cb + 16u * c, gr[c]);
          if (h == 0) {
            sts32(carry_sc + (uint32_t)(Y & 1) * 8u, __float_as_uint(asf));
            sts32(carry_sc + (uint32_t)(Y & 1) * 8u + 4u, __float_as_uint(asl));
          }
        }
      }
      } else {
      // ---- hot-row variant: runs longer than BK_WALK are handed to the whole warp
      const bool leader = valid && (is_head || j == 0);  // leader of a segment (or of its part inside this tile)
      int jj = j + 1;
      bool long_run = false;
      if (leader) {
        if (!is_head) {  // continues from the previous tile: the carried partial sum comes first (position order)
          const uint32_t cb = carry_base + (uint32_t)((Y + 1) & 1) * 256u + (uint32_t)(h * CPT) * 16u;
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const float4 cv = lds128(cb + 16u * c);
            gr[c].x = cv.x + gr[c].x; gr[c].y = cv.y + gr[c].y; gr[c].z = cv.z + gr[c].z; gr[c].w = cv.w + gr[c].w;
          }
          if (h == 0) {
            const float2 cs = lds64f(carry_sc + (uint32_t)((Y + 1) & 1) * 8u);
            asf = cs.x + asf;
            asl = cs.y + asl;
          }
        }
        if (!is_tail) {
          // short runs (the common case) are walked by their leader thread; a run that goes on past BK_WALK positions
          // is summed by the whole warp below - same ascending order, one column per lane instead of 32 per thread
          int steps = 0;
          while (jj < cnt && lds32(ms + 4u * (jj + 1)) == key) {
            if (COOP && steps == BK_WALK) {
              long_run = true;
              break;
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              const uint32_t cq = (uint32_t)(h * CPT + c);
              const float4 v = lds128(xs + (cq >> 3) * 16384u + (uint32_t)jj * 128u + sw32b_chunk(cq & 7u, (uint32_t)jj));
              gr[c].x += v.x; gr[c].y += v.y; gr[c].z += v.z; gr[c].w += v.w;
            }
            if (h == 0) {
              const float2 sv = lds64f(sc_base + 8u * jj);
              asf += sv.x;
              asl += sv.y;
            }
            ++jj;
            ++steps;
          }
          if (long_run) {  // hand the running sum to the warp through this position's slot of the (dead) x tile
#pragma unroll
            for (int c = 0; c < CPT; ++c) sts128(xs + xo[c], gr[c]);
          }
        }
      }
      if (COOP && lds32(ms + 4u * 258u) != 0u) {  // the tile repeats a row somewhere (flag set by the producers)
        const uint32_t lm = __ballot_sync(0xffffffffu, long_run);
        if (lm) {  // rare: kept out of line so that the common path's register allocation is unaffected
          int jj_io = jj;
          float asf_io = asf, asl_io = asl;
          bk_sum_long_runs(lm, xs, ms, sc_base, cnt, h, q, lane, key, jj_io, asf_io, asl_io);
          if (long_run) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) gr[c] = lds128(xs + xo[c]);
            jj = jj_io;
            asf = asf_io;
            asl = asl_io;
          }
        }
      }
            if (leader) {
        const bool closed = lds32(ms + 4u * (jj + 1)) != key;
        if (closed) {
          const int64_t opos = (int64_t)it.p0 + jj - 1;  // sorted position that closes the segment
          if (P.out_rows) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) st4(P.out_rows + opos * BK_K + 4 * (h * CPT + c), gr[c]);
          }
          if (P.do_update) {  // new row -> this thread's slot of the (dead) x tile; stored 256 B at a time below
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              float4 nv;
              nv.x = opt_update(xr[c].x, gr[c].x, P.o);
              nv.y = opt_update(xr[c].y, gr[c].y, P.o);
              nv.z = opt_update(xr[c].z, gr[c].z, P.o);
              nv.w = opt_update(xr[c].w, gr[c].w, P.o);
              sts128(xs + xo[c], nv);
            }
            wkey = key;
          }
          if (h == 0) {
            if (P.out_scal) *reinterpret_cast<float2*>(P.out_scal + 2 * opos) = make_float2(asf, asl);
            if (P.scal && P.do_update) {
              const float2 sold = lds64f(scal_st + (uint32_t)s * 1024u + 8u * j);
              float2 nv;
              nv.x = opt_update(sold.x, asf, P.o);
              nv.y = P.g_lin ? opt_update(sold.y, asl, P.o) : sold.y;
              *reinterpret_cast<float2*>(P.scal + 2 * (int64_t)key) = nv;
            }
          }
        } else {
          const uint32_t cb = carry_base + (uint32_t)(Y & 1) * 256u + (uint32_t)(h * CPT) * 16u;
#pragma unroll
          for (int c = 0; c < CPT; ++c) sts128(cb + 16u * c, gr[c]);
          if (h == 0) {
            sts32(carry_sc + (uint32_t)(Y & 1) * 8u, __float_as_uint(asf));
            sts32(carry_sc + (uint32_t)(Y & 1) * 8u + 4u, __float_as_uint(asl));
          }
        }
      }
      }
      if (P.do_update) {
        // coalesced write-back: 16 lanes per finished row store its 256 bytes contiguously
        const uint32_t wb = wr_base + (uint32_t)(Y & 1) * 512u;
        if (h == 0) sts32(wb + 4u * j, wkey);
        named_bar_sync(3, BK_EPI);
#pragma unroll
        for (int i = 0; i < 2048 / BK_EPI; ++i) {
          const uint32_t wk = lds32(wb + 4u * (uint32_t)(r0 + (BK_EPI / 16) * i));
          if (wk != TW_NONE) {
            const float4 v = lds128(xs + so0 + (uint32_t)i * ((BK_EPI / 16) * 128u));
            st4(P.table + (int64_t)wk * BK_K + 4 * qc, v);
          }
        }
      }
      mbar_arrive(empty(s));
      }
      ++tin;
      const bool unit_end = it.last_in_unit();
      if (tin == BK_DRAIN || unit_end) {  // drain the GEMM 2 accumulator (fp32 round-to-nearest adds)
        ok = ok && mbar_wait(d2_full(gb), ((uint32_t)(G >> 1)) & 1u);
        tc_fence_after();
        // accumulator = [x_hi*g_hi + x_lo*g_hi | x_hi*g_lo]: columns n and 32 + n belong together
        uint32_t w0[DWC], w1[DWC];
        const uint32_t ta = tm_lane + 128u + (uint32_t)(64 * gb + DWC * h);
        tmem_ld_cols<DWC>(ta, w0);
        tmem_ld_cols<DWC>(ta + 32u, w1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(d2_empty(gb));
#pragma unroll
        for (int i = 0; i < DWC; ++i) dwacc[i] += __uint_as_float(w0[i]) + __uint_as_float(w1[i]);
        ++G;
        tin = 0;
        if (unit_end) {
          if (j < BK_K) {
            float* dst = P.slabs + ((int64_t)it.u * BK_K + j) * BK_N1 + DWC * h;
#pragma unroll
            for (int i4 = 0; i4 < DWC / 4; ++i4)
              st4(dst + 4 * i4, make_float4(dwacc[4 * i4], dwacc[4 * i4 + 1], dwacc[4 * i4 + 2], dwacc[4 * i4 + 3]));
          }
#pragma unroll
          for (int i = 0; i < DWC; ++i) dwacc[i] = 0.f;
        }
      }
      ++Y;
      it.next();
      if (it_next.valid()) it_next.next();
    }
  }
  if (!ok && P.status) atomicOr(P.status, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_free_cols(tmem_base, 256);
  }
}

constexpr size_t BK_SMEM = 1024 + BK_NS * BK_STAGE + BK_XT + 4 * BK_GT + 16384 + 4 * BK_META + 2048 + 512 + 16 +
                           BK_NS * 1024 + 1024 + 8 * 17 + 64;

// ------------------------------------------------------------------------------------------------ UMMA layout probe
// D[128, 32] = At^T @ Bt for At [K, 128], Bt [K, 32] (K <= 64, multiple of 8) with both operands MN-major, built with
// exactly the tile layout and descriptors of GEMM 2 above.  Test-only entry point (tests/test_tower_gpu.py).
__global__ void __launch_bounds__(160, 1) umma_probe_kernel(const float* __restrict__ At, const float* __restrict__ Bt,
                                                            int K, int variant, float* __restrict__ D,
                                                            int32_t* status) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 65536u, bar = base + 65536u + 16384u, slot = bar + 8u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < K * 128; i += 160) {
    const int kk = i / 128, mm = i - kk * 128;
    sts32(a_base + (uint32_t)(mm >> 5) * 16384u + (uint32_t)kk * 128u +
              sw32b_chunk((uint32_t)((mm & 31) >> 2), (uint32_t)kk) + 4u * (mm & 3),
          __float_as_uint(At[i]));
  }
  for (int i = tid; i < K * 32; i += 160) {
    const int kk = i / 32, nn = i - kk * 32;
    sts32(b_base + (uint32_t)kk * 128u + sw32b_chunk((uint32_t)(nn >> 2), (uint32_t)kk) + 4u * (nn & 3), __float_as_uint(Bt[i]));
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 4) tmem_alloc_cols(slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(slot);
  bool ok = true;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = umma_idesc_tf32_major(128, 32, 1, 1);
    const uint32_t lbo = variant == 0 ? 16384u : 512u, sbo = variant == 0 ? 512u : 16384u;
    for (int ks = 0; ks < K / 8; ++ks) {
      umma_tf32(tmem_base, umma_desc_make(a_base + (uint32_t)ks * 1024u, lbo, sbo, 1u),
                umma_desc_make(b_base + (uint32_t)ks * 1024u, lbo, sbo, 1u), idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  if (warp < 4) {
    ok = mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16), v);
    tmem_ld_wait();
    for (int n = 0; n < 32; ++n) D[(32 * warp + lane) * 32 + n] = __uint_as_float(v[n]);
  }
  if (!ok && status) atomicOr(status, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_free_cols(tmem_base, 32);
  }
}

// ---- owner-side plan for row-sharded tables: from the ids of ALL ranks (gids [W*b, m], rank-major) keep the entries
// this rank owns (id mod W == rank) in ascending global position, key them by owner-local row, sort
struct TowerOwned {
  const int32_t* gids;
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

__global__ void __launch_bounds__(256) tower_shard_keys_kernel(const int32_t* __restrict__ gids,
                                                               const int64_t* __restrict__ local_offs, uint32_t m,
                                                               int W, int wshift, const int32_t* __restrict__ own_gpos,
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
      shard_of((int64_t)gids[gp], W, wshift, owner, lr);
      keys[i] = (uint32_t)(local_offs[(uint32_t)gp % m] + lr);
      pos[i] = gp;
    } else {
      keys[i] = sentinel;
      pos[i] = 0;
    }
  }
}

struct TowerPlanWs {
  uint32_t* keys_in;
  int32_t* pos_in;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static TowerPlanWs tower_plan_layout(int64_t N, void* basep) {
  TowerPlanWs w;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)N, 0, 32);
  w.cub_bytes = sort_bytes;
  char* b = (char*)basep;
  size_t off = 0;
  const size_t arr = align_up((size_t)N * 4, 256);
  w.keys_in = (uint32_t*)(b + off); off += arr;
  w.pos_in = (int32_t*)(b + off); off += arr;
  w.cub_temp = (void*)(b + off); off += align_up(sort_bytes, 256);
  w.total = off;
  return w;
}

}  // namespace rm

extern "C" {

int32_t rm_tower_units_per_field(int64_t B, int32_t unit) { return unit > 0 ? (int32_t)((B + unit - 1) / unit) : 0; }

size_t rm_tower_plan_workspace_bytes(int64_t N) {
  if (N <= 0) return 256;
  return rm::tower_plan_layout(N, nullptr).total;
}

// Sort the (table row, position) pairs of one batch and cut every field's range into work units.
//   sorted_keys [N] uint32, sorted_pos [N] int32, field_bounds [m+1] int32, unit_bounds [m*(upf+1) + 1] int32 (cuts + hot-row flag)
int rm_tower_plan(const int64_t* ids, const int64_t* table_offsets, int64_t B, int32_t m, int64_t total_rows,
                  int32_t unit, void* workspace, size_t workspace_bytes, uint32_t* sorted_keys, int32_t* sorted_pos,
                  int32_t* field_bounds, int32_t* unit_bounds, int32_t* status, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(ids && table_offsets && workspace && sorted_keys && sorted_pos && field_bounds && unit_bounds,
               "null pointer");
  RM_CHECK_ARG(B > 0 && m > 0 && total_rows > 0 && unit >= BK_TILE, "bad shape");
  const int64_t N = B * m;
  RM_UNSUPPORTED(N < ((int64_t)1 << 31) - 1, "B*m must be < 2^31 - 1");
  RM_UNSUPPORTED(total_rows < ((int64_t)1 << 32) - 1, "total_rows must be < 2^32 - 1 (32-bit sort keys + sentinel)");
  cudaStream_t st = (cudaStream_t)stream;
  TowerPlanWs w = tower_plan_layout(N, workspace);
  if (workspace_bytes < w.total) {
    set_error("rm_tower_plan: workspace %zu < required %zu", workspace_bytes, w.total);
    return RM_E_WORKSPACE;
  }
  const uint32_t sentinel = (uint32_t)total_rows;
  tower_keys_kernel<<<grid_for(N, 256, 8), 256, 0, st>>>(ids, table_offsets, (uint32_t)N, (uint32_t)m, sentinel,
                                                         w.keys_in, w.pos_in, status);
  RM_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 32 && ((int64_t)1 << end_bit) <= total_rows) ++end_bit;
  size_t bytes = w.cub_bytes;
  RM_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, bytes, (const uint32_t*)w.keys_in, sorted_keys,
                                          (const int32_t*)w.pos_in, sorted_pos, (int)N, 0, end_bit, st));
  count_launch();
  const int upf = rm_tower_units_per_field(B, unit);
  tower_bounds_kernel<<<1, 1024, (m + 1) * sizeof(int32_t), st>>>(sorted_keys, (int32_t)N, table_offsets, m, unit, upf,
                                                                  field_bounds, unit_bounds);
  RM_LAUNCH_CHECK();
  tower_hot_flag_kernel<<<grid_for(N / BK_HOT_RUN + 1, 256, 4), 256, 0, st>>>(sorted_keys, field_bounds, m,
                                                                             unit_bounds + (size_t)m * (upf + 1));
  RM_LAUNCH_CHECK();
  return 0;
}

size_t rm_tower_shard_plan_workspace_bytes(int64_t Ntot, int64_t N_cap) {
  if (Ntot <= 0 || N_cap <= 0) return 256;
  size_t sort_bytes = 0, sel_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)N_cap, 0, 32);
  thrust::counting_iterator<int32_t> counting(0);
  rm::TowerOwned own{nullptr, nullptr, 1, 1, 0, 0};
  cub::DeviceSelect::If(nullptr, sel_bytes, counting, (int32_t*)nullptr, (int32_t*)nullptr, (int)Ntot, own);
  const size_t cubb = sort_bytes > sel_bytes ? sort_bytes : sel_bytes;
  return rm::align_up((size_t)Ntot * 4, 256) + 2 * rm::align_up((size_t)N_cap * 4, 256) + rm::align_up(cubb, 256) + 256;
}

// Owner-side plan of the fused backward for row-sharded tables (see rm_shard_plan for the conventions): outputs as
// rm_tower_plan with keys = owner-local rows and sorted_pos = GLOBAL positions gp = src_rank*(b*m) + p, over the fixed
// capacity N_cap (entries past the owned count carry the sentinel key and sort last); n_own[1] = owned count.
// gids are int32 (the all-gather moves half the bytes of the reference's int64 ids; table sizes are < 2^31).
// `Bcap` = per-field capacity used to size the work units (rm_tower_units_per_field(Bcap, unit)).
int rm_tower_shard_plan(const int32_t* gids, int64_t Ntot, int32_t m, int32_t W, int32_t rank, const int64_t* feat_sizes,
                        const int64_t* local_offsets_m1, int64_t total_local, int64_t N_cap, int64_t Bcap, int32_t unit,
                        void* workspace, size_t workspace_bytes, uint32_t* sorted_keys, int32_t* sorted_gpos,
                        int32_t* field_bounds, int32_t* unit_bounds, int32_t* n_own, int32_t* status, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(gids && feat_sizes && local_offsets_m1 && workspace && sorted_keys && sorted_gpos && field_bounds &&
                   unit_bounds && n_own, "null pointer");
  RM_CHECK_ARG(Ntot > 0 && m > 0 && N_cap > 0 && total_local > 0 && rank >= 0 && rank < W && unit >= BK_TILE && Bcap > 0,
               "bad shape");
  RM_UNSUPPORTED(W >= 1 && W <= 8, "world size must be <= 8");
  RM_UNSUPPORTED(Ntot < ((int64_t)1 << 31) - 1 && N_cap <= Ntot, "W*B*m must be < 2^31 - 1 and N_cap <= W*B*m");
  RM_UNSUPPORTED(total_local < ((int64_t)1 << 31), "local rows must be < 2^31");
  const size_t need = rm_tower_shard_plan_workspace_bytes(Ntot, N_cap);
  if (workspace_bytes < need) {
    set_error("rm_tower_shard_plan: workspace %zu < required %zu", workspace_bytes, need);
    return RM_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* b = (char*)workspace;
  int32_t* own_gpos = (int32_t*)b; b += align_up((size_t)Ntot * 4, 256);
  uint32_t* keys_in = (uint32_t*)b; b += align_up((size_t)N_cap * 4, 256);
  int32_t* pos_in = (int32_t*)b; b += align_up((size_t)N_cap * 4, 256);
  void* cub_temp = (void*)b;
  size_t cub_bytes = (size_t)((char*)workspace + workspace_bytes - b);
  const int wshift = world_shift(W);
  thrust::counting_iterator<int32_t> counting(0);
  TowerOwned own{gids, feat_sizes, (uint32_t)m, (int)W, wshift, (int)rank};
  size_t bytes = cub_bytes;
  RM_CUDA(cub::DeviceSelect::If(cub_temp, bytes, counting, own_gpos, n_own, (int)Ntot, own, st));
  count_launch();
  int end_bit = 1;
  while (end_bit < 31 && ((int64_t)1 << end_bit) <= total_local) ++end_bit;
  const uint32_t sentinel = (uint32_t)total_local;  // sorts behind every owned row
  tower_shard_keys_kernel<<<grid_for(N_cap, 256, 8), 256, 0, st>>>(gids, local_offsets_m1, (uint32_t)m, (int)W, wshift, own_gpos,
                                                                  n_own, (int32_t)N_cap, sentinel, keys_in, pos_in, status);
  RM_LAUNCH_CHECK();
  bytes = cub_bytes;
  RM_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, bytes, (const uint32_t*)keys_in, sorted_keys, (const int32_t*)pos_in,
                                          sorted_gpos, (int)N_cap, 0, end_bit, st));
  count_launch();
  const int upf = rm_tower_units_per_field(Bcap, unit);
  tower_bounds_local_kernel<<<1, 1024, (m + 1) * sizeof(int32_t), st>>>(sorted_keys, (int32_t)N_cap, local_offsets_m1,
                                                                        (uint32_t)total_local, m, unit, upf, field_bounds,
                                                                        unit_bounds);
  RM_LAUNCH_CHECK();
  tower_hot_flag_kernel<<<grid_for(N_cap / BK_HOT_RUN + 1, 256, 4), 256, 0, st>>>(sorted_keys, field_bounds, m,
                                                                                 unit_bounds + (size_t)m * (upf + 1));
  RM_LAUNCH_CHECK();
  return 0;
}

size_t rm_tower_bwd_workspace_bytes(int64_t B, int32_t m, int32_t unit) {
  const int upf = rm_tower_units_per_field(B, unit);
  return 256 + (size_t)m * 2 * rm::BK_K * 128 + (size_t)m * upf * rm::BK_K * rm::BK_N1 * 4;
}

// Fused backward + optimizer update (see the header of this file).  k = 64 and N1 = 32 only.
int rm_tower_bwd_update(float* table, float* scal, const uint32_t* sorted_keys, const int32_t* sorted_pos,
                        const int32_t* unit_bounds, const float* g1, const float* S, const float* g_fm,
                        const float* g_lin, const float* W1, int64_t B, int32_t m, int32_t k, int32_t N1, int32_t unit,
                        int32_t opt, float lr, float l2, int32_t variant, float* dW1, float* out_rows, float* out_scal,
                        int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(table && sorted_keys && sorted_pos && unit_bounds && g1 && S && g_fm && W1 && dW1 && workspace,
               "null pointer");
  RM_CHECK_ARG(variant >= RM_TOWER_BWD_AUTO && variant <= RM_TOWER_BWD_HOT, "unknown variant");
  RM_CHECK_ARG(B > 0 && m > 0 && unit >= BK_TILE, "bad shape");
  RM_UNSUPPORTED(k == BK_K && N1 == BK_N1, "the fused tower backward is built for k = 64, first hidden layer = 32");
  RM_UNSUPPORTED(aligned16(table) && aligned16(g1) && aligned16(S) && aligned16(workspace) &&
                     (!out_rows || aligned16(out_rows)) && (!scal || (reinterpret_cast<uintptr_t>(scal) & 7) == 0),
                 "tower backward needs 16-byte aligned rows");
  const size_t need = rm_tower_bwd_workspace_bytes(B, m, unit);
  if (workspace_bytes < need) {
    set_error("rm_tower_bwd_update: workspace %zu < required %zu", workspace_bytes, need);
    return RM_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  TowerBwdParams P;
  P.do_update = opt != RM_OPT_NONE;
  const int rc = make_params(P.do_update ? opt : RM_OPT_GD, lr, l2, &P.o);
  if (rc) return rc;
  const int upf = rm_tower_units_per_field(B, unit);
  uint32_t* wpack = (uint32_t*)((char*)workspace + 256);
  float* slabs = (float*)((char*)workspace + 256 + (size_t)m * 2 * BK_K * 128);
  tower_pack_w1_kernel<<<grid_for((int64_t)m * 2 * BK_K * 8, 256, 8), 256, 0, st>>>(W1, m, wpack);
  RM_LAUNCH_CHECK();
  P.table = table; P.scal = scal; P.keys = sorted_keys; P.spos = sorted_pos; P.ub = unit_bounds; P.g1 = g1; P.S = S;
  P.g_fm = g_fm; P.g_lin = g_lin; P.wpack = wpack; P.slabs = slabs; P.out_rows = out_rows; P.out_scal = out_scal;
  P.status = status; P.m = m; P.upf = upf; P.n_units = m * upf;

  RM_SMEM_ATTR_ONCE(BK_SMEM, tower_bwd_kernel<false>);
  RM_SMEM_ATTR_ONCE(BK_SMEM, tower_bwd_kernel<true>);
  const int grid = P.n_units < RM_NUM_SMS ? P.n_units : RM_NUM_SMS;
  // AUTO: both variants are launched and the plan's hot-row flag (device side) picks one - the other returns at once
  // (~14 us).  A caller that knows what its ids look like (e.g. from the previous steps' flags) launches one variant only;
  // either variant is correct on any input, they differ in speed.
  P.force = variant != RM_TOWER_BWD_AUTO || BK_NH != 2;
  if (variant != RM_TOWER_BWD_HOT || BK_NH != 2) {
    tower_bwd_kernel<false><<<grid, BK_THREADS, BK_SMEM, st>>>(P);
    RM_LAUNCH_CHECK();
  }
  if (variant != RM_TOWER_BWD_PLAIN && BK_NH == 2) {
    tower_bwd_kernel<true><<<grid, BK_THREADS, BK_SMEM, st>>>(P);
    RM_LAUNCH_CHECK();
  }
  tower_dw_reduce_kernel<<<grid_for((int64_t)m * BK_K * BK_N1, 256, 8), 256, 0, st>>>(slabs, m, upf, dW1);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_umma_probe(const float* At, const float* Bt, int32_t K, int32_t variant, float* D, int32_t* status,
                  void* stream) {
  using namespace rm;
  RM_CHECK_ARG(At && Bt && D && K >= 8 && K <= 64 && K % 8 == 0, "bad probe arguments");
  const int smem = 1024 + 65536 + 16384 + 64;
  RM_SMEM_ATTR_ONCE(smem, umma_probe_kernel);
  umma_probe_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(At, Bt, K, variant, D, status);
  RM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
