// Fused DeepFM front end + first MLP layer ("tower" forward), one kernel:
//
//   gather   tf.nn.embedding_lookup per field            recman/tf/core/layers.py:117-128, 238-261
//   FM       sum-square second order + bias first order  layers.py:457-478
//   linear   first-order term (k=1 lookups + dense . w)   layers.py:330-347
//   DNN[0]   y1 = [embeds | dense] @ W1 + b1 (pre-act)    layers.py:589-609 (first matmul + bias_add)
//
// The embedding rows of a 128-sample tile are copied field by field with cp.async straight into the K-major
// SWIZZLE_128B shared-memory layout the tensor core reads (a row of k = 64 floats is two 128-byte swizzle rows), so a
// row is fetched from HBM once and feeds the FM sums (CUDA cores), the first-order lookups and the layer-1 GEMM
// (tcgen05, 3xTF32: a = trunc(a) + lo, w = rna(w) + lo, fp32 accumulation in TMEM) without ever being written back -
// the [B, m*k] row buffer is optional (only the unfused backward needs it).  The two k=1 tables are one interleaved
// [rows, 2] array: one 8-byte lookup per id instead of two 128-byte sector fetches.
//
// Roofline: HBM.  Algorithmic bytes per sample: m*(8 + 4k + 8) + 4*n_dense (+ 4k S + 4 N1 y1 + 8 written).
//
// CTA = 128 samples, 320 threads:
//   warps 0-7  producers: cp.async gather (2 fields in flight), then per landed field: S/Q accumulation, lo tile,
//              optional row-buffer store; afterwards the epilogue (TMEM -> +b1 + dense part -> y1).
//   warp  8    lane 0 issues tcgen05.mma (M=128, N=N1PAD, K=8), 24 per field at k=64; owns TMEM.
//   warp  9    lane 0 streams the pre-packed W1 field images (cp.async.bulk + mbarrier tx-count).
// The TMEM accumulation truncates (round toward zero, bias ~ 0.5 ulp per accumulate), so fields rotate over NACC
// accumulators that the epilogue adds in fp32 round-to-nearest: <= 24*ceil(m/NACC) accumulates each.
// Every mbarrier wait is bounded (status bit 2).
#include "tower_common.cuh"

namespace rm {

constexpr int TF_ROWS = 128;
constexpr int TF_STAGES = 4;   // gather stages (one field of the tile each)
constexpr int TF_DEPTH = 3;    // fields in flight per producer thread
constexpr int TF_WSLOTS = 2;   // W1 field images in flight
constexpr int TF_PRODUCERS = 256;
constexpr int TF_THREADS = 320;

constexpr int TF_MAX_PEERS = 8;
constexpr uint32_t TF_ROW_BITS = 29;  // row code = owner rank << 29 | row inside the owner's table
constexpr uint32_t TF_ROW_MASK = (1u << TF_ROW_BITS) - 1u;

struct TowerFwdParams {
  const float* table[TF_MAX_PEERS];  // [rows, k] of every owner rank (W == 1: the one table); peer-mapped memory for W > 1
  const float* scal[TF_MAX_PEERS];   // [rows, 2] = (bias, linear weight) per owner, nullable
  const int64_t* offs;               // W == 1: [m+1] global row offsets; W > 1: [m] owner-local first rows
  const int64_t* feat_sizes;         // W > 1: [m] global table sizes (range check)
  int W, wshift;
  const int64_t* ids;
  const float* dense;
  const float* lin_dense;
  int lin_dense_stride;
  const uint32_t* wpack;
  const float* W1;
  const float* b1;
  float* x;
  int64_t ld;
  float* y1;
  float* fm_out;
  float* lin_out;
  float* sum_out;
  int32_t* status;
  int64_t B;
  int m, nd, N1, N1PAD, NACC;
  uint32_t tmem_cols;
};

// W1 field images for the forward: B operand = W1_f^T, K-major.  Per field f: [blk][part hi|lo][n < N1PAD][128 B] - the hi
// and lo rows of a block are adjacent, so [w_hi | w_lo] is ONE operand of N = 2*N1PAD rows;
// chunk ch of row n holds W1[f*k + blk*32 + 4ch .. +3][n] at chunk position ch ^ (n & 7).
__global__ void __launch_bounds__(256) tower_pack_w1t_kernel(const float* __restrict__ W1, int m, int k, int N1,
                                                             int N1PAD, uint32_t* __restrict__ out) {
  const int KB = k / 32;
  const int64_t total = (int64_t)m * 2 * KB * N1PAD * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i & 7);
    int64_t t = i >> 3;
    const int n = (int)(t % N1PAD); t /= N1PAD;
    const int part = (int)(t & 1); t >>= 1;
    const int blk = (int)(t % KB);
    const int f = (int)(t / KB);
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kk = f * k + blk * 32 + 4 * ch + e;
      const float w = n < N1 ? W1[(int64_t)kk * N1 + n] : 0.f;
      const uint32_t hi = f32_to_tf32(w);
      v[e] = part ? f32_to_tf32(w - __uint_as_float(hi)) : hi;
    }
    uint32_t* dst = out + ((((int64_t)f * KB + blk) * 2 + part) * N1PAD + n) * 32 + ((ch ^ (n & 7)) << 2);
    *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

template <int KB>
__global__ void __launch_bounds__(TF_THREADS, 1) tower_fwd_kernel(const TowerFwdParams P) {
  constexpr int K = 32 * KB;
  constexpr int R = 4;                  // rows per producer thread: rg + 32 i, 16-byte chunk c8 of every 32-column block
  constexpr uint32_t XT = KB * 16384u;  // bytes of one [128 x k] tile: KB blocks of [128 rows x 128 B]
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t WB = 2u * KB * (uint32_t)P.N1PAD * 128u;  // bytes of one W1 field image: per block [hi rows | lo rows]
  const uint32_t lo_base = base + TF_STAGES * XT;          // 2 remainder ("lo") tiles of one 32-column block each
  const uint32_t w_base = lo_base + 2u * 16384u;
  const uint32_t scal_base = w_base + TF_WSLOTS * WB;      // [stage][128] float2
  const uint32_t rowidx_base = scal_base + TF_STAGES * 1024u;
  const uint32_t w1d_base = rowidx_base + (((uint32_t)TF_ROWS * P.m * 4u + 15u) & ~15u);
  const uint32_t b1_base = w1d_base + (uint32_t)P.nd * P.N1PAD * 4u;
  const uint32_t rowsum_base = b1_base + (uint32_t)P.N1PAD * 4u;  // [128] second-order terms
  const uint32_t bar_base = (rowsum_base + TF_ROWS * 4u + 7u) & ~7u;
  auto empty = [&](int s) { return bar_base + 8u * s; };
  auto wfull = [&](int s) { return bar_base + 8u * (TF_STAGES + s); };
  auto wempty = [&](int s) { return bar_base + 8u * (TF_STAGES + TF_WSLOTS + s); };
  auto lo_ready = [&](int s) { return bar_base + 8u * (TF_STAGES + 2 * TF_WSLOTS + s); };
  auto lo_free = [&](int s) { return bar_base + 8u * (TF_STAGES + 2 * TF_WSLOTS + 2 + s); };
  const uint32_t accum_full = bar_base + 8u * (TF_STAGES + 2 * TF_WSLOTS + 4);
  const uint32_t tmem_slot = accum_full + 8u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b0 = (int64_t)blockIdx.x * TF_ROWS;
  const int m = P.m;

  if (tid == 0) {
    for (int s = 0; s < TF_STAGES; ++s) mbar_init(empty(s), 1);
    for (int s = 0; s < TF_WSLOTS; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(lo_ready(s), TF_PRODUCERS);
      mbar_init(lo_free(s), 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
  }
  // row index of every (sample, field) of the tile: offs[f] + id, TW_NONE when the id is outside its table
  for (int idx = tid; idx < TF_ROWS * m; idx += TF_THREADS) {
    const int r = idx / m, f = idx - r * m;
    const int64_t b = b0 + r;
    uint32_t row = TW_NONE;
    if (b < P.B) {
      const int64_t id = P.ids[b * m + f];
      if (P.W == 1) {
        const int64_t lo = P.offs[f], hi = P.offs[f + 1];
        if (id >= 0 && id < hi - lo) row = (uint32_t)(lo + id);
      } else if (id >= 0 && id < P.feat_sizes[f]) {
        // row-sharded tables: global row id lives on rank id mod W at local row id div W
        int owner;
        int64_t lr;
        shard_of(id, P.W, P.wshift, owner, lr);
        row = ((uint32_t)owner << TF_ROW_BITS) | (uint32_t)(P.offs[f] + lr);
      }
      if (row == TW_NONE && P.status) atomicOr(P.status, 1);
    }
    sts32(rowidx_base + 4u * idx, row);
  }
  for (int idx = tid; idx < P.nd * P.N1PAD; idx += TF_THREADS) {
    const int j = idx / P.N1PAD, n = idx - j * P.N1PAD;
    const float w = n < P.N1 ? P.W1[((int64_t)m * K + j) * P.N1 + n] : 0.f;
    sts32(w1d_base + 4u * idx, __float_as_uint(w));
  }
  for (int n = tid; n < P.N1PAD; n += TF_THREADS) sts32(b1_base + 4u * n, __float_as_uint(n < P.N1 ? P.b1[n] : 0.f));
  __syncthreads();
  if (warp == 8) tmem_alloc_cols(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = lds32(tmem_slot);
  bool ok = true;

  if (warp < 8) {
    // ================================================================== producers
    const int c8 = tid & 7, rg = tid >> 3;  // rows rg + 32 i: (row & 7) == (rg & 7) for every i
    const uint32_t dst0 = (uint32_t)rg * 128u + ((uint32_t)(c8 ^ (rg & 7)) << 4);
    float4 S[KB][R], Q[KB][R];
#pragma unroll
    for (int blk = 0; blk < KB; ++blk)
#pragma unroll
      for (int i = 0; i < R; ++i) S[blk][i] = Q[blk][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float bias_acc = 0.f, lin_acc = 0.f;  // threads 0..127: row tid

    auto issue = [&](int f) {
      const int s = f % TF_STAGES;
      ok = ok && mbar_wait(empty(s), (((uint32_t)(f / TF_STAGES)) & 1u) ^ 1u);
      const uint32_t xs = base + (uint32_t)s * XT + dst0;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int r = rg + 32 * i;
        const uint32_t row = lds32(rowidx_base + 4u * (uint32_t)(r * m + f));
        const bool live = row != TW_NONE;
        const float* src = P.table[live ? (row >> TF_ROW_BITS) : 0u] + (int64_t)(live ? (row & TF_ROW_MASK) : 0u) * K + 4 * c8;
#pragma unroll
        for (int blk = 0; blk < KB; ++blk)
          cp_async16(xs + (uint32_t)blk * 16384u + (uint32_t)i * 4096u, src + 32 * blk, live ? 16u : 0u);
      }
      if (P.scal[0] && tid < TF_ROWS) {
        const uint32_t row = lds32(rowidx_base + 4u * (uint32_t)(tid * m + f));
        const bool live = row != TW_NONE;
        cp_async8(scal_base + (uint32_t)s * 1024u + 8u * tid,
                  P.scal[live ? (row >> TF_ROW_BITS) : 0u] + 2 * (int64_t)(live ? (row & TF_ROW_MASK) : 0u), live ? 8u : 0u);
      }
    };

    for (int f = 0; f < TF_DEPTH && f < m; ++f) {
      issue(f);
      cp_async_commit();
    }
    int t = 0;  // 32-column blocks handed to the tensor core so far
    for (int f = 0; f < m; ++f) {
      cp_async_wait<TF_DEPTH - 1>();  // all but the newest TF_DEPTH-1 groups: field f has landed (this thread's chunks)
      const int s = f % TF_STAGES;
#pragma unroll
      for (int blk = 0; blk < KB; ++blk, ++t) {
        // the remainder tile alternates between two buffers: block t+1 is prepared while the MMAs of block t run
        const int ls = t & 1;
        ok = ok && mbar_wait(lo_free(ls), (((uint32_t)(t >> 1)) & 1u) ^ 1u);
        const uint32_t xs = base + (uint32_t)s * XT + (uint32_t)blk * 16384u + dst0;
        const uint32_t ld = lo_base + (uint32_t)ls * 16384u + dst0;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float4 v = lds128(xs + (uint32_t)i * 4096u);
          S[blk][i].x += v.x; S[blk][i].y += v.y; S[blk][i].z += v.z; S[blk][i].w += v.w;
          Q[blk][i].x += v.x * v.x; Q[blk][i].y += v.y * v.y; Q[blk][i].z += v.z * v.z; Q[blk][i].w += v.w * v.w;
          sts128(ld + (uint32_t)i * 4096u, trunc_lo4(v));
          if (P.x) {
            const int64_t b = b0 + rg + 32 * i;
            if (b < P.B) st4(P.x + b * P.ld + (int64_t)f * K + 32 * blk + 4 * c8, v);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(lo_ready(ls));
      }
      if (P.scal[0] && tid < TF_ROWS) {
        const float2 sv = lds64f(scal_base + (uint32_t)s * 1024u + 8u * tid);
        bias_acc += sv.x;
        lin_acc += sv.y;
      }
      if (f + TF_DEPTH < m) issue(f + TF_DEPTH);
      cp_async_commit();  // one (possibly empty) group per field keeps the wait count uniform
    }

    // ---- FM second order, field sums, first-order logits
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float second = 0.f;
#pragma unroll
      for (int blk = 0; blk < KB; ++blk) {
        const float4 s4 = S[blk][i], q4 = Q[blk][i];
        second += 0.5f * (s4.x * s4.x - q4.x) + 0.5f * (s4.y * s4.y - q4.y) + 0.5f * (s4.z * s4.z - q4.z) +
                  0.5f * (s4.w * s4.w - q4.w);
      }
      second = group_sum<8>(second);
      const int r = rg + 32 * i;
      if (c8 == 0) sts32(rowsum_base + 4u * r, __float_as_uint(second));
      const int64_t b = b0 + r;
      if (P.sum_out && b < P.B) {
#pragma unroll
        for (int blk = 0; blk < KB; ++blk) st4(P.sum_out + b * K + 32 * blk + 4 * c8, S[blk][i]);
      }
    }
    named_bar_sync(1, TF_PRODUCERS);
    if (tid < TF_ROWS) {
      const int64_t b = b0 + tid;
      if (b < P.B) {
        for (int j = 0; j < P.nd; ++j) {
          const float dv = P.dense[b * P.nd + j];
          if (P.x) P.x[b * P.ld + (int64_t)m * K + j] = dv;
          if (P.lin_dense) lin_acc += dv * P.lin_dense[(int64_t)j * P.lin_dense_stride];
        }
        if (P.fm_out) P.fm_out[b] = bias_acc + __uint_as_float(lds32(rowsum_base + 4u * tid));
        if (P.lin_out) P.lin_out[b] = lin_acc;
      }
    }

    // ================================================================== epilogue: y1 = sum of accumulators + b1 + dense part
    ok = ok && mbar_wait(accum_full, 0);
    tc_fence_after();
    const int q = warp & 3, h = warp >> 2;
    const int r = 32 * q + lane;
    const int64_t b = b0 + r;
    const int nacc = m < P.NACC ? m : P.NACC;
    for (int cc = h; cc < P.N1PAD / 16; cc += 2) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      for (int a = 0; a < nacc; ++a) {
        // an accumulator is [a_hi*w_hi + a_lo*w_hi | a_hi*w_lo]: columns n and N1PAD + n belong together
        uint32_t v[16], w[16];
        const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a * 2 * P.N1PAD + 16 * cc);
        tmem_ld16(ta, v);
        tmem_ld16(ta + (uint32_t)P.N1PAD, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v[j]) + __uint_as_float(w[j]);
      }
      if (b < P.B) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(lds32(b1_base + 4u * (16 * cc + j)));
        for (int jd = 0; jd < P.nd; ++jd) {
          const float dv = P.dense[b * P.nd + jd];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            acc[j] = fmaf(dv, __uint_as_float(lds32(w1d_base + 4u * (jd * P.N1PAD + 16 * cc + j))), acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = 16 * cc + j;
          if (n < P.N1) P.y1[b * P.N1 + n] = acc[j];
        }
      }
    }
  } else if (warp == 8) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_wide = umma_idesc_tf32(128, 2 * P.N1PAD);  // a_hi x [w_hi | w_lo]
      const uint32_t idesc_hi = umma_idesc_tf32(128, P.N1PAD);        // a_lo x w_hi
      const uint32_t wblk = 2u * (uint32_t)P.N1PAD * 128u;            // one block of a field image: hi rows, lo rows
      int t = 0;
      for (int f = 0; f < m; ++f) {
        const int s = f % TF_STAGES, ws = f % TF_WSLOTS;
        ok = ok && mbar_wait(wfull(ws), ((uint32_t)(f / TF_WSLOTS)) & 1u);
        const uint32_t acc = tmem_base + (uint32_t)((f % P.NACC) * 2 * P.N1PAD);
        const uint32_t xa = base + (uint32_t)s * XT;
        const uint32_t wa = w_base + (uint32_t)ws * WB;
#pragma unroll
        for (int blk = 0; blk < KB; ++blk, ++t) {
          const int ls = t & 1;
          ok = ok && mbar_wait(lo_ready(ls), ((uint32_t)(t >> 1)) & 1u);  // block landed (all producers) + remainder tile
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = umma_desc(xa + (uint32_t)blk * 16384u + (uint32_t)ks * 32u);
            const uint64_t a_lo = umma_desc(lo_base + (uint32_t)ls * 16384u + (uint32_t)ks * 32u);
            const uint64_t b_w = umma_desc(wa + (uint32_t)blk * wblk + (uint32_t)ks * 32u);
            const uint32_t accumulate = (f >= P.NACC || blk > 0 || ks > 0) ? 1u : 0u;
            umma_tf32(acc, a_hi, b_w, idesc_wide, accumulate);
            umma_tf32(acc, a_lo, b_w, idesc_hi, 1u);
          }
          umma_commit(lo_free(ls));
        }
        umma_commit(empty(s));
        umma_commit(wempty(ws));
      }
      umma_commit(accum_full);
    }
  } else {
    // ================================================================== W1 image loader
    if (lane == 0) {
      for (int f = 0; f < m; ++f) {
        const int ws = f % TF_WSLOTS;
        ok = ok && mbar_wait(wempty(ws), (((uint32_t)(f / TF_WSLOTS)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(wfull(ws), WB);
        bulk_g2s(w_base + (uint32_t)ws * WB, P.wpack + (size_t)f * (WB / 4), WB, wfull(ws));
      }
    }
  }
  if (!ok && P.status) atomicOr(P.status, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_free_cols(tmem_base, P.tmem_cols);
  }
}

static size_t tower_fwd_smem(int KB, int m, int nd, int N1PAD) {
  const size_t XT = (size_t)KB * 16384;
  const size_t WB = 2 * (size_t)KB * N1PAD * 128;
  size_t s = 1024;  // alignment slack
  s += TF_STAGES * XT + 2 * 16384 + TF_WSLOTS * WB + TF_STAGES * 1024;
  s += ((size_t)TF_ROWS * m * 4 + 15) & ~(size_t)15;
  s += (size_t)nd * N1PAD * 4 + (size_t)N1PAD * 4 + TF_ROWS * 4 + 8;
  s += 8 * (TF_STAGES + 2 * TF_WSLOTS + 6);
  return s;
}

static int tower_n1pad(int N1) { return (N1 + 15) / 16 * 16; }

}  // namespace rm

extern "C" {

int rm_tower_supported(int32_t m, int32_t k, int32_t n_dense, int32_t N1) {
  using namespace rm;
  if (!(k == 32 || k == 64) || m < 1 || m > 64 || n_dense < 0 || n_dense > 64 || N1 < 1 || N1 > 64) return 0;
  return tower_fwd_smem(k / 32, m, n_dense, tower_n1pad(N1)) <= 227 * 1024 ? 1 : 0;
}

size_t rm_tower_fwd_workspace_bytes(int32_t m, int32_t k, int32_t N1) {
  if (m <= 0 || k <= 0 || N1 <= 0) return 256;
  return 256 + (size_t)m * 2 * (k / 32) * rm::tower_n1pad(N1) * 128;
}

static int tower_fwd_impl(const float* const* tables, const float* const* scals, int W, const int64_t* feat_sizes,
                          const int64_t* offs, const int64_t* ids, const float* dense, const float* lin_dense,
                          int32_t lin_dense_stride, int32_t n_dense, const float* W1, const float* b1, int32_t N1,
                          int64_t B, int32_t m, int32_t k, float* x, int64_t ld, float* y1, float* fm_out,
                          float* lin_out, float* sum_out, int32_t* status, void* workspace, size_t workspace_bytes,
                          void* stream, const char* who) {
  using namespace rm;
  RM_CHECK_ARG(tables && offs && ids && W1 && b1 && y1 && workspace, "null pointer");
  RM_CHECK_ARG(B >= 0 && m > 0 && k > 0 && n_dense >= 0 && N1 > 0, "bad shape");
  RM_CHECK_ARG(n_dense == 0 || dense, "dense pointer missing");
  RM_CHECK_ARG(!x || ld >= (int64_t)m * k + n_dense, "ld smaller than m*k+n_dense");
  RM_UNSUPPORTED(rm_tower_supported(m, k, n_dense, N1), "tower kernels need k in {32, 64}, m <= 64, N1 <= 64");
  RM_UNSUPPORTED(W >= 1 && W <= TF_MAX_PEERS && true, "world size must be <= 8");
  RM_CHECK_ARG(W == 1 || feat_sizes, "feat_sizes missing");
  RM_UNSUPPORTED((!x || (aligned16(x) && ld % 4 == 0)) && (!sum_out || aligned16(sum_out)) && aligned16(workspace),
                 "tower forward needs 16-byte aligned rows");
  const size_t need = rm_tower_fwd_workspace_bytes(m, k, N1);
  if (workspace_bytes < need) {
    set_error("%s: workspace %zu < required %zu", who, workspace_bytes, need);
    return RM_E_WORKSPACE;
  }
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int KB = k / 32;
  const int N1PAD = tower_n1pad(N1);
  uint32_t* wpack = (uint32_t*)((char*)workspace + 256);
  tower_pack_w1t_kernel<<<grid_for((int64_t)m * 2 * KB * N1PAD * 8, 256, 8), 256, 0, st>>>(W1, m, k, N1, N1PAD, wpack);
  RM_LAUNCH_CHECK();
  TowerFwdParams P;
  for (int r = 0; r < TF_MAX_PEERS; ++r) {
    P.table[r] = r < W ? tables[r] : nullptr;
    P.scal[r] = (r < W && scals) ? scals[r] : nullptr;
    RM_CHECK_ARG(r >= W || (P.table[r] && aligned16(P.table[r])), "null / misaligned table");
    RM_CHECK_ARG(r >= W || !scals || (P.scal[r] && (reinterpret_cast<uintptr_t>(P.scal[r]) & 7) == 0),
                 "null / misaligned k=1 table");
  }
  P.W = W;
  P.wshift = world_shift(W);
  P.feat_sizes = feat_sizes;
  P.offs = offs; P.ids = ids; P.dense = dense; P.lin_dense = lin_dense;
  P.lin_dense_stride = lin_dense_stride; P.wpack = wpack; P.W1 = W1; P.b1 = b1; P.x = x; P.ld = ld; P.y1 = y1;
  P.fm_out = fm_out; P.lin_out = lin_out; P.sum_out = sum_out; P.status = status; P.B = B; P.m = m; P.nd = n_dense;
  P.N1 = N1; P.N1PAD = N1PAD;
  int nacc = 512 / (2 * N1PAD);  // an accumulator is 2*N1PAD columns wide: [.. x w_hi | .. x w_lo]
  if (nacc > 8) nacc = 8;
  P.NACC = nacc;
  uint32_t cols = 32;
  while (cols < (uint32_t)(nacc * 2 * N1PAD)) cols <<= 1;
  P.tmem_cols = cols;
  const size_t smem = tower_fwd_smem(KB, m, n_dense, N1PAD);
  const int grid = (int)ceil_div(B, TF_ROWS);
  if (KB == 1) {
    RM_SMEM_ATTR_ONCE(smem, tower_fwd_kernel<1>);
    tower_fwd_kernel<1><<<grid, TF_THREADS, smem, st>>>(P);
  } else {
    RM_SMEM_ATTR_ONCE(smem, tower_fwd_kernel<2>);
    tower_fwd_kernel<2><<<grid, TF_THREADS, smem, st>>>(P);
  }
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_tower_fwd(const float* table, const float* scal, const int64_t* table_offsets, const int64_t* ids,
                 const float* dense, const float* lin_dense, int32_t lin_dense_stride, int32_t n_dense,
                 const float* W1, const float* b1, int32_t N1, int64_t B, int32_t m, int32_t k, float* x, int64_t ld,
                 float* y1, float* fm_out, float* lin_out, float* sum_out, int32_t* status, void* workspace,
                 size_t workspace_bytes, void* stream) {
  const float* tabs[1] = {table};
  const float* scs[1] = {scal};
  return tower_fwd_impl(tabs, scal ? scs : nullptr, 1, nullptr, table_offsets, ids, dense, lin_dense, lin_dense_stride,
                        n_dense, W1, b1, N1, B, m, k, x, ld, y1, fm_out, lin_out, sum_out, status, workspace,
                        workspace_bytes, stream, "rm_tower_fwd");
}

/* Row-sharded tables over NVLink peer memory: row `id` of field f is read from tables[id % W] +
 * (local_offsets[f] + id / W) * k (scals alike).  tables / scals are HOST arrays of W device pointers. */
int rm_tower_fwd_p2p(const float* const* tables, const float* const* scals, int32_t W, const int64_t* feat_sizes,
                     const int64_t* local_offsets, const int64_t* ids, const float* dense, const float* lin_dense,
                     int32_t lin_dense_stride, int32_t n_dense, const float* W1, const float* b1, int32_t N1, int64_t B,
                     int32_t m, int32_t k, float* y1, float* fm_out, float* lin_out, float* sum_out, int32_t* status,
                     void* workspace, size_t workspace_bytes, void* stream) {
  return tower_fwd_impl(tables, scals, W, feat_sizes, local_offsets, ids, dense, lin_dense, lin_dense_stride, n_dense, W1,
                        b1, N1, B, m, k, nullptr, 0, y1, fm_out, lin_out, sum_out, status, workspace, workspace_bytes,
                        stream, "rm_tower_fwd_p2p");
}

}  // extern "C"
