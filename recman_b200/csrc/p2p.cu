// (e) Row-sharded embedding tables over NVLink peer memory: no all-to-all, no pack/unpack pass.
//
// Row r of a table lives on rank r mod W at local row r div W (W a power of two <= 8: one NVSwitch box).  Every
// rank maps its peers' tables (cudaIpc handles, exchanged once) and the fused front end reads each row straight from
// its owner - local HBM for its own rows, NVLink loads for the others - so the lookup, the [embeds | dense] row
// buffer, the FM sums and the first-order term are ONE kernel on every rank and the NVLink transfers overlap the
// local HBM work warp by warp.  Nothing in the single-process reference corresponds to this file
// (recman/tf/core/layers.py:238-261 is the lookup it shards).
//
// Gradients travel the other way through per-rank gradient-row buffers G[b*m, k+4] that the owners read over NVLink
// inside the deterministic segmented reduce (scatter.cu: rm_shard_plan / rm_segment_reduce_p2p).
#include "common.cuh"

#define RM_MAX_PEERS 8

namespace rm {

struct PeerTables {
  const float* tab[RM_MAX_PEERS];
  const float* bias[RM_MAX_PEERS];
  const float* lin[RM_MAX_PEERS];
};

// one row group (LPR lanes) per sample; identical to gather_fm_kernel (gather.cu) except for the row address
template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) gather_fm_p2p_kernel(
    const PeerTables pt, int W, int wshift, const int64_t* __restrict__ feat_sizes, const int64_t* __restrict__ local_offs,
    const int64_t* __restrict__ ids, const float* __restrict__ dense, const float* __restrict__ lin_dense, int n_dense,
    int64_t B, int m, int k, float* __restrict__ x, int64_t ld, float* __restrict__ fm_out, float* __restrict__ lin_out,
    float* __restrict__ sum_out, int32_t* status) {
  const int lir = threadIdx.x % LPR;
  const int k4 = k >> 2;
  const bool col_ok = lir < k4;
  const bool has_bias = pt.bias[0] != nullptr, has_lin = pt.lin[0] != nullptr;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
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
        int owner[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int f = f0 + u;
          ok[u] = false;
          row[u] = 0;
          owner[u] = 0;
          if (f < m) {
            const int64_t id = my_ids[f];
            ok[u] = (id >= 0) && (id < feat_sizes[f]);
            if (ok[u]) {
              int64_t lr;
              shard_of(id, W, wshift, owner[u], lr);
              row[u] = local_offs[f] + lr;
            } else if (lir == 0 && status) {
              atomicOr(status, 1);
            }
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
            if (col_ok) v[u] = ldg_stream4(pt.tab[owner[u]] + row[u] * (int64_t)k + 4 * lir);
            if (lir == (u % LPR)) {  // spread the k=1 lookups over the lanes of the group
              if (has_bias) bv[u] = ldg_stream1(pt.bias[owner[u]] + row[u]);
              if (has_lin) lv[u] = ldg_stream1(pt.lin[owner[u]] + row[u]);
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

static inline int pow2ceil_p(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace rm

extern "C" {

int rm_p2p_alloc(size_t bytes, void** ptr, uint8_t* handle64) {
  using namespace rm;
  RM_CHECK_ARG(ptr && handle64 && bytes > 0, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  RM_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("rm_p2p_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    return (int)e;
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

int rm_p2p_open(const uint8_t* handle64, void** ptr) {
  using namespace rm;
  RM_CHECK_ARG(ptr && handle64, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  RM_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int rm_p2p_close(void* ptr) {
  using namespace rm;
  if (ptr) RM_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int rm_p2p_free(void* ptr) {
  using namespace rm;
  if (ptr) RM_CUDA(cudaFree(ptr));
  return 0;
}

int rm_gather_fm_fwd_p2p(const float* const* tables, const float* const* bias_tables, const float* const* lin_tables,
                         int32_t W, const int64_t* feat_sizes, const int64_t* local_offsets, const int64_t* ids,
                         const float* dense, const float* lin_dense, int32_t n_dense, int64_t B, int32_t m, int32_t k,
                         float* x, int64_t ld, float* fm_out, float* lin_out, float* sum_out, int32_t* status,
                         void* stream) {
  using namespace rm;
  RM_CHECK_ARG(tables && feat_sizes && local_offsets && ids && x, "null pointer");
  RM_CHECK_ARG(B >= 0 && m > 0 && k > 0 && n_dense >= 0, "bad shape");
  RM_CHECK_ARG(n_dense == 0 || dense, "dense pointer missing");
  RM_CHECK_ARG(ld >= (int64_t)m * k + n_dense, "ld smaller than m*k+n_dense");
  RM_UNSUPPORTED(W >= 1 && W <= RM_MAX_PEERS && true, "world size must be <= 8");
  RM_UNSUPPORTED(k % 4 == 0 && k <= 128, "fused front end needs k % 4 == 0 and k <= 128");
  RM_UNSUPPORTED(ld % 4 == 0 && aligned16(x) && (!sum_out || aligned16(sum_out)),
                 "fused front end needs 16-byte aligned rows (ld % 4 == 0)");
  if (B == 0) return 0;
  PeerTables pt;
  for (int r = 0; r < RM_MAX_PEERS; ++r) {
    pt.tab[r] = r < W ? tables[r] : nullptr;
    pt.bias[r] = (r < W && bias_tables) ? bias_tables[r] : nullptr;
    pt.lin[r] = (r < W && lin_tables) ? lin_tables[r] : nullptr;

    RM_CHECK_ARG(r >= W || (pt.tab[r] && aligned16(pt.tab[r])), "null / misaligned peer table");
    RM_CHECK_ARG(r >= W || !bias_tables || pt.bias[r], "null peer bias table");
    RM_CHECK_ARG(r >= W || !lin_tables || pt.lin[r], "null peer linear table");
  }
  const int wshift = world_shift(W);
  cudaStream_t st = (cudaStream_t)stream;
  const int lpr = pow2ceil_p(k / 4);
  // (a deep-queue cp.async variant of this kernel was measured SLOWER at W = 2 - 0.72 ms vs 0.54 ms: the NVLink reads are
  // not bound by registers held per load in flight - and was removed; profiles/README.md)
  const int grid = grid_for(B, 256 / (lpr > 32 ? 32 : lpr), 8);
#define RM_GP(L)                                                                                                      \
  case L:                                                                                                             \
    gather_fm_p2p_kernel<L, 4, 4><<<grid, 256, 0, st>>>(pt, W, wshift, feat_sizes, local_offsets, ids, dense, lin_dense,  \
                                                        n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status);        \
    break
  switch (lpr) {
    RM_GP(1);
    RM_GP(2);
    RM_GP(4);
    RM_GP(8);
    RM_GP(16);
    default:
      gather_fm_p2p_kernel<32, 4, 4><<<grid, 256, 0, st>>>(pt, W, wshift, feat_sizes, local_offsets, ids, dense, lin_dense,
                                                           n_dense, B, m, k, x, ld, fm_out, lin_out, sum_out, status);
  }
#undef RM_GP
  RM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
