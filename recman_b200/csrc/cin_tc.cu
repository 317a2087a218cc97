// K5 tensor-core path: CIN layer forward on tcgen05 (sm_100a), TMEM accumulators, hand-written PTX.
//
// recman/tf/core/layers.py:711-739:  F[(b,d), n] = act( sum_{p,q} x0[b,p,d]*xk[b,q,d] * W[(p,q), n] + bias[n] )
// is a GEMM with M = B*D rows, N <= 256 columns and K = m*H, whose A operand Z = x0 (x) xk is never
// materialised (3 GB at config C3): it is synthesised tile by tile straight into the swizzled shared-memory
// layout the tensor core reads.
//
// CTA = 256 rows of M x all N columns, 1 CTA per SM, 320 threads:
//   warps 0-7  producers: thread r owns row r.  It keeps x0[b, 0..m-1, d] in registers, walks q, multiplies by
//              xk[b,q,d] and writes 16-byte chunks of the A tile (K-major, SWIZZLE_128B) into a 3-stage ring;
//              afterwards the same warps are the epilogue (tcgen05.ld -> +bias -> act -> global).
//   warp  8    lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=NPAD, K=8) into two TMEM
//              accumulators (rows 0-127 at column 0, rows 128-255 at column 256); the warp owns TMEM alloc/free.
//   warp  9    lane 0 streams the pre-packed W stage images with cp.async.bulk (1-D TMA) + mbarrier tx-count.
//
// The reduction index is re-ordered to k'' = q*MPAD + p (MPAD = m rounded up to 4; W rows are permuted and
// zero-padded to match by cin_pack_w_kernel) so that a thread's x0 register index is a compile-time constant.
//
// Precision.  tcgen05 has no fp32 MMA.  RM_CIN_TF32: one pass, operands rounded to tf32 (rna).
// RM_CIN_3XTF32 (parity mode): a = a_hi + a_lo, w = w_hi + w_lo and D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
// with fp32 accumulation in TMEM (the dropped a_lo*w_lo term is 2^-22 relative).  A 128-byte smem row then
// holds 16 k'' of "hi" in its first 64 bytes and the same 16 k'' of "lo" in the last 64, so both modes use the
// same SWIZZLE_128B tiles and descriptors.
//
// TMEM accumulation rounds toward zero - a bias that grows linearly with the number of accumulate steps (measured
// 3.3e-5 of max|F| at k'' = 2800).  In parity mode the reduction is therefore cut into splits of <= TC_SPLIT_K k'':
// CTA (tile, split) accumulates its k'' range in TMEM and writes the raw partial sums; cin_splitk_finish_kernel adds
// the splits in order in fp32 round-to-nearest (deterministic), then bias + activation.  One split = the fused epilogue.
//
// (Measured alternative: one CTA running its tile's splits one after the other, adding each drained accumulator to an
// L2-resident running sum - no partial tensors through HBM, but the producer warps are also the epilogue warps, so
// synthesis and MMAs stall at every split boundary: forward 3.12 ms per C3 step against 2.04 ms with (tile, split) CTAs.)
//
// Every mbarrier wait is bounded: a pipeline bug sets *status and lets the kernel drain instead of hanging.
#include "tc_common.cuh"

namespace rm {

constexpr int TC_BM = 256;          // rows per CTA (two M=128 accumulators)
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = TC_BM * 128;  // A tile: 256 rows x 128 B
constexpr int TC_THREADS = 320;
constexpr int TC_SPLIT_K = 512;     // parity mode: k'' accumulated in TMEM per CTA (truncation bias < 1e-5 of max|F|)

// ------------------------------------------------------------------------------------------------ W packing
// Stage image `it` = NPAD rows x 128 B, 16-byte chunk c of row n stored at chunk (c ^ (n & 7)).
//   TF32 : chunk c holds k'' = it*32 + 4c .. +3                       (tf32-rounded)
//   3x   : chunk c<4 holds hi of k'' = it*16 + 4c .. +3, chunk c>=4 the matching lo
// with k'' = q*MPAD + p  ->  W[(p*H + q), n]  (zero for p >= m, q >= H, n >= N).
__global__ void __launch_bounds__(256) cin_pack_w_kernel(const float* __restrict__ W, int m, int H, int N, int NPAD,
                                                         int MPAD, int split3, int n_stages,
                                                         uint32_t* __restrict__ out) {
  const int64_t total = (int64_t)n_stages * NPAD * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int64_t rowi = i >> 3;
    const int n = (int)(rowi % NPAD);
    const int it = (int)(rowi / NPAD);
    const bool lo = split3 && c >= 4;
    const int kbase = split3 ? it * 16 + 4 * (c & 3) : it * 32 + 4 * c;
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = kbase + j;
      const int q = kk / MPAD, p = kk - q * MPAD;
      float w = 0.f;
      if (q < H && p < m && n < N) w = W[((int64_t)p * H + q) * N + n];
      const uint32_t hi = f32_to_tf32(w);
      v[j] = lo ? f32_to_tf32(w - __uint_as_float(hi)) : hi;
    }
    uint32_t* dst = out + ((int64_t)it * NPAD + n) * 32 + ((c ^ (n & 7)) << 2);
    *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------------------ main kernel
struct TcParams {
  const float* x0;
  int64_t bs0;
  const float* xk;
  int64_t bsk;
  const uint32_t* wpack;
  const float* bias;
  float* out;
  float* pre;
  int32_t* status;
  float* partial;  // n_split > 1: [n_split][B, N, D] raw partial sums (no bias / activation)
  int64_t Mrows;   // B*D
  int m, H, D, N, NPAD, act, n_stages;
  int n_split, stages_per_split;
};

// out = act(bias + sum over splits, in order) - fp32 round-to-nearest, fixed order
__global__ void __launch_bounds__(256) cin_splitk_finish_kernel(const float* __restrict__ partial, int n_split,
                                                                int64_t total, int N, int D,
                                                                const float* __restrict__ bias, int act,
                                                                float* __restrict__ out, float* __restrict__ pre) {
  const int64_t total4 = total >> 2;  // D % 4 == 0 is not required: vector path only when total % 4 == 0 and D % 4 == 0
  const bool vec = (total & 3) == 0 && (D & 3) == 0;
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 acc = __ldcs(reinterpret_cast<const float4*>(partial) + i);
      for (int s = 1; s < n_split; ++s) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(partial + (int64_t)s * total) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const float bv = __ldg(bias + (int)(((i << 2) / D) % N));
      acc.x += bv; acc.y += bv; acc.z += bv; acc.w += bv;
      if (pre) reinterpret_cast<float4*>(pre)[i] = acc;
      reinterpret_cast<float4*>(out)[i] =
          make_float4(tc_act(acc.x, act), tc_act(acc.y, act), tc_act(acc.z, act), tc_act(acc.w, act));
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      float acc = partial[i];
      for (int s = 1; s < n_split; ++s) acc += partial[(int64_t)s * total + i];
      acc += __ldg(bias + (int)((i / D) % N));
      if (pre) pre[i] = acc;
      out[i] = tc_act(acc, act);
    }
  }
}

template <int MP4, bool SPLIT3>
__global__ void __launch_bounds__(TC_THREADS, 1) cin_fwd_tc_kernel(const TcParams P) {
  constexpr int CPS = SPLIT3 ? 4 : 8;  // 16-byte k-chunks of one row produced per stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t b_bytes = (uint32_t)P.NPAD * 128u;
  const uint32_t stage_bytes = (TC_A_BYTES + b_bytes + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC_STAGES * stage_bytes;
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto full_b = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  auto empty = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + s); };
  const uint32_t accum_full = bar_base + 8u * (3 * TC_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (3 * TC_STAGES + 1);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_a(s), 256);
      mbar_init(full_b(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 8) {  // one warp allocates all 512 TMEM columns (1 CTA per SM by construction: > 114 KB smem)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // this CTA's slice of the reduction: stages [stage0, stage0 + n_stages) of the whole K, tile blockIdx.x
  const int split = (int)blockIdx.y;
  const int stage0 = split * P.stages_per_split;
  const int n_stages = min(P.stages_per_split, P.n_stages - stage0);
  const int g0 = stage0 * CPS, g1 = (stage0 + n_stages) * CPS;  // 16-byte k-chunks [g0, g1) of a row (g = q*MP4 + c)
  bool ok = true;

  if (warp < 8) {
    // ===================================================================== producers: A tile synthesis
    const int r = tid;
    const int64_t R = (int64_t)blockIdx.x * TC_BM + r;
    const bool valid = R < P.Mrows;
    const int64_t b = valid ? R / P.D : 0;
    const int d = valid ? (int)(R - b * P.D) : 0;
    float x0r[MP4 * 4];
#pragma unroll
    for (int j = 0; j < MP4 * 4; ++j) x0r[j] = (valid && j < P.m) ? P.x0[b * P.bs0 + (int64_t)j * P.D + d] : 0.f;
    const float* xkp = P.xk + b * P.bsk + d;
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t rx = (uint32_t)(r & 7);
    int cnt = 0;  // chunks produced so far (same value in every producer thread)
    const int q_begin = g0 / MP4, q_end = min(P.H, (g1 + MP4 - 1) / MP4);
    for (int q = q_begin; q < q_end; ++q) {
      const float xkv = valid ? __ldg(xkp + (int64_t)q * P.D) : 0.f;
#pragma unroll
      for (int c = 0; c < MP4; ++c) {
        const int g = q * MP4 + c;
        if (g < g0 || g >= g1) continue;  // another split's chunk (only in the first / last q of the range)
        const int slot = cnt % CPS;
        const int it = cnt / CPS;
        const int s = it % TC_STAGES;
        if (slot == 0) ok = mbar_wait(empty(s), (((uint32_t)(it / TC_STAGES)) & 1u) ^ 1u) && ok;
        const float v0 = x0r[4 * c + 0] * xkv, v1 = x0r[4 * c + 1] * xkv, v2 = x0r[4 * c + 2] * xkv,
                    v3 = x0r[4 * c + 3] * xkv;
        uint4 hi, lo;
        split_trunc4(make_float4(v0, v1, v2, v3), hi, lo);
        uint8_t* arow = gen_base + (size_t)s * stage_bytes + row_off;
        *reinterpret_cast<uint4*>(arow + ((((uint32_t)slot) ^ rx) << 4)) = hi;
        if (SPLIT3) *reinterpret_cast<uint4*>(arow + ((((uint32_t)(slot + 4)) ^ rx) << 4)) = lo;
        if (slot == CPS - 1) {
          fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
          mbar_arrive(full_a(s));
        }
        ++cnt;
      }
    }
    if (cnt % CPS != 0) {  // zero-fill the unused k-chunks of the last stage (its W rows are zero too)
      const int it = cnt / CPS;
      const int s = it % TC_STAGES;
      uint8_t* arow = gen_base + (size_t)s * stage_bytes + row_off;
      for (int slot = cnt % CPS; slot < CPS; ++slot) {
        *reinterpret_cast<uint4*>(arow + ((((uint32_t)slot) ^ rx) << 4)) = make_uint4(0, 0, 0, 0);
        if (SPLIT3) *reinterpret_cast<uint4*>(arow + ((((uint32_t)(slot + 4)) ^ rx) << 4)) = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
      mbar_arrive(full_a(s));
    }

    // ===================================================================== epilogue: TMEM -> registers -> global
    ok = mbar_wait(accum_full, 0) && ok;
    tc_fence_after();
    const int h = warp >> 2, quad = warp & 3;
    const int64_t Re = (int64_t)blockIdx.x * TC_BM + h * 128 + quad * 32 + lane;
    const bool evalid = Re < P.Mrows;
    const int64_t be = evalid ? Re / P.D : 0;
    const int de = evalid ? (int)(Re - be * P.D) : 0;
    const int64_t obase = be * (int64_t)P.N * P.D + de;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * 256);
    for (int col0 = 0; col0 < P.NPAD; col0 += 16) {
      uint32_t a[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
            "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
          : "r"(taddr0 + (uint32_t)col0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (evalid) {
        if (P.n_split > 1) {  // raw partial sums; cin_splitk_finish_kernel adds the splits, bias and activation
          float* part = P.partial + (int64_t)split * P.Mrows * P.N;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = col0 + j;
            if (n < P.N) __stcs(part + obase + (int64_t)n * P.D, __uint_as_float(a[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = col0 + j;
            if (n < P.N) {
              const float v = __uint_as_float(a[j]) + __ldg(P.bias + n);
              const int64_t o = obase + (int64_t)n * P.D;
              if (P.pre) P.pre[o] = v;
              P.out[o] = tc_act(v, P.act);
            }
          }
        }
      }
    }
  } else if (warp == 8) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, P.NPAD);
      for (int it = 0; it < n_stages; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t par = ((uint32_t)(it / TC_STAGES)) & 1u;
        ok = mbar_wait(full_a(s), par) && ok;
        ok = mbar_wait(full_b(s), par) && ok;
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t b_addr = a_addr + TC_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < (SPLIT3 ? 2 : 4); ++ks) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t acc = tmem_base + (uint32_t)(h * 256);
            const uint32_t first = (it > 0 || ks > 0) ? 1u : 0u;
            const uint32_t ah = a_addr + (uint32_t)h * (128u * 128u) + (uint32_t)ks * 32u;
            const uint32_t bh = b_addr + (uint32_t)ks * 32u;
            umma_tf32(acc, umma_desc(ah), umma_desc(bh), idesc, first);
            if (SPLIT3) {
              umma_tf32(acc, umma_desc(ah + 64u), umma_desc(bh), idesc, 1u);  // a_lo * w_hi
              umma_tf32(acc, umma_desc(ah), umma_desc(bh + 64u), idesc, 1u);  // a_hi * w_lo
            }
          }
        }
        umma_commit(empty(s));  // arrives once every MMA issued so far has finished reading stage s
      }
      umma_commit(accum_full);
    }
  } else {
    // ===================================================================== W loader (1-D bulk copies)
    if (lane == 0) {
      for (int it = 0; it < n_stages; ++it) {
        const int s = it % TC_STAGES;
        ok = mbar_wait(empty(s), (((uint32_t)(it / TC_STAGES)) & 1u) ^ 1u) && ok;
        mbar_arrive_expect_tx(full_b(s), b_bytes);
        bulk_g2s(smem_base + (uint32_t)s * stage_bytes + TC_A_BYTES,
                 P.wpack + (size_t)(stage0 + it) * (b_bytes / 4), b_bytes, full_b(s));
      }
    }
  }
  if (!ok && P.status) atomicOr(P.status, 2);  // a bounded wait expired: results are invalid
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int tc_stages_for(int m, int H, int split3) {
  const int mp4 = (m + 3) / 4;
  const int cps = split3 ? 4 : 8;
  return (H * mp4 + cps - 1) / cps;
}
static int tc_npad(int N) { return (N + 15) / 16 * 16; }

bool cin_tc_supported(int64_t B, int m, int H, int D, int N) {
  (void)B;
  (void)H;
  (void)D;
  return m >= 1 && m <= 32 && N >= 1 && N <= 256;
}

// parity mode: number of k'' splits (one TMEM accumulation each) and stages per split
static void tc_split_for(int n_stages, int split3, int* n_split, int* stages_per_split) {
  const int k_per_stage = split3 ? 16 : 32;
  int ns = split3 ? (n_stages * k_per_stage + TC_SPLIT_K - 1) / TC_SPLIT_K : 1;
  if (ns < 1) ns = 1;
  const int sps = (n_stages + ns - 1) / ns;
  *n_split = (n_stages + sps - 1) / sps;
  *stages_per_split = sps;
}

// workspace: [status int32 x 64 (256 B)] [packed W stage images] [split-K partial sums (parity mode, K > TC_SPLIT_K)]
size_t cin_tc_fwd_workspace(int64_t B, int m, int H, int D, int N, int precision) {
  const int split3 = precision == RM_CIN_3XTF32;
  const int n_stages = tc_stages_for(m, H, split3);
  int n_split, sps;
  tc_split_for(n_stages, split3, &n_split, &sps);
  size_t bytes = 256 + align_up((size_t)n_stages * tc_npad(N) * 128, 256);
  if (n_split > 1) bytes += align_up((size_t)n_split * (size_t)B * N * D * sizeof(float), 256);
  return bytes;
}

template <int MP4>
static int launch_tc(const TcParams& P, bool split3, dim3 grid, size_t smem, cudaStream_t st) {
  if (split3) {
    RM_SMEM_ATTR_ONCE(smem, cin_fwd_tc_kernel<MP4, true>);
    cin_fwd_tc_kernel<MP4, true><<<grid, TC_THREADS, smem, st>>>(P);
  } else {
    RM_SMEM_ATTR_ONCE(smem, cin_fwd_tc_kernel<MP4, false>);
    cin_fwd_tc_kernel<MP4, false><<<grid, TC_THREADS, smem, st>>>(P);
  }
  RM_LAUNCH_CHECK();
  return 0;
}

int cin_fwd_tc(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* bias,
               int64_t B, int m, int H, int D, int N, int act, int precision, float* out, float* pre, void* workspace,
               size_t workspace_bytes, cudaStream_t st) {
  const bool split3 = precision == RM_CIN_3XTF32;
  const size_t need = cin_tc_fwd_workspace(B, m, H, D, N, precision);
  RM_CHECK_ARG(workspace != nullptr, "null workspace");
  if (workspace_bytes < need) {
    set_error("rm_cin_layer_fwd: workspace %zu < required %zu", workspace_bytes, need);
    return RM_E_WORKSPACE;
  }
  const int NPAD = tc_npad(N);
  const int MP4 = (m + 3) / 4;
  const int n_stages = tc_stages_for(m, H, split3);
  int32_t* status = (int32_t*)workspace;
  uint32_t* wpack = (uint32_t*)((char*)workspace + 256);
  RM_CUDA(cudaMemsetAsync(status, 0, 256, st));
  cin_pack_w_kernel<<<grid_for((int64_t)n_stages * NPAD * 8, 256, 8), 256, 0, st>>>(W, m, H, N, NPAD, MP4 * 4,
                                                                                    split3 ? 1 : 0, n_stages, wpack);
  RM_LAUNCH_CHECK();
  TcParams P;
  P.x0 = x0; P.bs0 = bs0; P.xk = xk; P.bsk = bsk; P.wpack = wpack; P.bias = bias; P.out = out; P.pre = pre;
  P.status = status; P.Mrows = B * (int64_t)D; P.m = m; P.H = H; P.D = D; P.N = N; P.NPAD = NPAD; P.act = act;
  P.n_stages = n_stages;
  tc_split_for(n_stages, split3 ? 1 : 0, &P.n_split, &P.stages_per_split);
  P.partial = P.n_split > 1 ? (float*)((char*)workspace + 256 + align_up((size_t)n_stages * NPAD * 128, 256)) : nullptr;
  const uint32_t stage_bytes = (uint32_t)((TC_A_BYTES + NPAD * 128 + 1023) / 1024 * 1024);
  const size_t smem = (size_t)TC_STAGES * stage_bytes + 8 * (3 * TC_STAGES + 2) + 1024;  // + alignment slack
  const dim3 grid((unsigned)ceil_div(P.Mrows, TC_BM), (unsigned)P.n_split);
  int rc;
  switch (MP4) {
    case 1: rc = launch_tc<1>(P, split3, grid, smem, st); break;
    case 2: rc = launch_tc<2>(P, split3, grid, smem, st); break;
    case 3: rc = launch_tc<3>(P, split3, grid, smem, st); break;
    case 4: rc = launch_tc<4>(P, split3, grid, smem, st); break;
    case 5: rc = launch_tc<5>(P, split3, grid, smem, st); break;
    case 6: rc = launch_tc<6>(P, split3, grid, smem, st); break;
    case 7: rc = launch_tc<7>(P, split3, grid, smem, st); break;
    default: rc = launch_tc<8>(P, split3, grid, smem, st); break;
  }
  if (rc || P.n_split == 1) return rc;
  const int64_t total = P.Mrows * N;
  cin_splitk_finish_kernel<<<grid_for(total / 4 + 1, 256, 8), 256, 0, st>>>(P.partial, P.n_split, total, N, D, bias, act,
                                                                            out, pre);
  RM_LAUNCH_CHECK();
  return 0;
}

}  // namespace rm
