// K5 backward on tcgen05 (sm_100a): the two GEMMs of one CIN layer's backward, Z never materialised.
//
//   dF = dout * act'(pre)                                          (cin_simt.cu: cin_dF_kernel, + dbias)
//   A:  dZ[(b,d), k''] = sum_n dF[(b,d), n] * W''[k'', n]          GEMM  M = B*D, N = k'' tile, K = n
//       dxk[b,q,d]  = sum_p dZ[(b,d),(q,p)] * x0[b,p,d]            contracted in the epilogue, straight from TMEM
//       dx0[b,p,d] += sum_q dZ[(b,d),(q,p)] * xk[b,q,d]
//   B:  dW''[k'', n]   = sum_{(b,d)} Z''[(b,d), k''] * dF[(b,d), n]  GEMM  M = k'' tile, N = n, K = (b,d)
//
// with k'' = q*MPAD + p as in the forward (cin_tc.cu).  Both kernels reuse the forward's machinery: one thread per
// operand row writing K-major SWIZZLE_128B tiles (hi|lo halves for 3xTF32), mbarrier ring, one MMA-issuing lane,
// TMEM accumulators, bounded waits.
//
// The tensor core accumulates fp32 with round-toward-zero (bias linear in the number of accumulate steps, see
// cin_tc.cu), so kernel B - whose reduction runs over all B*D rows - accumulates at most slab_rows = 512 rows in TMEM and
// the slabs are then summed in fp32 round-to-nearest, in slab order (deterministic) by cin_dw_unpermute_kernel.
#include "tc_common.cuh"

namespace rm {

constexpr int TB_STAGES = 3;
constexpr int TB_THREADS = 320;   // kernel A: 8 producer/epilogue warps + MMA warp + W loader warp
constexpr int TB_THREADS_B = 544;  // kernel B: 8 A-producer/epilogue warps + 8 B-producer warps + MMA warp

// ------------------------------------------------------------------------------------------------ pack W'' (kernel A)
// image (qg, st): NMMA rows (j -> k'' = qg*NMMA + j) x 128 B; chunk c of row j at (c ^ (j&7)).
//   3x : c<4 hi of n = st*16 + 4c + e ; c>=4 lo of n = st*16 + 4(c-4) + e        TF32: n = st*32 + 4c + e
__global__ void __launch_bounds__(256) cin_pack_wT_kernel(const float* __restrict__ W, int m, int H, int N, int MPAD,
                                                          int QG, int NMMA, int spq, int n_qg, int split3,
                                                          uint32_t* __restrict__ out) {
  const int64_t total = (int64_t)n_qg * spq * NMMA * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int64_t rowi = i >> 3;
    const int j = (int)(rowi % NMMA);
    const int64_t img = rowi / NMMA;
    const int st = (int)(img % spq);
    const int qg = (int)(img / spq);
    const int q = qg * QG + j / MPAD, p = j % MPAD;
    const bool lo = split3 && c >= 4;
    const int nbase = split3 ? st * 16 + 4 * (c & 3) : st * 32 + 4 * c;
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = nbase + e;
      float w = 0.f;
      if (q < H && p < m && n < N) w = W[((int64_t)p * H + q) * N + n];
      uint32_t hi, l;
      split_tf32(w, hi, l);
      v[e] = lo ? f32_to_tf32(__uint_as_float(l)) : hi;
    }
    uint32_t* dst = out + (img * NMMA + j) * 32 + ((c ^ (j & 7)) << 2);
    *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

struct TbParams {
  const float* x0;
  int64_t bs0;
  const float* xk;
  int64_t bsk;
  const float* dF;       // [B, N, D]
  const uint32_t* wpack;  // kernel A: W'' stage images
  float* dx0;             // [B, m, D]   (+=)
  float* dxk;             // [B, H, D]   (=), batch stride dbsk
  int64_t dbsk;
  float* partial;         // kernel B: [slabs, KPADT, NPAD]
  int32_t* status;
  int64_t Mrows;
  int m, H, D, N, NPAD, MPAD, QG, NMMA, spq, n_qg, KPADT, slab_rows;
};

// ------------------------------------------------------------------------------------------------ kernel A
template <int MP4, bool SPLIT3>
__global__ void __launch_bounds__(TB_THREADS, 1) cin_bwd_dx_tc_kernel(const TbParams P) {
  constexpr int KS = SPLIT3 ? 16 : 32;  // n's per stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = 256u * 128u;
  const uint32_t b_bytes = (uint32_t)P.NMMA * 128u;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TB_STAGES * stage_bytes;
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto full_b = [&](int s) { return bar_base + 8u * (TB_STAGES + s); };
  auto empty = [&](int s) { return bar_base + 8u * (2 * TB_STAGES + s); };
  const uint32_t accum_full = bar_base + 8u * (3 * TB_STAGES);
  const uint32_t tmem_empty = bar_base + 8u * (3 * TB_STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (3 * TB_STAGES + 2);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < TB_STAGES; ++s) {
      mbar_init(full_a(s), 256);
      mbar_init(full_b(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(accum_full, 1);
    mbar_init(tmem_empty, 256);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 8) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  bool ok = true;
  const int total_stages = P.n_qg * P.spq;

  if (warp < 8) {
    const int r = tid;
    const int64_t R = (int64_t)blockIdx.x * 256 + r;
    const bool valid = R < P.Mrows;
    const int64_t b = valid ? R / P.D : 0;
    const int d = valid ? (int)(R - b * P.D) : 0;
    float x0r[MP4 * 4], dx0r[MP4 * 4];
#pragma unroll
    for (int j = 0; j < MP4 * 4; ++j) {
      x0r[j] = (valid && j < P.m) ? P.x0[b * P.bs0 + (int64_t)j * P.D + d] : 0.f;
      dx0r[j] = 0.f;
    }
    const float* dFp = P.dF + (b * P.N) * (int64_t)P.D + d;  // + n*D
    const float* xkp = P.xk + b * P.bsk + d;                  // + q*D
    float* dxkp = P.dxk + b * P.dbsk + d;
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t rx = (uint32_t)(r & 7);
    const int h = warp >> 2, quad = warp & 3;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * 256);
    int it = 0;
    // software pipeline: the dF values of the next stage are in flight while this stage is split and stored (the tile
    // is the same for every q group, so "next" wraps around at the end of a group)
    float vn[KS];
    auto fetch = [&](int st) {
#pragma unroll
      for (int e = 0; e < KS; ++e) {
        const int n = st * KS + e;
        vn[e] = (valid && n < P.N) ? __ldg(dFp + (int64_t)n * P.D) : 0.f;
      }
    };
    fetch(0);
    for (int qg = 0; qg < P.n_qg; ++qg) {
      // ---- produce the dF tile stage by stage (identical for every q group; re-read from L2) ----
      for (int st = 0; st < P.spq; ++st, ++it) {
        const int s = it % TB_STAGES;
        float v[KS];
#pragma unroll
        for (int e = 0; e < KS; ++e) v[e] = vn[e];
        fetch(st + 1 < P.spq ? st + 1 : 0);
        ok = mbar_wait(empty(s), (((uint32_t)(it / TB_STAGES)) & 1u) ^ 1u) && ok;
        uint8_t* arow = gen_base + (size_t)s * stage_bytes + row_off;
#pragma unroll
        for (int c = 0; c < KS / 4; ++c) {
          uint4 hi, lo;
          split_trunc4(make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]), hi, lo);
          *reinterpret_cast<uint4*>(arow + ((((uint32_t)c) ^ rx) << 4)) = hi;
          if (SPLIT3) *reinterpret_cast<uint4*>(arow + ((((uint32_t)(c + 4)) ^ rx) << 4)) = lo;
        }
        fence_proxy_async_smem();
        mbar_arrive(full_a(s));
      }
      // ---- epilogue of this q group: contract dZ (TMEM) with x0 (registers) and xk, two q per TMEM round trip ----
      ok = mbar_wait(accum_full, (uint32_t)qg & 1u) && ok;
      tc_fence_after();
      constexpr int MPADc = MP4 * 4;
      for (int qi0 = 0; qi0 < P.QG; qi0 += 2) {  // QG is a multiple of 4
        const int q0 = qg * P.QG + qi0;
        if (q0 >= P.H) break;  // warp-uniform
        const bool two = q0 + 1 < P.H;
        const float xk0 = valid ? __ldg(xkp + (int64_t)q0 * P.D) : 0.f;
        const float xk1 = (valid && two) ? __ldg(xkp + (int64_t)(q0 + 1) * P.D) : 0.f;
        uint32_t a[2 * MPADc];
        tmem_ld_cols<2 * MPADc>(taddr0 + (uint32_t)(qi0 * MPADc), a);
        tmem_ld_wait();
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int j = 0; j < MPADc; ++j) {
          const float z0 = __uint_as_float(a[j]), z1 = __uint_as_float(a[MPADc + j]);
          acc0 = fmaf(z0, x0r[j], acc0);
          acc1 = fmaf(z1, x0r[j], acc1);
          dx0r[j] = fmaf(z0, xk0, dx0r[j]);
          dx0r[j] = fmaf(z1, xk1, dx0r[j]);
        }
        if (valid) {
          dxkp[(int64_t)q0 * P.D] = acc0;
          if (two) dxkp[(int64_t)(q0 + 1) * P.D] = acc1;
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);  // the accumulators may be overwritten by the next q group
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < MP4 * 4; ++j)
        if (j < P.m) {
          float* o = P.dx0 + (b * P.m + j) * (int64_t)P.D + d;  // this CTA owns these rows
          *o += dx0r[j];
        }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, P.NMMA);
      int it = 0;
      for (int qg = 0; qg < P.n_qg; ++qg) {
        ok = mbar_wait(tmem_empty, ((uint32_t)qg & 1u) ^ 1u) && ok;  // epilogue of the previous group drained TMEM
        tc_fence_after();
        for (int st = 0; st < P.spq; ++st, ++it) {
          const int s = it % TB_STAGES;
          const uint32_t par = ((uint32_t)(it / TB_STAGES)) & 1u;
          ok = mbar_wait(full_a(s), par) && ok;
          ok = mbar_wait(full_b(s), par) && ok;
          tc_fence_after();
          const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
          const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
          for (int ks = 0; ks < (SPLIT3 ? 2 : 4); ++ks) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t acc = tmem_base + (uint32_t)(h * 256);
              const uint32_t accumulate = (st > 0 || ks > 0) ? 1u : 0u;
              const uint32_t ah = a_addr + (uint32_t)h * (128u * 128u) + (uint32_t)ks * 32u;
              const uint32_t bh = b_addr + (uint32_t)ks * 32u;
              umma_tf32(acc, umma_desc(ah), umma_desc(bh), idesc, accumulate);
              if (SPLIT3) {
                umma_tf32(acc, umma_desc(ah + 64u), umma_desc(bh), idesc, 1u);
                umma_tf32(acc, umma_desc(ah), umma_desc(bh + 64u), idesc, 1u);
              }
            }
          }
          umma_commit(empty(s));
        }
        umma_commit(accum_full);
      }
    }
  } else {
    if (lane == 0) {
      for (int it = 0; it < total_stages; ++it) {
        const int s = it % TB_STAGES;
        ok = mbar_wait(empty(s), (((uint32_t)(it / TB_STAGES)) & 1u) ^ 1u) && ok;
        mbar_arrive_expect_tx(full_b(s), b_bytes);
        bulk_g2s(smem_base + (uint32_t)s * stage_bytes + a_bytes, P.wpack + (size_t)it * (b_bytes / 4), b_bytes,
                 full_b(s));
      }
    }
  }
  if (!ok && P.status) atomicOr(P.status, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_free512(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ kernel B
// CTA (ktile, slab): accumulators [256 k'' rows x NPAD] over the slab's rows; needs D % 4 == 0.
// Warp roles: 0-7 build the A tile (thread t owns row k'' = k_lo + t: products x0[b,p,d]*xk[b,q,d] from a staged
// slice), 8-15 build the B tile (thread t owns row n = t - 256 of dF), 16 issues the MMAs.  The two producer groups
// run concurrently, and neither has a division or a 64-bit multiply in its stage loop: (sample, d) positions advance
// incrementally.  On-the-fly operands are split by truncation (the tensor core drops the low 13 mantissa bits of a
// tf32 operand anyway): hi = raw bits, lo = v - trunc(v), exact in fp32.
template <bool SPLIT3>
__global__ void __launch_bounds__(TB_THREADS_B, 1) cin_bwd_dw_tc_kernel(const TbParams P) {
  constexpr int KS = SPLIT3 ? 16 : 32;  // (b,d) rows per stage
  constexpr int KC = KS / 4;            // float4 chunks per row and stage
  constexpr int XS = KS + 4;            // padded row stride of the staged x0 / xk slices (bank-conflict free LDS.128)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = 256u * 128u;
  const uint32_t b_bytes = (uint32_t)P.NPAD * 128u;
  const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TB_STAGES * stage_bytes;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (TB_STAGES + s); };
  const uint32_t accum_full = bar_base + 8u * (2 * TB_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * TB_STAGES + 1);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  // staged x slices: 2 buffers x (MPAD + NQ) rows x XS floats, after the barriers
  const int ktile = blockIdx.x, slab = blockIdx.y;
  const int k_lo = ktile * 256;
  const int q_lo = k_lo / P.MPAD;
  const int q_hi = min(P.H - 1, (k_lo + 255) / P.MPAD);
  const int NQ = max(0, q_hi - q_lo + 1);
  const int xrows = P.MPAD + NQ;
  float* xs = reinterpret_cast<float*>(gen_base + TB_STAGES * stage_bytes + 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < TB_STAGES; ++s) {
      mbar_init(full(s), 512);
      mbar_init(empty(s), 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 16) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  bool ok = true;
  const int64_t r_begin = (int64_t)slab * P.slab_rows;
  const int64_t r_end = min(P.Mrows, r_begin + P.slab_rows);
  const int n_stages = (int)((r_end - r_begin + KS - 1) / KS);
  const uint32_t Du = (uint32_t)P.D;
  const uint32_t rb = (uint32_t)r_begin, re = (uint32_t)r_end;

  if (warp < 8) {
    // =========================== A producers: row k'' = k_lo + tid -> (q, p) ===========================
    const int kk = k_lo + tid;
    const int q = kk / P.MPAD, p = kk - q * P.MPAD;
    const bool a_ok = q < P.H && p < P.m;
    const int ql = q - q_lo;
    const uint32_t row_off = (uint32_t)tid * 128u;
    const uint32_t rx = (uint32_t)(tid & 7);
    const int n_chunks = xrows * KC;  // float4 chunks of the staged slices per stage (<= 3 per thread)
    // per-thread staging slots: chunk i = tid + u*256 -> (row, c4) is the same every stage; only (sample, d) moves
    const float* sp[3];   // nullptr = zero row
    int64_t sbs[3];       // batch stride of the slot's source
    int64_t sob[3];       // sample * batch stride
    uint32_t sd[3], sr[3];
    int sdst[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int i = tid + u * 256;
      sp[u] = nullptr;
      sbs[u] = 0; sob[u] = 0; sd[u] = 0; sr[u] = re; sdst[u] = 0;
      if (i < n_chunks) {
        const int row = i / KC, c4 = i - row * KC;
        sdst[u] = row * XS + 4 * c4;
        if (row < P.MPAD) {
          if (row < P.m) { sp[u] = P.x0 + (int64_t)row * P.D; sbs[u] = P.bs0; }
        } else {
          sp[u] = P.xk + (int64_t)(q_lo + row - P.MPAD) * P.D; sbs[u] = P.bsk;
        }
        sr[u] = rb + 4u * c4;
        const uint32_t b = sr[u] / Du;
        sd[u] = sr[u] - b * Du;
        sob[u] = (int64_t)b * sbs[u];
      }
    }
    const bool slot_live[3] = {tid < n_chunks, tid + 256 < n_chunks, tid + 512 < n_chunks};
    float4 stg[3];
    auto issue_loads = [&]() {  // loads of the next stage's slices, then advance the slots by KS rows
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        stg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sp[u] && sr[u] < re) stg[u] = ld4(sp[u] + sob[u] + sd[u]);
        sr[u] += KS;
        sd[u] += KS;
        while (sd[u] >= Du) {
          sd[u] -= Du;
          sob[u] += sbs[u];
        }
      }
    };
    if (n_stages > 0) issue_loads();
    const float* xa_src = xs + p * XS;
    const float* xb_src = xs + (P.MPAD + ql) * XS;
    for (int st = 0; st < n_stages; ++st) {
      const int s = st % TB_STAGES;
      const int boff = (st & 1) * xrows * XS;
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (slot_live[u]) *reinterpret_cast<float4*>(xs + boff + sdst[u]) = stg[u];
      named_bar_sync(1, 256);  // staged slices visible to all A producers
      if (st + 1 < n_stages) issue_loads();
      ok = mbar_wait(empty(s), (((uint32_t)(st / TB_STAGES)) & 1u) ^ 1u) && ok;
      uint8_t* arow = gen_base + (size_t)s * stage_bytes + row_off;
#pragma unroll
      for (int c4 = 0; c4 < KC; ++c4) {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_ok) {
          const float4 xa = *reinterpret_cast<const float4*>(xa_src + boff + 4 * c4);
          const float4 xb = *reinterpret_cast<const float4*>(xb_src + boff + 4 * c4);
          z = make_float4(xa.x * xb.x, xa.y * xb.y, xa.z * xb.z, xa.w * xb.w);
        }
        uint4 hi, lo;
        split_trunc4(z, hi, lo);
        *reinterpret_cast<uint4*>(arow + ((((uint32_t)c4) ^ rx) << 4)) = hi;
        if (SPLIT3) *reinterpret_cast<uint4*>(arow + ((((uint32_t)(c4 + 4)) ^ rx) << 4)) = lo;
      }
      fence_proxy_async_smem();
      mbar_arrive(full(s));
    }
    // ---- epilogue: partial[slab][k_lo + row][n] ----
    ok = mbar_wait(accum_full, 0) && ok;
    tc_fence_after();
    const int h = warp >> 2, quad = warp & 3;
    const int rowk = k_lo + h * 128 + quad * 32 + lane;
    float* dst = P.partial + ((int64_t)slab * P.KPADT + rowk) * P.NPAD;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * 256);
    for (int col0 = 0; col0 < P.NPAD; col0 += 16) {
      uint32_t a[16];
      tmem_ld16(taddr0 + (uint32_t)col0, a);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<uint4*>(dst + col0 + j) = make_uint4(a[j], a[j + 1], a[j + 2], a[j + 3]);
    }
  } else if (warp < 16) {
    // =========================== B producers: row n = tid - 256 of dF ===========================
    const int n = tid - 256;
    const bool b_ok = n < P.N;
    const bool b_row = n < P.NPAD;
    const uint32_t row_off = (uint32_t)n * 128u;
    const uint32_t rx = (uint32_t)(n & 7);
    const int64_t ND = (int64_t)P.N * P.D;
    uint32_t r = rb;
    uint32_t d = rb % Du;
    const float* dfp = P.dF + ((int64_t)(rb / Du) * P.N + (b_ok ? n : 0)) * P.D;  // (sample, n, 0)
    float4 df[2][KC];
    auto issue_loads = [&](float4 (&dst)[KC]) {
#pragma unroll
      for (int c4 = 0; c4 < KC; ++c4) {
        dst[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b_ok && r < re) dst[c4] = ld4(dfp + d);
        r += 4;
        d += 4;
        if (d >= Du) {  // D % 4 == 0: a chunk never straddles two samples
          d -= Du;
          dfp += ND;
        }
      }
    };
    auto stage = [&](int st, float4 (&cur)[KC], float4 (&nxt)[KC]) {
      const int s = st % TB_STAGES;
      if (st + 1 < n_stages) issue_loads(nxt);
      ok = mbar_wait(empty(s), (((uint32_t)(st / TB_STAGES)) & 1u) ^ 1u) && ok;
      if (b_row) {
        uint8_t* brow = gen_base + (size_t)s * stage_bytes + a_bytes + row_off;
#pragma unroll
        for (int c4 = 0; c4 < KC; ++c4) {
          uint4 hi, lo;
          split_trunc4(cur[c4], hi, lo);
          *reinterpret_cast<uint4*>(brow + ((((uint32_t)c4) ^ rx) << 4)) = hi;
          if (SPLIT3) *reinterpret_cast<uint4*>(brow + ((((uint32_t)(c4 + 4)) ^ rx) << 4)) = lo;
        }
        fence_proxy_async_smem();
      }
      mbar_arrive(full(s));
    };
    if (n_stages > 0) issue_loads(df[0]);
    int st = 0;
#pragma unroll 1
    for (; st + 1 < n_stages; st += 2) {  // unrolled by two: the prefetch registers alternate without moves
      stage(st, df[0], df[1]);
      stage(st + 1, df[1], df[0]);
    }
    if (st < n_stages) stage(st, df[0], df[1]);
  } else {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, P.NPAD);
      for (int st = 0; st < n_stages; ++st) {
        const int s = st % TB_STAGES;
        ok = mbar_wait(full(s), ((uint32_t)(st / TB_STAGES)) & 1u) && ok;
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
        for (int ks = 0; ks < (SPLIT3 ? 2 : 4); ++ks) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t acc = tmem_base + (uint32_t)(h * 256);
            const uint32_t accumulate = (st > 0 || ks > 0) ? 1u : 0u;
            const uint32_t ah = a_addr + (uint32_t)h * (128u * 128u) + (uint32_t)ks * 32u;
            const uint32_t bh = b_addr + (uint32_t)ks * 32u;
            umma_tf32(acc, umma_desc(ah), umma_desc(bh), idesc, accumulate);
            if (SPLIT3) {
              umma_tf32(acc, umma_desc(ah + 64u), umma_desc(bh), idesc, 1u);
              umma_tf32(acc, umma_desc(ah), umma_desc(bh + 64u), idesc, 1u);
            }
          }
        }
        umma_commit(empty(s));
      }
      umma_commit(accum_full);
    }
  }
  if (!ok && P.status) atomicOr(P.status, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_free512(tmem_base);
  }
}

// dW[(p*H + q), n] = sum over slabs (in order, fp32 round-to-nearest) of partial[slab][q*MPAD + p][n]
__global__ void __launch_bounds__(256) cin_dw_unpermute_kernel(const float* __restrict__ partial, int slabs, int KPADT,
                                                               int NPAD, int m, int H, int N, int MPAD,
                                                               float* __restrict__ dW) {
  const int64_t total = (int64_t)m * H * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int64_t pq = i / N;
    const int q = (int)(pq % H), p = (int)(pq / H);
    const int64_t off = ((int64_t)q * MPAD + p) * NPAD + n;
    float acc = 0.f;
    for (int s = 0; s < slabs; ++s) acc += partial[(int64_t)s * KPADT * NPAD + off];
    dW[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct TbLayout {
  int MPAD, MP4, QG, NMMA, spq, n_qg, NPAD, KPADT, n_ktiles, slabs, slab_rows;
  size_t off_status, off_dF, off_wpack, off_partial, total;
};

static TbLayout tb_layout(int64_t B, int m, int H, int D, int N, int precision) {
  TbLayout L;
  const bool split3 = precision == RM_CIN_3XTF32;
  const int KS = split3 ? 16 : 32;
  L.MP4 = (m + 3) / 4;
  L.MPAD = L.MP4 * 4;
  int qg = (256 / L.MPAD) / 4 * 4;  // multiple of 4 -> NMMA % 16 == 0
  if (qg < 4) qg = 4;
  if (qg * L.MPAD > 256) qg = 4;  // MPAD <= 32 -> 4*MPAD <= 128
  const int h4 = (H + 3) / 4 * 4;
  if (qg > h4) qg = h4;  // do not pad tiny H up to a full group
  L.QG = qg;
  L.NMMA = L.QG * L.MPAD;
  L.n_qg = (H + L.QG - 1) / L.QG;
  L.spq = (N + KS - 1) / KS;
  L.NPAD = (N + 15) / 16 * 16;
  const int kpp = H * L.MPAD;
  L.n_ktiles = (kpp + 255) / 256;
  L.KPADT = L.n_ktiles * 256;
  const int64_t Mrows = B * (int64_t)D;
  // slab size: 320..512 rows (the TMEM accumulation truncates: <= 512 accumulated rows keep the bias under 1e-5 of
  // max|dW| in parity mode), chosen so that the grid
  // n_ktiles x slabs fills whole waves of 148 CTAs (1 CTA / SM) and the per-CTA prologue/epilogue is amortised
  L.slab_rows = 0;
  {
    double best = -1.0;
    for (int cand = 320; cand <= 512; cand += 32) {
      const int64_t slabs = (Mrows + cand - 1) / cand;
      const int64_t total = slabs * L.n_ktiles;
      const int64_t waves = (total + RM_NUM_SMS - 1) / RM_NUM_SMS;
      const double score = (double)total / (double)(waves * RM_NUM_SMS) * cand / (cand + 96.0);
      if (score > best) {
        best = score;
        L.slab_rows = cand;
      }
    }
  }
  if (L.slab_rows < 32) L.slab_rows = 32;
  L.slabs = (int)((Mrows + L.slab_rows - 1) / L.slab_rows);
  if (L.slabs < 1) L.slabs = 1;
  size_t off = 0;
  L.off_status = off; off += 256;
  L.off_dF = off; off += align_up((size_t)B * N * D * 4, 256);
  L.off_wpack = off; off += align_up((size_t)L.n_qg * L.spq * L.NMMA * 128, 256);
  L.off_partial = off; off += align_up((size_t)L.slabs * L.KPADT * L.NPAD * 4, 256);
  L.total = off;
  return L;
}

bool cin_tc_bwd_supported(int64_t B, int m, int H, int D, int N) {
  (void)B;
  (void)H;
  return m >= 1 && m <= 32 && N >= 1 && N <= 256 && D % 4 == 0;
}

size_t cin_tc_bwd_workspace(int64_t B, int m, int H, int D, int N, int precision) {
  return tb_layout(B, m, H, D, N, precision).total;
}

template <int MP4>
static int launch_dx(const TbParams& P, bool split3, int grid, size_t smem, cudaStream_t st) {
  if (split3) {
    RM_SMEM_ATTR_ONCE(smem, cin_bwd_dx_tc_kernel<MP4, true>);
    cin_bwd_dx_tc_kernel<MP4, true><<<grid, TB_THREADS, smem, st>>>(P);
  } else {
    RM_SMEM_ATTR_ONCE(smem, cin_bwd_dx_tc_kernel<MP4, false>);
    cin_bwd_dx_tc_kernel<MP4, false><<<grid, TB_THREADS, smem, st>>>(P);
  }
  RM_LAUNCH_CHECK();
  return 0;
}

int cin_bwd_tc(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* pre,
               const float* dout, int64_t B, int m, int H, int D, int N, int act, int precision, float* dW,
               float* dbias, float* dx0, float* dxk, int64_t dbsk, void* workspace, size_t workspace_bytes,
               cudaStream_t st) {
  const bool split3 = precision == RM_CIN_3XTF32;
  const TbLayout L = tb_layout(B, m, H, D, N, precision);
  if (workspace_bytes < L.total) {
    set_error("rm_cin_layer_bwd: workspace %zu < required %zu", workspace_bytes, L.total);
    return RM_E_WORKSPACE;
  }
  RM_UNSUPPORTED(bs0 % 4 == 0 && bsk % 4 == 0 && aligned16(x0) && aligned16(xk), "tensor-core backward needs 16-byte aligned rows");
  RM_UNSUPPORTED(B * (int64_t)D < ((int64_t)1 << 31), "tensor-core backward needs B*D < 2^31 (32-bit row indices)");
  char* ws = (char*)workspace;
  int32_t* status = (int32_t*)(ws + L.off_status);
  float* dF = (float*)(ws + L.off_dF);
  uint32_t* wpack = (uint32_t*)(ws + L.off_wpack);
  float* partial = (float*)(ws + L.off_partial);
  RM_CUDA(cudaMemsetAsync(status, 0, 256, st));
  {
    const int rc = cin_dF_dbias(dout, pre, B, N, D, act, dF, dbias, partial, (size_t)L.slabs * L.KPADT * L.NPAD, st);
    if (rc) return rc;
  }
  cin_pack_wT_kernel<<<grid_for((int64_t)L.n_qg * L.spq * L.NMMA * 8, 256, 8), 256, 0, st>>>(
      W, m, H, N, L.MPAD, L.QG, L.NMMA, L.spq, L.n_qg, split3 ? 1 : 0, wpack);
  RM_LAUNCH_CHECK();
  TbParams P;
  P.x0 = x0; P.bs0 = bs0; P.xk = xk; P.bsk = bsk; P.dF = dF; P.wpack = wpack; P.dx0 = dx0; P.dxk = dxk; P.dbsk = dbsk;
  P.partial = partial; P.status = status; P.Mrows = B * (int64_t)D; P.m = m; P.H = H; P.D = D; P.N = N;
  P.NPAD = L.NPAD; P.MPAD = L.MPAD; P.QG = L.QG; P.NMMA = L.NMMA; P.spq = L.spq; P.n_qg = L.n_qg; P.KPADT = L.KPADT;
  P.slab_rows = L.slab_rows;
  // ---- kernel A ----
  {
    const uint32_t stage_bytes = (uint32_t)((256 * 128 + L.NMMA * 128 + 1023) / 1024 * 1024);
    const size_t smem = (size_t)TB_STAGES * stage_bytes + 8 * (3 * TB_STAGES + 3) + 1024;
    const int grid = (int)ceil_div(P.Mrows, 256);
    int rc;
    switch (L.MP4) {
      case 1: rc = launch_dx<1>(P, split3, grid, smem, st); break;
      case 2: rc = launch_dx<2>(P, split3, grid, smem, st); break;
      case 3: rc = launch_dx<3>(P, split3, grid, smem, st); break;
      case 4: rc = launch_dx<4>(P, split3, grid, smem, st); break;
      case 5: rc = launch_dx<5>(P, split3, grid, smem, st); break;
      case 6: rc = launch_dx<6>(P, split3, grid, smem, st); break;
      case 7: rc = launch_dx<7>(P, split3, grid, smem, st); break;
      default: rc = launch_dx<8>(P, split3, grid, smem, st); break;
    }
    if (rc) return rc;
  }
  // ---- kernel B ----
  {
    const int KS = split3 ? 16 : 32;
    const uint32_t stage_bytes = (uint32_t)((256 * 128 + L.NPAD * 128 + 1023) / 1024 * 1024);
    const int max_xrows = L.MPAD + 256 / L.MPAD + 2;
    const size_t smem = (size_t)TB_STAGES * stage_bytes + 128 + (size_t)2 * max_xrows * (KS + 4) * 4 + 1024;
    RM_UNSUPPORTED(smem <= 227 * 1024, "shared memory budget exceeded in the dW kernel");
    RM_UNSUPPORTED(max_xrows * (KS / 4) <= 3 * 256, "staged slice too large for the dW kernel's prefetch registers");
    dim3 grid((unsigned)L.n_ktiles, (unsigned)L.slabs);
    if (split3) {
      RM_SMEM_ATTR_ONCE(smem, cin_bwd_dw_tc_kernel<true>);
      cin_bwd_dw_tc_kernel<true><<<grid, TB_THREADS_B, smem, st>>>(P);
    } else {
      RM_SMEM_ATTR_ONCE(smem, cin_bwd_dw_tc_kernel<false>);
      cin_bwd_dw_tc_kernel<false><<<grid, TB_THREADS_B, smem, st>>>(P);
    }
    RM_LAUNCH_CHECK();
    cin_dw_unpermute_kernel<<<grid_for((int64_t)m * H * N, 256, 8), 256, 0, st>>>(partial, L.slabs, L.KPADT, L.NPAD, m, H,
                                                                                  N, L.MPAD, dW);
    RM_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace rm
