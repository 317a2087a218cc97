// K3: FM layer forward / backward on an already gathered [B, m, k] block.
// Replaces FMLayer.__call__ (recman/tf/core/layers.py:457-478): six TF
// elementwise/reduce ops become one pass.  Roofline: HBM; algorithmic bytes per
// sample fwd 4*m*k + 4*m + 4, bwd 4mk read + 4 + 4mk write.
#include "common.cuh"

namespace rm {

// vector path: LPR lanes (float4 each) per sample
template <int LPR, int U>
__global__ void __launch_bounds__(256) fm_fwd_vec_kernel(const float* __restrict__ e, int64_t ld,
                                                         const float* __restrict__ bias, int64_t B, int m, int k,
                                                         float* __restrict__ out, float* __restrict__ sum_out) {
  const int lir = threadIdx.x % LPR;
  const int k4 = k >> 2;
  const bool col_ok = lir < k4;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  const int64_t iters = (B + n_groups - 1) / n_groups;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t b = group + it * n_groups;
    const bool live = b < B;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = S;
    float bias_acc = 0.f;
    if (live) {
      const float* row = e + b * ld + 4 * lir;
      for (int f0 = 0; f0 < m; f0 += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (f0 + u < m && col_ok) v[u] = ld4(row + (int64_t)(f0 + u) * k);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          S.x += v[u].x; S.y += v[u].y; S.z += v[u].z; S.w += v[u].w;
          Q.x += v[u].x * v[u].x; Q.y += v[u].y * v[u].y; Q.z += v[u].z * v[u].z; Q.w += v[u].w * v[u].w;
        }
      }
      if (bias)
        for (int f = lir; f < m; f += LPR) bias_acc += bias[b * m + f];
      if (sum_out && col_ok) st4(sum_out + b * k + 4 * lir, S);
    }
    float second = 0.5f * (S.x * S.x - Q.x) + 0.5f * (S.y * S.y - Q.y) + 0.5f * (S.z * S.z - Q.z) +
                   0.5f * (S.w * S.w - Q.w);
    second = group_sum<LPR>(second);
    bias_acc = group_sum<LPR>(bias_acc);
    if (live && lir == 0) out[b] = bias_acc + second;
  }
}

// scalar path: one warp per sample, any k
__global__ void __launch_bounds__(256) fm_fwd_scalar_kernel(const float* __restrict__ e, int64_t ld,
                                                            const float* __restrict__ bias, int64_t B, int m, int k,
                                                            float* __restrict__ out, float* __restrict__ sum_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += n_warps) {
    float acc = 0.f;
    for (int c = lane; c < k; c += 32) {
      float s = 0.f, q = 0.f;
      for (int f = 0; f < m; ++f) {
        const float v = e[b * ld + (int64_t)f * k + c];
        s += v;
        q += v * v;
      }
      if (sum_out) sum_out[b * k + c] = s;
      acc += 0.5f * (s * s - q);
    }
    if (bias)
      for (int f = lane; f < m; f += 32) acc += bias[b * m + f];
    acc = warp_sum(acc);
    if (lane == 0) out[b] = acc;
  }
}

// backward, vector path.  S given or recomputed.
template <int LPR, int U>
__global__ void __launch_bounds__(256) fm_bwd_vec_kernel(const float* __restrict__ e, int64_t ld,
                                                         const float* __restrict__ sum, const float* __restrict__ gout,
                                                         int64_t B, int m, int k, float* __restrict__ de, int64_t d_ld,
                                                         float* __restrict__ d_bias, int accumulate) {
  const int lir = threadIdx.x % LPR;
  const int k4 = k >> 2;
  const bool col_ok = lir < k4;
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  for (int64_t b = group; b < B; b += n_groups) {
    const float g = gout[b];
    if (d_bias)
      for (int f = lir; f < m; f += LPR) d_bias[b * m + f] = g;
    if (!col_ok) continue;
    const float* row = e + b * ld + 4 * lir;
    float* drow = de + b * d_ld + 4 * lir;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sum) {
      S = ld4(sum + b * k + 4 * lir);
    } else {
      for (int f = 0; f < m; ++f) {
        const float4 v = ld4(row + (int64_t)f * k);
        S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
      }
    }
    for (int f0 = 0; f0 < m; f0 += U) {
      float4 v[U], a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        a[u] = v[u];
        if (f0 + u < m) {
          v[u] = ld4(row + (int64_t)(f0 + u) * k);
          if (accumulate) a[u] = ld4(drow + (int64_t)(f0 + u) * k);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (f0 + u < m) {
          float4 r;
          r.x = a[u].x + g * (S.x - v[u].x);
          r.y = a[u].y + g * (S.y - v[u].y);
          r.z = a[u].z + g * (S.z - v[u].z);
          r.w = a[u].w + g * (S.w - v[u].w);
          st4(drow + (int64_t)(f0 + u) * k, r);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) fm_bwd_scalar_kernel(const float* __restrict__ e, int64_t ld,
                                                            const float* __restrict__ sum,
                                                            const float* __restrict__ gout, int64_t B, int m, int k,
                                                            float* __restrict__ de, int64_t d_ld,
                                                            float* __restrict__ d_bias, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += n_warps) {
    const float g = gout[b];
    if (d_bias)
      for (int f = lane; f < m; f += 32) d_bias[b * m + f] = g;
    for (int c = lane; c < k; c += 32) {
      float s = 0.f;
      if (sum) {
        s = sum[b * k + c];
      } else {
        for (int f = 0; f < m; ++f) s += e[b * ld + (int64_t)f * k + c];
      }
      for (int f = 0; f < m; ++f) {
        const int64_t o = (int64_t)f * k + c;
        const float r = g * (s - e[b * ld + o]);
        de[b * d_ld + o] = accumulate ? de[b * d_ld + o] + r : r;
      }
    }
  }
}

static inline int pow2ceil_(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int LPR>
static int launch_fm_fwd(const float* e, int64_t ld, const float* bias, int64_t B, int m, int k, float* out,
                         float* sum_out, cudaStream_t st) {
  const int grid = grid_for(B, 256 / LPR, 8);
  fm_fwd_vec_kernel<LPR, 4><<<grid, 256, 0, st>>>(e, ld, bias, B, m, k, out, sum_out);
  RM_LAUNCH_CHECK();
  return 0;
}
template <int LPR>
static int launch_fm_bwd(const float* e, int64_t ld, const float* sum, const float* gout, int64_t B, int m, int k,
                         float* de, int64_t d_ld, float* d_bias, int accumulate, cudaStream_t st) {
  const int grid = grid_for(B, 256 / LPR, 8);
  fm_bwd_vec_kernel<LPR, 4><<<grid, 256, 0, st>>>(e, ld, sum, gout, B, m, k, de, d_ld, d_bias, accumulate);
  RM_LAUNCH_CHECK();
  return 0;
}

}  // namespace rm

extern "C" {

int rm_fm_fwd(const float* embeds, int64_t ld, const float* bias, int64_t B, int32_t m, int32_t k, float* out,
              float* sum_out, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(embeds && out, "null pointer");
  RM_CHECK_ARG(B >= 0 && m > 0 && k > 0 && ld >= (int64_t)m * k, "bad shape");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (k % 4 == 0) && k <= 128 && (ld % 4 == 0) && aligned16(embeds) && (!sum_out || aligned16(sum_out));
  if (vec) {
    switch (pow2ceil_(k / 4)) {
      case 1: return launch_fm_fwd<1>(embeds, ld, bias, B, m, k, out, sum_out, st);
      case 2: return launch_fm_fwd<2>(embeds, ld, bias, B, m, k, out, sum_out, st);
      case 4: return launch_fm_fwd<4>(embeds, ld, bias, B, m, k, out, sum_out, st);
      case 8: return launch_fm_fwd<8>(embeds, ld, bias, B, m, k, out, sum_out, st);
      case 16: return launch_fm_fwd<16>(embeds, ld, bias, B, m, k, out, sum_out, st);
      default: return launch_fm_fwd<32>(embeds, ld, bias, B, m, k, out, sum_out, st);
    }
  }
  fm_fwd_scalar_kernel<<<grid_for(B, 8, 8), 256, 0, st>>>(embeds, ld, bias, B, m, k, out, sum_out);
  RM_LAUNCH_CHECK();
  return 0;
}

int rm_fm_bwd(const float* embeds, int64_t ld, const float* sum, const float* gout, int64_t B, int32_t m, int32_t k,
              float* d_embeds, int64_t d_ld, float* d_bias, int32_t accumulate, void* stream) {
  using namespace rm;
  RM_CHECK_ARG(embeds && gout && d_embeds, "null pointer");
  RM_CHECK_ARG(B >= 0 && m > 0 && k > 0 && ld >= (int64_t)m * k && d_ld >= (int64_t)m * k, "bad shape");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (k % 4 == 0) && k <= 128 && (ld % 4 == 0) && (d_ld % 4 == 0) && aligned16(embeds) &&
                   aligned16(d_embeds) && (!sum || aligned16(sum));
  if (vec) {
    switch (pow2ceil_(k / 4)) {
      case 1: return launch_fm_bwd<1>(embeds, ld, sum, gout, B, m, k, d_embeds, d_ld, d_bias, accumulate, st);
      case 2: return launch_fm_bwd<2>(embeds, ld, sum, gout, B, m, k, d_embeds, d_ld, d_bias, accumulate, st);
      case 4: return launch_fm_bwd<4>(embeds, ld, sum, gout, B, m, k, d_embeds, d_ld, d_bias, accumulate, st);
      case 8: return launch_fm_bwd<8>(embeds, ld, sum, gout, B, m, k, d_embeds, d_ld, d_bias, accumulate, st);
      case 16: return launch_fm_bwd<16>(embeds, ld, sum, gout, B, m, k, d_embeds, d_ld, d_bias, accumulate, st);
      default: return launch_fm_bwd<32>(embeds, ld, sum, gout, B, m, k, d_embeds, d_ld, d_bias, accumulate, st);
    }
  }
  fm_bwd_scalar_kernel<<<grid_for(B, 8, 8), 256, 0, st>>>(embeds, ld, sum, gout, B, m, k, d_embeds, d_ld, d_bias,
                                                          accumulate);
  RM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
