// K5 (CUDA-core path): CIN layer forward / backward in plain fp32 FFMA.
//
// recman/tf/core/layers.py:711-751:  Z[(b,d),(p,q)] = x0[b,p,d]*xk[b,q,d];  F = act(Z.W + bias);
// out[b,n,d] = F[(b,d),n].  Z (3 GB at config C3) is never materialised: tiles of it are synthesised
// in shared memory from the x0 / xk rows of the CTA's samples.
//
// This path is (a) the fp32 ground truth the tcgen05 kernels are verified against on the device and
// (b) the backward used until the tensor-core backward lands.  Everything is deterministic: no atomics,
// batch reductions go through fixed-size slabs summed in slab order.
#include "cin.cuh"

namespace rm {

constexpr int CIN_BM = 128;  // rows (b,d) per CTA tile
constexpr int CIN_BN = 64;   // output columns per CTA tile
constexpr int CIN_BK = 16;   // reduction step
constexpr int CIN_AS = CIN_BM + 16;  // padded row stride of the A tile (conflict-free stores)

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == RM_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == RM_ACT_LEAKY_RELU) return fmaxf(0.2f * v, v);  // tf.nn.leaky_relu, alpha = 0.2
  return v;
}
__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == RM_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  if (act == RM_ACT_LEAKY_RELU) return pre > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// ---------------------------------------------------------------------------------------------
// forward:  grid (ceil(B/TB), ceil(N/64)), 256 threads; thread (ty,tx) owns rows ty*8.., cols tx*4..
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cin_fwd_simt_kernel(const float* __restrict__ x0, int64_t bs0,
                                                           const float* __restrict__ xk, int64_t bsk,
                                                           const float* __restrict__ W, const float* __restrict__ bias,
                                                           int B, int m, int H, int D, int N, int act, int TB,
                                                           float* __restrict__ out, float* __restrict__ pre) {
  extern __shared__ float smem[];
  float* x0s = smem;                       // [TB][m][D]
  float* xks = x0s + TB * m * D;           // [TB][H][D]
  float* As = xks + TB * H * D;            // [BK][CIN_AS]
  float* Bs = As + CIN_BK * CIN_AS;        // [BK][BN]
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * TB;
  const int n0 = blockIdx.y * CIN_BN;
  const int nb = min(TB, B - b0);
  const int rows_used = nb * D;
  const int K = m * H;
  for (int i = tid; i < TB * m * D; i += 256) {
    const int bl = i / (m * D), rem = i - bl * (m * D);
    x0s[i] = bl < nb ? x0[(int64_t)(b0 + bl) * bs0 + rem] : 0.f;
  }
  for (int i = tid; i < TB * H * D; i += 256) {
    const int bl = i / (H * D), rem = i - bl * (H * D);
    xks[i] = bl < nb ? xk[(int64_t)(b0 + bl) * bsk + rem] : 0.f;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += CIN_BK) {
    {  // synthesise the A tile: As[kk][r] = x0[b,p,d]*xk[b,q,d]
      const int kk = tid >> 4;
      const int kidx = k0 + kk;
      const bool kok = kidx < K;
      const int p = kok ? kidx / H : 0, q = kok ? kidx - p * H : 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (tid & 15) + 16 * i;
        float a = 0.f;
        if (kok && r < rows_used) {
          const int bl = r / D, d = r - bl * D;
          a = x0s[(bl * m + p) * D + d] * xks[(bl * H + q) * D + d];
        }
        As[kk * CIN_AS + r] = a;
      }
      // B tile: Bs[kk][n] = W[kidx][n0+n]
      const int nq = (tid & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + nq + j;
        Bs[kk * CIN_BN + nq + j] = (kok && n < N) ? __ldg(W + (int64_t)kidx * N + n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CIN_BK; ++kk) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk * CIN_AS + ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk * CIN_BN + tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    if (n >= N) continue;
    const float bv = __ldg(bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty * 8 + i;
      if (r < rows_used) {
        const int bl = r / D, d = r - bl * D;
        const int64_t o = ((int64_t)(b0 + bl) * N + n) * D + d;
        const float v = acc[i][j] + bv;
        if (pre) pre[o] = v;
        out[o] = act_fwd(v, act);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward 0: dF = dout * act'(pre)  (elementwise) and dbias[n] = sum_{b,d} dF[b,n,d]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cin_dF_kernel(const float* __restrict__ dout, const float* __restrict__ pre,
                                                     int64_t total, int act, float* __restrict__ dF) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    dF[i] = dout[i] * act_grad(pre[i], act);
}

// one CTA per output column n; fixed assignment of (b,d) pairs to threads + fixed tree -> deterministic
__global__ void __launch_bounds__(256) cin_dbias_kernel(const float* __restrict__ dF, int B, int N, int D,
                                                        float* __restrict__ dbias) {
  __shared__ float red[256];
  const int n = blockIdx.x;
  float acc = 0.f;
  const int64_t total = (int64_t)B * D;
  for (int64_t i = threadIdx.x; i < total; i += 256) {
    const int64_t b = i / D;
    const int d = (int)(i - b * D);
    acc += dF[(b * N + n) * D + d];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dbias[n] = red[0];
}

// ---------------------------------------------------------------------------------------------
// backward A: dZ = dF . W^T tile by tile, contracted on the fly into dx0 (+=) and dxk (=)
//   dx0[b,p,d] += sum_q dZ[(b,d),(p,q)] * xk[b,q,d]      dxk[b,q,d] = sum_p dZ[(b,d),(p,q)] * x0[b,p,d]
// One CTA owns TB samples and walks K tiles (p fixed, 64 q's): every (row, q) element has exactly one
// owner thread for the whole kernel, so the shared-memory accumulators need no atomics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cin_bwd_dx_simt_kernel(const float* __restrict__ x0, int64_t bs0,
                                                              const float* __restrict__ xk, int64_t bsk,
                                                              const float* __restrict__ W, const float* __restrict__ dF,
                                                              int B, int m, int H, int D, int N, int TB,
                                                              float* __restrict__ dx0, float* __restrict__ dxk,
                                                              int64_t dbsk) {
  extern __shared__ float smem[];
  float* x0s = smem;                    // [TB][m][D]
  float* xks = x0s + TB * m * D;        // [TB][H][D]
  float* dx0s = xks + TB * H * D;       // [TB][m][D]
  float* dxks = dx0s + TB * m * D;      // [TB][H][D]
  float* As = dxks + TB * H * D;        // [BK][CIN_AS]   dF tile  (n-chunk x rows)
  float* Bs = As + CIN_BK * CIN_AS;     // [BK][BN]       W^T tile (n-chunk x q-chunk)
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * TB;
  const int nb = min(TB, B - b0);
  const int rows_used = nb * D;
  for (int i = tid; i < TB * m * D; i += 256) {
    const int bl = i / (m * D), rem = i - bl * (m * D);
    x0s[i] = bl < nb ? x0[(int64_t)(b0 + bl) * bs0 + rem] : 0.f;
    dx0s[i] = 0.f;
  }
  for (int i = tid; i < TB * H * D; i += 256) {
    const int bl = i / (H * D), rem = i - bl * (H * D);
    xks[i] = bl < nb ? xk[(int64_t)(b0 + bl) * bsk + rem] : 0.f;
    dxks[i] = 0.f;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  for (int p = 0; p < m; ++p) {
    for (int q0 = 0; q0 < H; q0 += CIN_BN) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int n0 = 0; n0 < N; n0 += CIN_BK) {
        {
          const int nn = tid >> 4;
          const int n = n0 + nn;
          const bool nok = n < N;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = (tid & 15) + 16 * i;
            float a = 0.f;
            if (nok && r < rows_used) {
              const int bl = r / D, d = r - bl * D;
              a = dF[((int64_t)(b0 + bl) * N + n) * D + d];
            }
            As[nn * CIN_AS + r] = a;
          }
          const int qq = (tid & 15) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int q = q0 + qq + j;
            Bs[nn * CIN_BN + qq + j] = (nok && q < H) ? __ldg(W + (int64_t)(p * H + q) * N + n) : 0.f;
          }
        }
        __syncthreads();
#pragma unroll
        for (int nn = 0; nn < CIN_BK; ++nn) {
          float a[8], b[4];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = As[nn * CIN_AS + ty * 8 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[nn * CIN_BN + tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
      // contraction of the dZ micro-tile acc[i][j] = dZ[row ty*8+i][(p, q0 + tx*4 + j)]
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty * 8 + i;
        float s0 = 0.f;  // partial of dx0[b,p,d] over this thread's 4 q's
        if (r < rows_used) {
          const int bl = r / D, d = r - bl * D;
          const float x0v = x0s[(bl * m + p) * D + d];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int q = q0 + tx * 4 + j;
            if (q < H) {
              const int o = (bl * H + q) * D + d;
              dxks[o] += acc[i][j] * x0v;  // (row, q) has a single owner thread: plain RMW
              s0 += acc[i][j] * xks[o];
            }
          }
        }
        // reduce over the 16 tx lanes that share this row (lanes of one half-warp), fixed xor tree
        s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 4);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        if (tx == 0 && r < rows_used) {
          const int bl = r / D, d = r - bl * D;
          dx0s[(bl * m + p) * D + d] += s0;
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nb * m * D; i += 256) {
    const int bl = i / (m * D), rem = i - bl * (m * D);
    dx0[(int64_t)(b0 + bl) * m * D + rem] += dx0s[i];  // this CTA owns these samples
  }
  for (int i = tid; i < nb * H * D; i += 256) {
    const int bl = i / (H * D), rem = i - bl * (H * D);
    dxk[(int64_t)(b0 + bl) * dbsk + rem] = dxks[i];
  }
}

// ---------------------------------------------------------------------------------------------
// backward B: dW[(p,q), n] = sum_{(b,d)} Z[(b,d),(p,q)] * dF[(b,d), n]
// grid (K tiles (p, 64 q's), ceil(N/64), slabs); each CTA reduces its slab of samples into a partial
// tile; cin_dw_final_kernel sums the slabs in order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cin_bwd_dw_simt_kernel(const float* __restrict__ x0, int64_t bs0,
                                                              const float* __restrict__ xk, int64_t bsk,
                                                              const float* __restrict__ dF, int B, int m, int H, int D,
                                                              int N, int q_tiles, int slab_samples,
                                                              float* __restrict__ partial) {
  __shared__ float As[CIN_BK][CIN_BN + 4];  // [r][kk]  Z rows
  __shared__ float Bs[CIN_BK][CIN_BN + 4];  // [r][n]   dF rows
  const int tid = threadIdx.x;
  const int p = blockIdx.x / q_tiles;
  const int q0 = (blockIdx.x - p * q_tiles) * CIN_BN;
  const int n0 = blockIdx.y * CIN_BN;
  const int slab = blockIdx.z;
  const int bs = slab * slab_samples;
  const int be = min(B, bs + slab_samples);
  const int64_t r_begin = (int64_t)bs * D, r_end = (int64_t)be * D;
  const int ty = tid >> 4, tx = tid & 15;  // thread owns kk = ty*4.., n = tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += CIN_BK) {
    {
      // 16 rows x 64 columns per tile = 1024 elements, 4 per thread; consecutive threads take consecutive rows
      // (consecutive d of one sample) so that global reads coalesce along d.
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e;
        const int rr = idx & 15, c = idx >> 4;
        const int64_t r = r0 + rr;
        float a = 0.f, bv = 0.f;
        if (r < r_end) {
          const int64_t b = r / D;
          const int d = (int)(r - b * D);
          const int q = q0 + c;
          if (q < H) a = x0[b * bs0 + (int64_t)p * D + d] * xk[b * bsk + (int64_t)q * D + d];
          const int n = n0 + c;
          if (n < N) bv = dF[(b * N + n) * D + d];
        }
        As[rr][c] = a;
        Bs[rr][c] = bv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < CIN_BK; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[rr][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[rr][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int64_t KN = (int64_t)m * H * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= H) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) partial[(int64_t)slab * KN + (int64_t)(p * H + q) * N + n] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256) cin_dw_final_kernel(const float* __restrict__ partial, int64_t KN, int slabs,
                                                           float* __restrict__ dW) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < KN; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < slabs; ++s) acc += partial[(int64_t)s * KN + i];
    dW[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int simt_tb(int D) { return D <= CIN_BM ? CIN_BM / D : 0; }

// dF = dout * act'(pre) and dbias in one pass over [B, N*D]: CTA g owns a contiguous range of samples, thread t owns
// the positions idx = t (mod 256) of the N*D plane and accumulates them in shared memory (no conflicts, fixed order),
// then sums each n over d -> part[g, n]; cin_dbias_final_kernel adds the parts in CTA order.  Deterministic.
__global__ void __launch_bounds__(256) cin_dF_dbias_kernel(const float* __restrict__ dout, const float* __restrict__ pre,
                                                           int64_t B, int N, int D, int act, float* __restrict__ dF,
                                                           float* __restrict__ part) {
  extern __shared__ float sm_plane[];
  const int plane = N * D;
  for (int i = threadIdx.x; i < plane; i += 256) sm_plane[i] = 0.f;
  const int64_t chunk = (B + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = (int64_t)blockIdx.x * chunk;
  const int64_t b1 = b0 + chunk < B ? b0 + chunk : B;
  for (int64_t b = b0; b < b1; ++b) {
    const int64_t base = b * plane;
    for (int i = threadIdx.x; i < plane; i += 256) {
      const float v = dout[base + i] * act_grad(pre[base + i], act);
      dF[base + i] = v;
      sm_plane[i] += v;
    }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += 256) {
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc += sm_plane[n * D + d];
    part[(int64_t)blockIdx.x * N + n] = acc;
  }
}

__global__ void __launch_bounds__(256) cin_dbias_final_kernel(const float* __restrict__ part, int G, int N,
                                                              float* __restrict__ dbias) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int g = 0; g < G; ++g) acc += part[(int64_t)g * N + n];
  dbias[n] = acc;
}

int cin_dF_dbias(const float* dout, const float* pre, int64_t B, int N, int D, int act, float* dF, float* dbias,
                 float* scratch, size_t scratch_floats, cudaStream_t st) {
  const size_t smem = (size_t)N * D * sizeof(float);
  int64_t G = (int64_t)(scratch_floats / (size_t)N);
  if (G > 4 * RM_NUM_SMS) G = 4 * RM_NUM_SMS;
  if (G > B) G = B;
  if (G >= 1 && smem <= 200 * 1024) {
    RM_SMEM_ATTR_ONCE(smem, cin_dF_dbias_kernel);
    cin_dF_dbias_kernel<<<(unsigned)G, 256, smem, st>>>(dout, pre, B, N, D, act, dF, scratch);
    RM_LAUNCH_CHECK();
    cin_dbias_final_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(scratch, (int)G, N, dbias);
    RM_LAUNCH_CHECK();
    return 0;
  }
  const int64_t total = B * (int64_t)N * D;
  cin_dF_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(dout, pre, total, act, dF);
  RM_LAUNCH_CHECK();
  cin_dbias_kernel<<<N, 256, 0, st>>>(dF, (int)B, N, D, dbias);
  RM_LAUNCH_CHECK();
  return 0;
}

int cin_fwd_simt(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* bias,
                 int64_t B, int m, int H, int D, int N, int act, float* out, float* pre, cudaStream_t st) {
  const int TB = simt_tb(D);
  RM_UNSUPPORTED(TB > 0, "embedding size D must be <= 128");
  const size_t smem = ((size_t)TB * (m + H) * D + CIN_BK * CIN_AS + CIN_BK * CIN_BN) * sizeof(float);
  RM_UNSUPPORTED(smem <= 227 * 1024, "m + H too large for the shared-memory row cache");
  RM_SMEM_ATTR_ONCE(smem, cin_fwd_simt_kernel);
  dim3 grid((unsigned)ceil_div(B, TB), (unsigned)ceil_div(N, CIN_BN));
  cin_fwd_simt_kernel<<<grid, 256, smem, st>>>(x0, bs0, xk, bsk, W, bias, (int)B, m, H, D, N, act, TB, out, pre);
  RM_LAUNCH_CHECK();
  return 0;
}

CinBwdWs cin_bwd_layout(int64_t B, int m, int H, int D, int N, void* base) {
  CinBwdWs w;
  const int q_tiles = (int)ceil_div(H, CIN_BN);
  const int64_t tiles = (int64_t)m * q_tiles * ceil_div(N, CIN_BN);
  int slabs = (int)ceil_div(4 * RM_NUM_SMS, tiles);  // aim for >= 4 waves of CTAs
  if (slabs < 1) slabs = 1;
  if (slabs > B) slabs = (int)(B > 0 ? B : 1);
  w.slab_samples = (int)ceil_div(B > 0 ? B : 1, slabs);
  w.slabs = (int)ceil_div(B > 0 ? B : 1, w.slab_samples);
  char* p = (char*)base;
  size_t off = 0;
  w.dF = (float*)(p + off);
  off += align_up((size_t)B * N * D * 4, 256);
  w.partial = (float*)(p + off);
  off += align_up((size_t)w.slabs * m * H * N * 4, 256);
  w.total = off;
  return w;
}

int cin_bwd_simt(const float* x0, int64_t bs0, const float* xk, int64_t bsk, const float* W, const float* pre,
                 const float* dout, int64_t B, int m, int H, int D, int N, int act, float* dW, float* dbias,
                 float* dx0, float* dxk, int64_t dbsk, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int TB = simt_tb(D);
  RM_UNSUPPORTED(TB > 0, "embedding size D must be <= 128");
  CinBwdWs ws = cin_bwd_layout(B, m, H, D, N, workspace);
  if (workspace_bytes < ws.total) {
    set_error("rm_cin_layer_bwd: workspace %zu < required %zu", workspace_bytes, ws.total);
    return RM_E_WORKSPACE;
  }
  const size_t smem = ((size_t)2 * TB * (m + H) * D + CIN_BK * CIN_AS + CIN_BK * CIN_BN) * sizeof(float);
  RM_UNSUPPORTED(smem <= 227 * 1024, "m + H too large for the backward's shared-memory accumulators");
  {
    const int rc = cin_dF_dbias(dout, pre, B, N, D, act, ws.dF, dbias, ws.partial, (size_t)ws.slabs * m * H * N, st);
    if (rc) return rc;
  }
  RM_SMEM_ATTR_ONCE(smem, cin_bwd_dx_simt_kernel);
  cin_bwd_dx_simt_kernel<<<(unsigned)ceil_div(B, TB), 256, smem, st>>>(x0, bs0, xk, bsk, W, ws.dF, (int)B, m, H, D, N,
                                                                      TB, dx0, dxk, dbsk);
  RM_LAUNCH_CHECK();
  const int q_tiles = (int)ceil_div(H, CIN_BN);
  dim3 grid((unsigned)(m * q_tiles), (unsigned)ceil_div(N, CIN_BN), (unsigned)ws.slabs);
  cin_bwd_dw_simt_kernel<<<grid, 256, 0, st>>>(x0, bs0, xk, bsk, ws.dF, (int)B, m, H, D, N, q_tiles, ws.slab_samples,
                                               ws.partial);
  RM_LAUNCH_CHECK();
  const int64_t KN = (int64_t)m * H * N;
  cin_dw_final_kernel<<<grid_for(KN, 256, 8), 256, 0, st>>>(ws.partial, KN, ws.slabs, dW);
  RM_LAUNCH_CHECK();
  return 0;
}

}  // namespace rm
