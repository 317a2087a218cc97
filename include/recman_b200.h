/*
 * recman_b200.h - C ABI of librecman_b200.so: the B200 (sm_100a) kernels behind
 * the recman.th CTR hot path.
 *
 * The reference (dev-wei/recman) has no FFI / plugin / operator registry: its
 * only boundary is the Python layer vocabulary of recman/tf/core/layers.py,
 * whose numerical backend is TensorFlow library ops.  Every entry point below
 * therefore cites the reference *call site* whose TF op(s) it replaces
 * (file:line under the reference tree).  INTEGRATION.md shows the ctypes
 * binding a recman maintainer would add.
 *
 * Conventions
 *   - plain `extern "C"`, raw pointers + explicit sizes/strides, no torch types;
 *   - every pointer is a DEVICE pointer unless its comment says "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - the caller owns every buffer including workspaces (`*_workspace_bytes`
 *     queries); no hidden allocation, no hidden synchronisation, no environment
 *     variables, no mutable state that results depend on (the library keeps a
 *     launch counter and per-kernel "shared-memory opt-in done" flags) ->
 *     thread-safe per (stream, workspace);
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = RM_E_* below;
 *     `rm_last_error()` returns a thread-local message for the last failure;
 *   - there is no CPU path and no other-architecture path: `rm_device_check`
 *     refuses anything but compute capability 10.x.
 *   - all floating point is fp32 without fast-math (the one exception: the optimizer update's division is the
 *     2-ulp __fdividef, see csrc/optim.cuh), ids are int64 as in the
 *     reference (tf/inputs.py:158).
 */
#ifndef RECMAN_B200_H_
#define RECMAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RM_ABI_VERSION 4 /* 3: fused DeepFM tower (rm_tower_*), head (rm_deepfm_head); 4: rm_cin_pool_*, dense-feature gradients in rm_deepfm_head */

#define RM_E_INVALID (-1)     /* bad argument (null pointer, negative size, ...) */
#define RM_E_UNSUPPORTED (-2) /* shape outside what the kernels implement */
#define RM_E_ARCH (-3)        /* device is not sm_100 */
#define RM_E_WORKSPACE (-4)   /* workspace too small */

/* optimizer kinds: tf/core/utils.py:201-213 (create_optimizer) */
#define RM_OPT_ADAM 0
#define RM_OPT_ADAGRAD 1
#define RM_OPT_GD 2
#define RM_OPT_NONE (-1) /* rm_tower_bwd_update only: compute gradients, leave the tables untouched */
/* rm_tower_bwd_update `variant`: AUTO launches both kernel variants and lets the plan's hot-row flag pick one on the
 * device; PLAIN / HOT launch one only (a caller that knows its id distribution).  Same results either way. */
#define RM_TOWER_BWD_AUTO 0
#define RM_TOWER_BWD_PLAIN 1
#define RM_TOWER_BWD_HOT 2

/* activation kinds for rm_cin_* : hparams/xDeepFM.py:29,33 (tf.nn.leaky_relu, alpha 0.2) */
#define RM_ACT_IDENTITY 0
#define RM_ACT_RELU 1
#define RM_ACT_LEAKY_RELU 2

/* CIN contraction precision: tcgen05 has no fp32 MMA. */
#define RM_CIN_FP32_SIMT 0 /* CUDA-core fp32 FMA (verification path)             */
#define RM_CIN_3XTF32 1    /* tcgen05 kind::tf32, hi/lo split, 3 MMAs (parity)   */
#define RM_CIN_TF32 2      /* tcgen05 kind::tf32 single pass (fast, ~1e-3)        */

int rm_version(void);
const char* rm_last_error(void);
/* 0 when `device` is a compute-capability-10.x GPU, RM_E_ARCH otherwise. */
int rm_device_check(int device);
/* Kernels launched by this library in this process so far (every own <<<>>> site
 * counts 1, every cub:: device primitive call counts 1).  bench.py's gpu_launches. */
int64_t rm_launch_count(void);

/* ------------------------------------------------------------------------- *
 * K1  multi-table embedding gather (forward)
 * replaces: tf.nn.embedding_lookup per field + tf.concat(axis=1)
 *           recman/tf/core/layers.py:117-128 and :238-261
 * All m tables live in one [total_rows, k] array; field f owns rows
 * [table_offsets[f], table_offsets[f+1]).  out[b*out_stride + f*k + c] =
 * table[table_offsets[f] + ids[b*m+f]][c].  One launch for all fields; the
 * concat is never materialised separately.  k == 1 is the bias / first-order
 * weight lookup (layers.py:124-128, :418-439).
 * An id outside its table writes zeros and sets *status != 0 (status may be NULL).
 * ------------------------------------------------------------------------- */
int rm_gather_fwd(const float* table, const int64_t* table_offsets, const int64_t* ids,
                  int64_t B, int32_t m, int32_t k, float* out, int64_t out_stride,
                  int32_t* status, void* stream);

/* K1 pooled variant: tf.nn.embedding_lookup_sparse(combiner="sqrtn")
 * recman/tf/core/layers.py:144-169.  CSR: sample b pools
 * values[offsets[b]..offsets[b+1]) of the table starting at row `row_offset`
 * with `table_rows` rows: out[b*out_stride + c] = sum_j row_j[c] / sqrt(n_b)
 * (n_b == 0 -> zeros), accumulated in CSR order. */
int rm_gather_pooled_fwd(const float* table, int64_t row_offset, int64_t table_rows,
                         const int64_t* values, const int64_t* offsets, int64_t B, int32_t k,
                         float* out, int64_t out_stride, int32_t* status, void* stream);

/* ------------------------------------------------------------------------- *
 * K3  FM layer: recman/tf/core/layers.py:457-478 (FMLayer.__call__)
 * embeds[b*ld + f*k + c], bias[b*m + f] (NULL = no first-order term).
 * out[b] = sum_f bias + 0.5*sum_c[(sum_f e)^2 - sum_f e^2];
 * sum_out (nullable) [B,k] receives sum_f e for the backward.
 * ------------------------------------------------------------------------- */
int rm_fm_fwd(const float* embeds, int64_t ld, const float* bias, int64_t B, int32_t m, int32_t k,
              float* out, float* sum_out, void* stream);
/* backward: d_embeds[b,f,:] (+)= gout[b]*(S[b,:] - e[b,f,:]); d_bias[b,f] = gout[b].
 * `sum` may be NULL (S is recomputed).  accumulate != 0 adds into d_embeds. */
int rm_fm_bwd(const float* embeds, int64_t ld, const float* sum, const float* gout, int64_t B,
              int32_t m, int32_t k, float* d_embeds, int64_t d_ld, float* d_bias, int32_t accumulate,
              void* stream);

/* ------------------------------------------------------------------------- *
 * K1+K3 fused DeepFM front end (recman/tf/core/DeepFM.py:107-140):
 * gather all fields into the DNN input row  x[b] = [e_0 .. e_{m-1} | dense]
 * (DNNCombiner, layers.py:494-501), and from the rows still in registers
 * produce the FM logit (layers.py:457-478, bias term from `bias_table`
 * [total_rows] if non-NULL) and the first-order linear logit
 * sum_f lin_table[row] + sum_j dense[b,j]*lin_dense[j]  (layers.py:330-347).
 * Nullable: bias_table, lin_table, dense/lin_dense (n_dense = 0), fm_out,
 * lin_out, sum_out.  x row stride `ld` >= m*k + n_dense.
 * ------------------------------------------------------------------------- */
int rm_gather_fm_fwd(const float* table, const float* bias_table, const float* lin_table,
                     const int64_t* table_offsets, const int64_t* ids, const float* dense,
                     const float* lin_dense, int32_t n_dense, int64_t B, int32_t m, int32_t k,
                     float* x, int64_t ld, float* fm_out, float* lin_out, float* sum_out,
                     int32_t* status, void* stream);

/* ------------------------------------------------------------------------- *
 * K2  deterministic sparse embedding-gradient scatter-add
 * replaces: TF autodiff of tf.nn.embedding_lookup (IndexedSlices with duplicate
 *           rows summed) for recman/tf/core/layers.py:117-128.
 * Step 1 (plan, depends on ids only): key[p] = table_offsets[p % m] + ids[p]
 * (table_offsets NULL -> key = ids[p]); stable radix sort of (key, p); run-length
 * encode.  Outputs: sorted_pos[N], seg_start[N+1] (first n_unique+1 valid),
 * uniq_rows[N] ascending (first n_unique valid), n_unique[1].
 * Step 2 (reduce): out_rows[u,:] = sum over j in segment u, ascending position,
 * of grad[(p/m)*ld + (p%m)*k + :], p = sorted_pos[j] - a fixed order, no atomics.
 * N = B*m < 2^31, total_rows < 2^32 - 1.  An id outside its table (or, with table_offsets
 * NULL, outside [0, total_rows)) sets *status |= 1 (status nullable), is keyed with the
 * sentinel total_rows and dropped: it sorts behind every row and never appears in uniq_rows.
 * ------------------------------------------------------------------------- */
size_t rm_segment_plan_workspace_bytes(int64_t N);
int rm_segment_plan(const int64_t* ids, const int64_t* table_offsets, int64_t N, int32_t m,
                    int64_t total_rows, void* workspace, size_t workspace_bytes,
                    int32_t* sorted_pos, int32_t* seg_start, int64_t* uniq_rows, int32_t* n_unique,
                    int32_t* status, void* stream);
int rm_segment_reduce(const float* grad, int64_t ld, int32_t m, int32_t k, int64_t N,
                      const int32_t* sorted_pos, const int32_t* seg_start, const int32_t* n_unique,
                      float* out_rows, void* workspace, size_t workspace_bytes, void* stream);
/* Workspace of every segmented-reduce entry point (rm_segment_reduce, rm_emb_fm_bwd[_update],
 * rm_segment_reduce_p2p[_update]; N = number of positions / plan capacity).  Segments longer
 * than 16 positions (skewed ids: a hot row) are not walked serially: every 32 consecutive
 * positions are summed by one row group and the chunk partials are added in chunk order -
 * still a fixed, position-determined association (oracle/segment.py restates it).  A NULL
 * workspace disables this and walks every segment serially. */
size_t rm_segment_reduce_workspace_bytes(int64_t N, int32_t k);

/* Fused DeepFM embedding backward: the gradient row of position p=(b,f) is
 *   dx[b*ld + f*k + :] + g_fm[b] * (S[b,:] - x[b*ld + f*k + :])      (FM bwd, A5b)
 * and the k=1 tables (FM bias, linear weight) both receive g_fm[b] / g_lin[b].
 * Segment sums go to out_rows [n_unique,k], out_bias[n_unique], out_lin[n_unique]
 * (each nullable).  dx may be NULL (no DNN), g_fm NULL (no FM). */
int rm_emb_fm_bwd(const float* dx, const float* x, int64_t ld, const float* sum, const float* g_fm,
                  const float* g_lin, int32_t m, int32_t k, int64_t N, const int32_t* sorted_pos,
                  const int32_t* seg_start, const int32_t* n_unique, float* out_rows, float* out_bias,
                  float* out_lin, void* workspace, size_t workspace_bytes, void* stream);

/* rm_emb_fm_bwd fused with the optimizer (N1): instead of emitting the summed rows,
 * table[uniq_rows[u],:] (and bias_table / lin_table[uniq_rows[u]] when g_fm / g_lin
 * are given) receive the stateless first-step update of rm_sparse_opt_step in the
 * same pass; bit-identical to rm_emb_fm_bwd followed by rm_sparse_opt_step. */
int rm_emb_fm_bwd_update(const float* dx, const float* x, int64_t ld, const float* sum,
                         const float* g_fm, const float* g_lin, int32_t m, int32_t k, int64_t N,
                         const int32_t* sorted_pos, const int32_t* seg_start,
                         const int64_t* uniq_rows, const int32_t* n_unique, float* table,
                         float* bias_table, float* lin_table, int32_t opt, float lr, float l2,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * K4  DCN cross network.  Call site recman/tf/core/DCN.py:135-137 (the class
 * itself is absent from the reference; arithmetic = arXiv 1708.05123 eq. 3):
 *   x_{l+1} = x0*(x_l . w[l]) + b[l] + x_l,  logit = x_L . w_out + w0_out
 * x[b*ld + :d]; w,b [L,d]; w_out [d]; w0_out [1]; logit [B];
 * dots [B,L] receives s_l = x_l . w_l (saved for the backward).  d <= 2048.
 * ------------------------------------------------------------------------- */
int rm_cross_fwd(const float* x, int64_t ld, const float* w, const float* b, const float* w_out,
                 const float* w0_out, int64_t B, int32_t d, int32_t L, float* logit, float* dots,
                 void* stream);
/* backward.  gout [B].  dx[b*d_ld + :d] (+)= dL/dx.  Parameter gradients are
 * batch-reduced in a fixed order (per-block partials in `workspace`, then one
 * ordered pass): dw, db [L,d]; dw_out [d]; dw0_out [1]. */
size_t rm_cross_bwd_workspace_bytes(int64_t B, int32_t d, int32_t L);
int rm_cross_bwd(const float* x, int64_t ld, const float* w, const float* b, const float* w_out,
                 const float* dots, const float* gout, int64_t B, int32_t d, int32_t L, float* dx,
                 int64_t d_ld, int32_t accumulate, float* dw, float* db, float* dw_out,
                 float* dw0_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * K5  CIN layer: recman/tf/core/layers.py:711-751 (one iteration of the loop)
 *   Z[(b,d),(p,q)] = x0[b,p,d] * xk[b,q,d]           (outer product, p-major)
 *   F = act(Z . W + bias),  W [m*H, N]               (conv1d 1x1 == GEMM)
 *   out[b, n, d] = F[(b,d), n]                       (transpose to [B,N,D])
 * x0[b*x0_bstride + p*D + d], xk[b*xk_bstride + q*D + d] (both may be views into
 * wider rows: x0 into the front-end row buffer, xk into the previous layer's
 * [B,N,D] output).  Z is never materialised.  `pre` (nullable) receives the pre-activation
 * [B,N,D] for the backward.  Split-half + sum-pool: rm_cin_pool_fwd / _bwd below;
 * the cin_w head is a [B, sum H] x [sum H, 1] GEMV left to torch.
 * Parity mode (RM_CIN_3XTF32) accumulates at most 512 k'' per CTA in TMEM (the tensor core's fp32 accumulation
 * rounds toward zero) and adds the splits in fp32 round-to-nearest, in order: 1e-5 of max|F| at any K.
 * ------------------------------------------------------------------------- */
int rm_cin_layer_fwd(const float* x0, int64_t x0_bstride, const float* xk, int64_t xk_bstride, const float* W,
                     const float* bias, int64_t B, int32_t m, int32_t H, int32_t D, int32_t N,
                     int32_t act, int32_t precision, float* out, float* pre, void* workspace,
                     size_t workspace_bytes, void* stream);
size_t rm_cin_layer_workspace_bytes(int64_t B, int32_t m, int32_t H, int32_t D, int32_t N,
                                    int32_t precision);
/* backward of one CIN layer.  dout [B,N,D] is dL/d(out); pre as saved.
 * dW [m*H,N], dbias [N] are batch-reduced deterministically;
 * dx0 [B,m,D] is accumulated (+=) (x0 feeds every layer), dxk [B,H,D]
 * (batch stride dxk_bstride) is written. */
int rm_cin_layer_bwd(const float* x0, int64_t x0_bstride, const float* xk, int64_t xk_bstride, const float* W,
                     const float* pre, const float* dout, int64_t B, int32_t m, int32_t H,
                     int32_t D, int32_t N, int32_t act, int32_t precision, float* dW, float* dbias,
                     float* dx0, float* dxk, int64_t dxk_bstride, void* workspace,
                     size_t workspace_bytes, void* stream);
size_t rm_cin_layer_bwd_workspace_bytes(int64_t B, int32_t m, int32_t H, int32_t D, int32_t N,
                                        int32_t precision);
/* split-half + sum-pool of a layer's output (recman/tf/core/layers.py:738-751): feature maps n < n0 feed the next
 * layer (out[:, :n0] is used in place), pooled[b, n - n0] = sum_d out[b, n, d] for n >= n0 (n0 = 0: last layer).
 * Backward: dout[b,n,d] = n < n0 ? d_next[b*next_bstride + n*D + d] : d_pool[b, n - n0]; a null gradient is zero. */
int rm_cin_pool_fwd(const float* out, int64_t B, int32_t N, int32_t D, int32_t n0, float* pooled, void* stream);
int rm_cin_pool_bwd(const float* d_next, int64_t next_bstride, const float* d_pool, int64_t B, int32_t N, int32_t D,
                    int32_t n0, float* dout, void* stream);

/* ------------------------------------------------------------------------- *
 * N1  optimizer step on K2's (rows, sums) output and on dense parameters.
 * The reference constructs a NEW optimizer every batch
 * (recman/tf/core/xDeepFM.py:116-126) so every step is a first step with
 * zero slots; the kernels implement exactly that (oracle.fresh_optimizer_step).
 * Sparse: for u < *n_unique: table[uniq_rows[u], :] <- step(., rows[u,:] + l2*table).
 * Dense:  p[i] <- step(p[i], g[i] + l2*p[i]).
 * ------------------------------------------------------------------------- */
int rm_sparse_opt_step(float* table, int32_t k, const int64_t* uniq_rows, const float* rows,
                       const int32_t* n_unique, int64_t max_rows, int32_t opt, float lr, float l2,
                       void* stream);
/* The same with rows `row_stride` floats apart (the k = 1 tables interleaved as [rows, 2], see rm_tower_fwd). */
int rm_sparse_opt_step_strided(float* table, int32_t k, int64_t row_stride, const int64_t* uniq_rows,
                               const float* rows, const int32_t* n_unique, int64_t max_rows, int32_t opt,
                               float lr, float l2, void* stream);
int rm_dense_opt_step(float* p, const float* g, int64_t n, int32_t opt, float lr, float l2,
                      void* stream);
/* The same update for `count` tensors in one launch (per 96 tensors): ps / gs / ns are HOST arrays of device
 * pointers and element counts; bit-identical to `count` calls of rm_dense_opt_step. */
int rm_dense_opt_step_multi(float* const* ps, const float* const* gs, const int64_t* ns, int32_t count,
                            int32_t opt, float lr, float l2, void* stream);

/* ------------------------------------------------------------------------- *
 * (e) multi-GPU: row-sharded tables (row r of a table lives on rank r mod W at
 * local row r div W).  Nothing in the single-process reference corresponds to
 * this; these are the staging kernels either side of the NCCL all-to-all of
 * ids / vectors / gradient rows.  A routed row is KP = k+4 floats:
 * [e_0..e_{k-1} | bias | lin | 0 | 0].  pos[j] = b*m + f of routed row j.
 * rm_unpack_rows: x[b*ld + f*k + :k] = recv[j,:k]; bias_out[pos] = recv[j,k];
 *                 lin_out[pos] = recv[j,k+1] (both nullable).
 * rm_pack_grad_rows (pos NULL = identity): send[j,:k] = dx[b*ld+f*k+:] + g_fm[b]*(S[b,:]-x[b*ld+f*k+:]);
 *                 send[j,k] = g_fm[b]; send[j,k+1] = g_lin[b]  (dx, g_fm, g_lin nullable).
 * The owner side uses rm_gather_fwd (out_stride = KP) and rm_segment_plan/reduce
 * (m = 1, k = KP) directly on the exchange buffers.
 * ------------------------------------------------------------------------- */
int rm_unpack_rows(const float* recv, int64_t n, int32_t KP, const int32_t* pos, int32_t m, int32_t k,
                   float* x, int64_t ld, float* bias_out, float* lin_out, void* stream);
int rm_pack_grad_rows(const float* dx, const float* x, int64_t ld, const float* sum, const float* g_fm,
                      const float* g_lin, int64_t n, int32_t KP, const int32_t* pos, int32_t m, int32_t k,
                      float* send, void* stream);

/* ------------------------------------------------------------------------- *
 * A8' backward of a NARROW dense layer y[B,N] = x[B,d] W[d,N] + b, N <= 64, N % 4 == 0
 * (DNN, recman/tf/core/layers.py:576-609; DeepFM's default hidden units are (32, 32),
 * recman/tf/core/DeepFM.py:35).  The batch is the only large dimension of these
 * products; exact fp32 FMA.  The forward and wide layers stay on cuBLAS.
 *   rm_linear_bwd_input : dx[b, c] = sum_n g[b,n] * W[c,n]  for c < d; columns
 *                         d <= c < d_ld (row padding, d_ld % 4 == 0) are set to 0.
 *   rm_linear_bwd_weight: dW[f, n] = sum_b x[b*ld + f] * g[b,n]  for f < K; the
 *                         batch is reduced in fixed slabs summed in slab order
 *                         (deterministic); workspace from the _workspace_bytes query.
 * ------------------------------------------------------------------------- */
int rm_linear_bwd_input(const float* g, int64_t B, int32_t N, const float* W, int32_t d, float* dx, int64_t d_ld,
                        void* stream);
/* rm_linear_bwd_input with DeepFM's FM backward (A5b) fused into the epilogue: the first m*k input columns are the
 * embedding block of the row buffer x[B, x_ld], and
 *   out[b, f*k + j] = sum_n g[b,n] * W[f*k + j, n] + g_fm[b] * (sum[b, j] - x[b*x_ld + f*k + j])
 * is the complete gradient of embedding row (b, f): out [B, m*k] is exactly the [B*m, k] gradient-row buffer that
 * rm_segment_reduce_p2p[_update] (W = 1: this rank only) consumes - no separate dx pass, no pack pass. */
int rm_linear_bwd_input_fm(const float* g, int64_t B, int32_t N, const float* W, int32_t m, int32_t k, const float* x,
                           int64_t x_ld, const float* sum, const float* g_fm, float* out, void* stream);
size_t rm_linear_bwd_weight_workspace_bytes(int64_t B, int32_t K, int32_t N);
int rm_linear_bwd_weight(const float* x, int64_t ld, const float* g, int64_t B, int32_t K, int32_t N, float* dW,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * (e') row-sharded tables over NVLink PEER MEMORY (any W <= 8, one NVSwitch box;
 * powers of two use mask / shift, other sizes one integer division per id): the lookup and the gradient reduction read their rows straight
 * from the owning rank - no all-to-all, no pack/unpack pass, no host sync.
 *
 * rm_p2p_alloc: cudaMalloc + cudaIpcGetMemHandle (handle64: 64 host bytes to be
 *   exchanged between the ranks); rm_p2p_open maps a peer's allocation
 *   (cudaIpcOpenMemHandle, lazy peer access); rm_p2p_close / rm_p2p_free undo.
 * rm_gather_fm_fwd_p2p: rm_gather_fm_fwd where row `id` of field f is read from
 *   tables[id % W] + (local_offsets[f] + id / W) * k  (same for bias / lin).
 *   tables / bias_tables / lin_tables are HOST arrays of W device pointers
 *   (bias_tables, lin_tables nullable); feat_sizes [m] are the GLOBAL table
 *   sizes (range check), local_offsets [m] the owner-local first rows.
 * rm_shard_plan: the owner-side K2 plan.  gids [W*b*m] are the ids of ALL ranks
 *   (rank-major, e.g. an all-gather); entries with id % W == rank are selected
 *   in ascending global position gp = src_rank*(b*m) + p, keyed by owner-local
 *   row, stably sorted and run-length encoded.  The owned count stays on the
 *   device: the sort runs over the fixed capacity N_cap (sentinel keys behind
 *   the live entries); more than N_cap owned entries set *status |= 4.
 *   Outputs as rm_segment_plan (sorted_gpos holds global positions) + n_own[1].
 * rm_segment_reduce_p2p: rm_segment_reduce whose row gp is read from rank
 *   gp / rows_per_rank's gradient buffer G[r] + (gp % rows_per_rank)*KP,
 *   a KP = k+4 float row [g_0..g_{k-1} | g_bias | g_lin | 0 | 0] as written by
 *   rm_pack_grad_rows(pos = NULL).  G is a HOST array of W device pointers.
 *   Summation order = ascending global position, i.e. the single-GPU order of
 *   the concatenated batch, independent of W.
 *   gscal (nullable, LOCAL memory): the two k=1 gradients are per-SAMPLE values, so
 *   instead of riding in every row they can be all-gathered once as
 *   gscal[W * rows_per_rank / m, 2] = (g_bias, g_lin) per global sample; then
 *   KP >= k suffices (KP = k: 256-byte rows at k = 64, one NVLink request fewer
 *   per row) and m (fields per sample) maps a position to its sample.
 * ------------------------------------------------------------------------- */
int rm_p2p_alloc(size_t bytes, void** ptr, uint8_t* handle64);
int rm_p2p_open(const uint8_t* handle64, void** ptr);
int rm_p2p_close(void* ptr);
int rm_p2p_free(void* ptr);
int rm_gather_fm_fwd_p2p(const float* const* tables, const float* const* bias_tables,
                         const float* const* lin_tables, int32_t W, const int64_t* feat_sizes,
                         const int64_t* local_offsets, const int64_t* ids, const float* dense,
                         const float* lin_dense, int32_t n_dense, int64_t B, int32_t m, int32_t k,
                         float* x, int64_t ld, float* fm_out, float* lin_out, float* sum_out,
                         int32_t* status, void* stream);
size_t rm_shard_plan_workspace_bytes(int64_t Ntot, int64_t N_cap);
int rm_shard_plan(const int64_t* gids, int64_t Ntot, int32_t m, int32_t W, int32_t rank,
                  const int64_t* feat_sizes, const int64_t* local_offsets, int64_t total_local,
                  int64_t N_cap, void* workspace, size_t workspace_bytes, int32_t* sorted_gpos,
                  int32_t* seg_start, int64_t* uniq_rows, int32_t* n_unique, int32_t* n_own,
                  int32_t* status, void* stream);
int rm_segment_reduce_p2p(const float* const* G, const float* gscal, int32_t m, int32_t W,
                          int64_t rows_per_rank, int32_t KP, int32_t k, int64_t N_cap, const int32_t* sorted_gpos,
                          const int32_t* seg_start, const int32_t* n_unique, float* out_rows,
                          float* out_bias, float* out_lin, void* workspace, size_t workspace_bytes,
                          void* stream);
/* ... fused with the optimizer update of the owner's local tables (bias_table / lin_table nullable). */
int rm_segment_reduce_p2p_update(const float* const* G, const float* gscal, int32_t m, int32_t W,
                                 int64_t rows_per_rank, int32_t KP, int32_t k, int64_t N_cap, const int32_t* sorted_gpos,
                                 const int32_t* seg_start, const int64_t* uniq_rows,
                                 const int32_t* n_unique, float* table, float* bias_table,
                                 float* lin_table, int32_t opt, float lr, float l2, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * T   fused DeepFM "tower": front end + first DNN layer on tcgen05 (forward), and the
 *     whole sparse backward + optimizer in one sorted pass (backward).
 * replaces, forward : FeatEmbeddingLayer.__call__ layers.py:238-261 (gather), FMLayer
 *     layers.py:457-478, LinearLayer layers.py:330-347 and the first matmul + bias_add of
 *     DNN.__call__ layers.py:589-609, composed as in tf/core/DeepFM.py:107-163;
 * replaces, backward: TF autodiff of the same ops (IndexedSlices gradient of
 *     tf.nn.embedding_lookup with duplicate rows summed) + the per-batch optimizer of
 *     xDeepFM.py:116-126.
 * The two k=1 tables (embedding bias layers.py:124-128, first-order weight :418-439) are ONE
 * interleaved [rows, 2] array `scal` = (bias, weight): one 8-byte lookup per id.
 * Forward: y1[b, :] = [embeds | dense] @ W1 + b1 (pre-activation, 3xTF32 on the tensor
 * core, fp32 result), fm_out / lin_out / sum_out as rm_gather_fm_fwd; the row buffer x is
 * optional (NULL: never written).  k in {32, 64}, N1 <= 64.
 * Backward: rm_tower_plan sorts (row, position) and cuts work units; rm_tower_bwd_update
 * then gathers each touched row once, forms g1 @ W1_f^T + g_fm * (S - row) per position
 * on the tensor core, sums positions of the same row in ascending position order
 * (deterministic), applies the stateless first-step optimizer in place and accumulates
 * dW1[:m*k] = x^T @ g1.  k = 64, N1 = 32.  out_rows / out_scal (nullable): summed
 * gradient row / (bias, weight) gradient written at the sorted position that closes its
 * segment (tests); opt == RM_OPT_NONE skips the update.
 * unit_bounds holds m*(upf+1) cuts (upf = rm_tower_units_per_field(B, unit)) followed by ONE flag
 * word the plan sets when some row collects more than 32 positions of the batch (skewed ids): the
 * backward then runs its variant that sums long runs warp-cooperatively (same ascending order:
 * bit-identical results, ~2x faster under Zipf ids).  Allocate m*(upf+1) + 1 int32.
 * ------------------------------------------------------------------------- */
int rm_tower_supported(int32_t m, int32_t k, int32_t n_dense, int32_t N1);
size_t rm_tower_fwd_workspace_bytes(int32_t m, int32_t k, int32_t N1);
int rm_tower_fwd(const float* table, const float* scal, const int64_t* table_offsets,
                 const int64_t* ids, const float* dense, const float* lin_dense,
                 int32_t lin_dense_stride, int32_t n_dense, const float* W1, const float* b1,
                 int32_t N1, int64_t B, int32_t m, int32_t k, float* x, int64_t ld, float* y1,
                 float* fm_out, float* lin_out, float* sum_out, int32_t* status, void* workspace,
                 size_t workspace_bytes, void* stream);
int32_t rm_tower_units_per_field(int64_t B, int32_t unit);
size_t rm_tower_plan_workspace_bytes(int64_t N);
int rm_tower_plan(const int64_t* ids, const int64_t* table_offsets, int64_t B, int32_t m,
                  int64_t total_rows, int32_t unit, void* workspace, size_t workspace_bytes,
                  uint32_t* sorted_keys, int32_t* sorted_pos, int32_t* field_bounds,
                  int32_t* unit_bounds, int32_t* status, void* stream);
size_t rm_tower_bwd_workspace_bytes(int64_t B, int32_t m, int32_t unit);
int rm_tower_bwd_update(float* table, float* scal, const uint32_t* sorted_keys,
                        const int32_t* sorted_pos, const int32_t* unit_bounds, const float* g1,
                        const float* S, const float* g_fm, const float* g_lin, const float* W1,
                        int64_t B, int32_t m, int32_t k, int32_t N1, int32_t unit, int32_t opt,
                        float lr, float l2, int32_t variant, float* dW1, float* out_rows, float* out_scal,
                        int32_t* status, void* workspace, size_t workspace_bytes, void* stream);
/* Row-sharded tables (row r of a table on rank r mod W at local row r div W) over NVLink peer memory:
 * rm_tower_fwd_p2p = rm_tower_fwd with every row read from its owner (tables / scals: HOST arrays of W device
 * pointers into the ranks' cudaIpc-mapped shards; feat_sizes [m] global sizes, local_offsets [m] owner-local first
 * rows); rm_tower_shard_plan = the owner-side rm_tower_plan over the ids of ALL ranks (gids [W*b*m] int32, rank-major):
 * owned entries in ascending global position gp = src_rank*(b*m) + p, keyed by owner-local row, sorted, cut into
 * units (capacity N_cap, *status |= 4 when exceeded; Bcap sizes the unit grid).  rm_tower_bwd_update then runs on the
 * owner with B = Bcap and g1 / S / g_fm / g_lin holding the all-gathered per-sample operands of all W*b samples. */
int rm_tower_fwd_p2p(const float* const* tables, const float* const* scals, int32_t W,
                     const int64_t* feat_sizes, const int64_t* local_offsets, const int64_t* ids,
                     const float* dense, const float* lin_dense, int32_t lin_dense_stride,
                     int32_t n_dense, const float* W1, const float* b1, int32_t N1, int64_t B, int32_t m,
                     int32_t k, float* y1, float* fm_out, float* lin_out, float* sum_out,
                     int32_t* status, void* workspace, size_t workspace_bytes, void* stream);
size_t rm_tower_shard_plan_workspace_bytes(int64_t Ntot, int64_t N_cap);
int rm_tower_shard_plan(const int32_t* gids, int64_t Ntot, int32_t m, int32_t W, int32_t rank,
                        const int64_t* feat_sizes, const int64_t* local_offsets, int64_t total_local,
                        int64_t N_cap, int64_t Bcap, int32_t unit, void* workspace,
                        size_t workspace_bytes, uint32_t* sorted_keys, int32_t* sorted_gpos,
                        int32_t* field_bounds, int32_t* unit_bounds, int32_t* n_own, int32_t* status,
                        void* stream);

/* ------------------------------------------------------------------------- *
 * H   fused DeepFM head: everything between the first DNN layer's pre-activation y1 and
 *     the loss, forward and backward, for hidden_units = (32, 32).
 * replaces: the rest of DNN.__call__ layers.py:589-609 (activation, second matmul,
 *     output projection), the add_n of the towers tf/core/DeepFM.py:150-160,
 *     PredictionLayer layers.py:796-808, create_loss utils.py:192-198 (Keras BCE with
 *     clip eps = 1e-7, or MSE) and TF's autodiff of all of it.
 * logit = lin + w0 + fm + dnn(y1); pred = sigmoid(logit) (task 0) or logit (task 1).
 * labels == NULL: forward only (logit / pred, nullable).  Otherwise also loss[1] (batch
 * mean), g[B] = grad_scale * dL/dlogit (the FM and first-order gradient), g1[B,32] =
 * grad_scale * dL/dy1, dW2[32,32], db2[32], dw3[32], dscal[1] (= db3 = dw0), db1[32].
 * dense [B, n_dense] (nullable): the samples' dense features; with it the kernel also emits the two
 * gradients of the first layer / first-order term that involve them, dW1_dense[n_dense,32] =
 * dense^T g1 and dlin_dense[n_dense] = dense^T g (layers.py:418-439, 589-602 autodiff).
 * Batch reductions are CTA partials added in a fixed association: run-to-run identical.
 * ------------------------------------------------------------------------- */
int rm_deepfm_head_supported(int32_t N1, int32_t N2);
size_t rm_deepfm_head_workspace_bytes(int64_t B, int32_t n_dense);
int rm_deepfm_head(const float* y1, const float* fm, const float* lin, const float* w0,
                   const float* W2, const float* b2, const float* w3, const float* b3,
                   const float* labels, const float* dense, int32_t n_dense, int64_t B, int32_t N1,
                   int32_t N2, int32_t act, int32_t task, float grad_scale, float* logit,
                   float* pred, float* loss, float* g1, float* g, float* dW2, float* db2, float* dw3,
                   float* dscal, float* db1, float* dW1_dense, float* dlin_dense, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Test-only: D[128,32] = At^T @ Bt (At [K,128], Bt [K,32]) through the MN-major SWIZZLE_128B
 * operand layout of the tower backward's weight-gradient GEMM (variant 0 = the layout used). */
int rm_umma_probe(const float* At, const float* Bt, int32_t K, int32_t variant, float* D,
                  int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RECMAN_B200_H_ */
