"""torch.autograd.Function shells around the C-ABI kernels.

Embedding-table gradients never become dense tensors: K2 (sort -> segment sum) produces
``ops.SparseGrad`` objects that are attached to the table parameter (``param.rm_sparse_grads``) and the
autograd engine receives ``None`` for the table.  The dense part of a table gradient, if any (the
reference's whole-table L2 term, layers.py:188-193), flows through ordinary autograd into ``param.grad``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
from torch.autograd import Function

from . import _C, ops


def attach_sparse_grad(param: torch.Tensor, sg: ops.SparseGrad) -> None:
    lst = getattr(param, "rm_sparse_grads", None)
    if lst is None:
        lst = []
        param.rm_sparse_grads = lst
    lst.append(sg)


def pop_sparse_grads(param: torch.Tensor) -> List[ops.SparseGrad]:
    lst = getattr(param, "rm_sparse_grads", None) or []
    param.rm_sparse_grads = []
    return lst


def dense_table_grad(param: torch.Tensor, consume: bool = False) -> torch.Tensor:
    """Dense view of everything a table received this step (tests, small tables): param.grad + scattered rows."""
    total = param.shape[0]
    out = torch.zeros(total, param.numel() // total, dtype=param.dtype, device=param.device)
    if param.grad is not None:
        out += param.grad.reshape(total, -1)
    for sg in (pop_sparse_grads(param) if consume else getattr(param, "rm_sparse_grads", None) or []):
        out += sg.to_dense(total)
    return out.reshape(param.shape)


# --------------------------------------------------------------------------- #
# embedding layer (A1-A4)
# --------------------------------------------------------------------------- #
@dataclass
class SparseRun:
    """A run of consecutive one-hot fields: output columns [col, col+n), id columns [id_col, id_col+n)."""

    col: int
    n: int
    id_col: int
    offsets: torch.Tensor  # int64 [n+1] device: global row offset of each field's table


@dataclass
class MultiField:
    col: int
    row_offset: int
    rows: int
    csr_index: int  # which (values, offsets) pair


@dataclass
class EmbeddingLayout:
    m: int
    k: int
    total_rows: int
    runs: List[SparseRun] = field(default_factory=list)
    multi: List[MultiField] = field(default_factory=list)


def _run_ids(sparse_ids: torch.Tensor, run: SparseRun) -> torch.Tensor:
    if run.id_col == 0 and run.n == sparse_ids.shape[1]:
        return sparse_ids if sparse_ids.is_contiguous() else sparse_ids.contiguous()
    return sparse_ids[:, run.id_col : run.id_col + run.n].contiguous()


def _expand_csr(values, offsets, B):
    counts = offsets[1:] - offsets[:-1]
    sample = torch.repeat_interleave(torch.arange(B, device=values.device), counts)
    scale = torch.rsqrt(counts.clamp_min(1).to(torch.float32))
    return sample, scale[sample]


class EmbeddingLayerFunction(Function):
    """FeatEmbeddingLayer: all fields -> embeds [B,m,k] (+ bias [B,m]); backward = deterministic scatter-add."""

    @staticmethod
    def forward(ctx, table, bias_table, layout: EmbeddingLayout, status, sparse_ids, *csr):
        B = sparse_ids.shape[0] if sparse_ids is not None else csr[1].numel() - 1
        m, k = layout.m, layout.k
        out = torch.empty(B, m * k, dtype=torch.float32, device=table.device)
        bias_out = torch.empty(B, m, dtype=torch.float32, device=table.device) if bias_table is not None else None
        run_ids = []
        for run in layout.runs:
            ids = _run_ids(sparse_ids, run)
            run_ids.append(ids)
            ops.gather(table, run.offsets, ids, out=out[:, run.col * k :], status=status)
            if bias_out is not None:
                ops.gather(bias_table.reshape(-1, 1), run.offsets, ids, out=bias_out[:, run.col :], status=status)
        for mf in layout.multi:
            values, offsets = csr[2 * mf.csr_index], csr[2 * mf.csr_index + 1]
            ops.gather_pooled(table, mf.row_offset, mf.rows, values, offsets, out=out[:, mf.col * k :], status=status)
            if bias_out is not None:
                ops.gather_pooled(bias_table.reshape(-1, 1), mf.row_offset, mf.rows, values, offsets,
                                  out=bias_out[:, mf.col :], status=status)
        ctx.layout = layout
        ctx.status = status
        ctx.table = table
        ctx.bias_table = bias_table
        ctx.run_ids = run_ids
        ctx.csr = csr
        ctx.B = B
        embeds = out.view(B, m, k)
        if bias_out is None:
            return embeds
        return embeds, bias_out.view(B, m, 1)

    @staticmethod
    def backward(ctx, g_embeds, g_bias=None):
        layout: EmbeddingLayout = ctx.layout
        B, m, k = ctx.B, layout.m, layout.k
        table, bias_table = ctx.table, ctx.bias_table
        ge = g_embeds.reshape(B, m * k)
        if not ge.is_contiguous():
            ge = ge.contiguous()
        gb = None
        if g_bias is not None and bias_table is not None:
            gb = g_bias.reshape(B, m)
            if not gb.is_contiguous():
                gb = gb.contiguous()
        for run, ids in zip(layout.runs, ctx.run_ids):
            plan = ops.segment_plan(ids, run.offsets, layout.total_rows, status=ctx.status)
            rows = ops.segment_reduce(ge[:, run.col * k :], plan, k, ld=m * k)
            attach_sparse_grad(table, ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique))
            if gb is not None:
                brow = ops.segment_reduce(gb[:, run.col :], plan, 1, ld=m)
                attach_sparse_grad(bias_table, ops.SparseGrad(plan.uniq_rows, brow.reshape(-1), plan.n_unique))
        for mf in layout.multi:
            values, offsets = ctx.csr[2 * mf.csr_index], ctx.csr[2 * mf.csr_index + 1]
            if values.numel() == 0:
                continue
            sample, scale = _expand_csr(values, offsets, B)
            # a tag id outside its table gets the sentinel key (dropped by the plan) instead of a foreign row
            inside = (values >= 0) & (values < mf.rows)
            keys = torch.where(inside, values + mf.row_offset, torch.full_like(values, layout.total_rows)).contiguous()
            plan = ops.segment_plan(keys, None, layout.total_rows, status=ctx.status)
            grows = (ge[:, mf.col * k : (mf.col + 1) * k][sample] * scale[:, None]).contiguous()
            rows = ops.segment_reduce(grows, plan, k, ld=k)
            attach_sparse_grad(table, ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique))
            if gb is not None:
                brows = (gb[:, mf.col][sample] * scale).contiguous()
                brow = ops.segment_reduce(brows, plan, 1, ld=1)
                attach_sparse_grad(bias_table, ops.SparseGrad(plan.uniq_rows, brow.reshape(-1), plan.n_unique))
        return (None,) * (5 + len(ctx.csr))


# --------------------------------------------------------------------------- #
# FM (A5)
# --------------------------------------------------------------------------- #
class FMFunction(Function):
    @staticmethod
    def forward(ctx, embeds, bias):
        out, S = ops.fm_fwd(embeds, bias)
        ctx.save_for_backward(embeds, S)
        ctx.has_bias = bias is not None
        ctx.bias_shape = None if bias is None else bias.shape
        return out.reshape(-1, 1)

    @staticmethod
    def backward(ctx, gout):
        embeds, S = ctx.saved_tensors
        de, dbias = ops.fm_bwd(embeds, S, gout.reshape(-1).contiguous(), want_bias=ctx.has_bias)
        return de, (dbias.reshape(ctx.bias_shape) if ctx.has_bias else None)


# --------------------------------------------------------------------------- #
# cross network (A6)
# --------------------------------------------------------------------------- #
class CrossFunction(Function):
    """x may be a padded row buffer [B, ld] of which the first d columns are the input (d=None: all)."""

    @staticmethod
    def forward(ctx, x, w, b, w_out, w0_out, d=None):
        xin = x if d is None else x[:, :d]
        logit, dots = ops.cross_fwd(xin, w, b, w_out.reshape(-1), w0_out)
        ctx.save_for_backward(x, w, b, w_out, dots)
        ctx.d = d
        return logit.reshape(-1, 1)

    @staticmethod
    def backward(ctx, gout):
        x, w, b, w_out, dots = ctx.saved_tensors
        d = ctx.d
        xin = x if d is None else x[:, :d]
        dx_full = torch.empty_like(x, memory_format=torch.contiguous_format)
        dx_view = dx_full if d is None else dx_full[:, :d]
        if d is not None and x.shape[1] > d:
            dx_full[:, d:].zero_()
        _, dw, db, dwo, dw0 = ops.cross_bwd(xin, w, b, w_out.reshape(-1), dots, gout.reshape(-1).contiguous(),
                                            dx=dx_view, accumulate=False)
        return dx_full, dw, db, dwo.reshape(w_out.shape), dw0, None


# --------------------------------------------------------------------------- #
# CIN layer (A7)
# --------------------------------------------------------------------------- #
class CINLayerFunction(Function):
    """One CIN layer: (x0 [B,m,D], xk [B,H,D], W [m*H,N], bias [N]) -> act(Z.W + bias) as [B,N,D]."""

    @staticmethod
    def forward(ctx, x0, xk, W, bias, act, precision):
        out, pre = ops.cin_layer_fwd(x0, xk, W, bias, act, precision, want_pre=True)
        ctx.save_for_backward(x0, xk, W, pre)
        ctx.act, ctx.precision = act, precision
        return out

    @staticmethod
    def backward(ctx, dout):
        x0, xk, W, pre = ctx.saved_tensors
        dx0 = torch.zeros_like(x0, memory_format=torch.contiguous_format)
        dxk = torch.empty(xk.shape, dtype=xk.dtype, device=xk.device)
        dW, dbias = ops.cin_layer_bwd(x0, xk, W, pre, dout, ctx.act, ctx.precision, dx0, dxk)
        return dx0, dxk, dW, dbias, None, None


class CINLayerPoolFunction(Function):
    """One CIN layer with its split-half + sum-pool (layers.py:711-751):
    (x0, xk, W, bias) -> (next = out[:, :n0] [B,n0,D] feeding the next layer, pooled [B, N-n0] = sum_d out[:, n0:]).
    n0 = 0 for the last layer (``next`` is then an empty tensor).  The backward builds dout in one pass from the two
    gradients instead of autograd's zero-fill + slice-add + expand chain."""

    @staticmethod
    def forward(ctx, x0, xk, W, bias, act, precision, n0):
        out, pre = ops.cin_layer_fwd(x0, xk, W, bias, act, precision, want_pre=True)
        pooled = ops.cin_pool_fwd(out, n0)
        ctx.save_for_backward(x0, xk, W, pre)
        ctx.act, ctx.precision, ctx.n0 = act, precision, n0
        ctx.set_materialize_grads(False)
        return out[:, :n0], pooled

    @staticmethod
    def backward(ctx, d_next, d_pool):
        x0, xk, W, pre = ctx.saved_tensors
        B, N, D = pre.shape
        if d_next is not None and (d_next.stride(2) != 1 or d_next.stride(1) != D):
            d_next = d_next.contiguous()
        dout = ops.cin_pool_bwd(d_next if ctx.n0 else None, d_pool, B, N, D, ctx.n0, pre.device)
        dx0 = torch.zeros_like(x0, memory_format=torch.contiguous_format)
        dxk = torch.empty(xk.shape, dtype=xk.dtype, device=xk.device)
        dW, dbias = ops.cin_layer_bwd(x0, xk, W, pre, dout, ctx.act, ctx.precision, dx0, dxk)
        return dx0, dxk, dW, dbias, None, None, None


# --------------------------------------------------------------------------- #
# fused DeepFM / DCN / xDeepFM front end (K1 + K3 + first-order, one launch)
# --------------------------------------------------------------------------- #
class FmBack:
    """Side channel between DeepFM's front end and its first (narrow) MLP layer: with both at hand, the layer's
    input-gradient kernel adds the FM backward term in its epilogue and writes the complete embedding-row gradients
    G [B*m, k] directly (rm_linear_bwd_input_fm) - the front end's backward then only reduces G.  ``g_fm`` arrives through
    a tensor hook on the FM logit, which autograd fires before it walks back into the MLP; when it has not (other graph
    shapes), both sides fall back to their separate kernels."""

    def __init__(self):
        # no reference to the row buffer itself: it is an output of the front-end Function whose ctx holds this object,
        # and such a cycle would leave the step's tensors to the garbage collector (fatal inside CUDA-graph capture)
        self.x_ptr = 0
        self.S = self.g_fm = self.G = None
        self.m = self.k = 0
        self.alloc = None  # () -> G buffer [B*m, k] (peer-shared memory when the tables are row-sharded)

    def on_g_fm(self, g):
        self.g_fm = g.reshape(-1).contiguous()


class GradTap(Function):
    """Identity on the final logit whose backward hands d(loss)/d(logit) to an ``FmBack`` before autograd walks into
    the MLP (a plain node instead of a tensor hook: hooks break CUDA-graph capture of the step)."""

    @staticmethod
    def forward(ctx, logit, fm_back):
        ctx.fm_back = fm_back
        return logit.view_as(logit)

    @staticmethod
    def backward(ctx, g):
        ctx.fm_back.on_g_fm(g)
        return g, None


class FrontEndFunction(Function):
    """ids -> (xbuf [B, ld] = [embeds | dense | 0-pad], fm [B,1], lin [B,1]).

    ``xbuf`` rows are 16-byte aligned (ld % 4 == 0); consumers read ``xbuf[:, :d]``.  Backward takes
    d(xbuf) [B, ld] (only the first m*k columns are used), d(fm), d(lin) and runs the fused K2.

    ``lin_table`` / ``lin_dense`` are (detached) views of the ``linear_w`` parameter ``W_lin`` - id rows in the
    embedding table's row numbering, then one weight per dense feature.  Their gradients go back to
    ``W_lin`` itself: sparse rows for the id part, ``W_lin.rm_dense_tail = (first_row, grad)`` for the tail.
    """

    @staticmethod
    def forward(ctx, table, bias_table, W_lin, lin_table, lin_dense, offsets, total_rows, status, ids, dense,
                fused_opt=None, fm_back=None):
        x, fm, lin, S = ops.gather_fm_fwd(table, bias_table, lin_table, offsets, ids, dense, lin_dense, status=status)
        ctx.fm_back = fm_back
        if fm_back is not None:
            B_, m_ = ids.shape
            k_ = table.shape[1]
            fm_back.x_ptr, fm_back.S, fm_back.m, fm_back.k = x.data_ptr(), S, m_, k_
            fm_back.g_fm = fm_back.G = None
            dev_ = x.device
            fm_back.alloc = lambda: torch.empty(B_ * m_, k_, dtype=torch.float32, device=dev_)
        ctx.table, ctx.bias_table, ctx.W_lin = table, bias_table, W_lin
        ctx.has_lin = lin_table is not None
        ctx.n_lin_dense = 0 if lin_dense is None else lin_dense.numel()
        ctx.total_rows = total_rows
        ctx.fused_opt = fused_opt  # (opt kind, lr): apply the sparse update inside the backward kernel (N1)
        ctx.lin_table = lin_table
        # K2's plan needs the ids only: build it on the side stream, under the front-end kernel and the MLP forward
        ctx.status = status
        ctx.plan = (ops.segment_plan(ids, offsets, total_rows, side=True, status=status)
                    if any(ctx.needs_input_grad) else None)
        ctx.save_for_backward(x, S, ids, offsets, dense)
        ctx.set_materialize_grads(False)
        return x, fm.reshape(-1, 1), lin.reshape(-1, 1)

    @staticmethod
    def backward(ctx, dx, dfm, dlin):
        x, S, ids, offsets, dense = ctx.saved_tensors
        k = ctx.table.shape[1]
        ld = x.shape[1]
        plan, ctx.plan = ctx.plan, None
        if plan is None:
            plan = ops.segment_plan(ids, offsets, ctx.total_rows, status=ctx.status)
        if dx is not None and (dx.stride(1) != 1 or dx.stride(0) != ld):
            dx = dx.contiguous()
        g_fm = None if dfm is None else dfm.reshape(-1).contiguous()
        g_lin = None if dlin is None else dlin.reshape(-1).contiguous()
        want_bias = ctx.bias_table is not None and g_fm is not None
        want_lin = ctx.has_lin and g_lin is not None
        fb = ctx.fm_back
        if fb is not None and fb.G is not None:
            # the first MLP layer already wrote the complete row gradients (dx + FM backward): reduce them
            G, fb.G = fb.G, None
            B_, m_ = ids.shape
            zeros = torch.zeros(B_, dtype=torch.float32, device=x.device) if (g_fm is None or g_lin is None) else None
            gscal = torch.stack([g_fm if g_fm is not None else zeros, g_lin if g_lin is not None else zeros], 1).contiguous()
            if want_lin and ctx.W_lin is not None and ctx.n_lin_dense and dense is not None:
                ctx.W_lin.rm_dense_tail = (ctx.total_rows, dense.t() @ g_lin)
            if ctx.fused_opt is not None:
                kind, lr = ctx.fused_opt
                ops.segment_reduce_p2p_update([G.data_ptr()], B_ * m_, k, k, plan, ctx.table.data,
                                              ctx.bias_table.data if want_bias else None,
                                              ctx.lin_table if want_lin else None, kind, lr, gscal=gscal, m=m_)
                return (None,) * 12
            rows, ob, ol = ops.segment_reduce_p2p([G.data_ptr()], B_ * m_, k, k, plan, want_bias, want_lin, gscal=gscal,
                                                  m=m_)
            attach_sparse_grad(ctx.table, ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique))
            if want_bias:
                attach_sparse_grad(ctx.bias_table, ops.SparseGrad(plan.uniq_rows, ob, plan.n_unique))
            if want_lin and ctx.W_lin is not None:
                attach_sparse_grad(ctx.W_lin, ops.SparseGrad(plan.uniq_rows, ol, plan.n_unique))
            return (None,) * 12
        if ctx.fused_opt is not None:
            kind, lr = ctx.fused_opt
            ops.emb_fm_bwd_update(dx, x, ld, S, g_fm, g_lin if want_lin else None, plan, k, ctx.table.data,
                                  ctx.bias_table.data if want_bias else None,
                                  ctx.lin_table if want_lin else None, kind, lr)
            if want_lin and ctx.W_lin is not None and ctx.n_lin_dense and dense is not None:
                ctx.W_lin.rm_dense_tail = (ctx.total_rows, dense.t() @ g_lin)
            return (None,) * 12
        rows, ob, ol = ops.emb_fm_bwd(dx, x, ld, S, g_fm, g_lin if want_lin else None, plan, k, True, want_bias,
                                      want_lin)
        attach_sparse_grad(ctx.table, ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique))
        if want_bias:
            attach_sparse_grad(ctx.bias_table, ops.SparseGrad(plan.uniq_rows, ob, plan.n_unique))
        if want_lin and ctx.W_lin is not None:
            attach_sparse_grad(ctx.W_lin, ops.SparseGrad(plan.uniq_rows, ol, plan.n_unique))
            if ctx.n_lin_dense and dense is not None:
                ctx.W_lin.rm_dense_tail = (ctx.total_rows, dense.t() @ g_lin)
        return (None,) * 12


class FirstLinearFunction(Function):
    """y = xbuf[:, :d] @ W + b with the input gradient written straight into a padded [B, ld] buffer."""

    @staticmethod
    def forward(ctx, xbuf, W, b, d, fm_back=None):
        x = xbuf[:, :d]
        ctx.save_for_backward(xbuf, W)
        ctx.d = d
        ctx.fm_back = fm_back
        return torch.addmm(b, x, W)

    @staticmethod
    def backward(ctx, g):
        xbuf, W = ctx.saved_tensors
        d = ctx.d
        g = g.contiguous()
        db = g.sum(0)
        ld = xbuf.shape[1]
        if ops.narrow_linear_ok(g.shape[1]) and xbuf.is_contiguous() and W.is_contiguous():
            # narrow first layer (DeepFM default 32 units): the batch is the only large dimension of both products
            dW = ops.linear_bwd_weight(xbuf, ld, d, g)
            fb = ctx.fm_back
            if (fb is not None and fb.g_fm is not None and fb.x_ptr == xbuf.data_ptr() and fb.k % 4 == 0
                    and fb.g_fm.shape[0] == xbuf.shape[0]):
                # DeepFM: FM backward fused into the epilogue -> complete embedding-row gradients, no dx pass at all
                G = fb.alloc()
                ops.linear_bwd_input_fm(g, W, fb.m, fb.k, xbuf, fb.S, fb.g_fm, out=G)
                fb.G, fb.g_fm = G, None
                return None, dW, db, None, None
            dxbuf = ops.linear_bwd_input(g, W, d_ld=ld)
            return dxbuf, dW, db, None, None
        dW = xbuf[:, :d].t() @ g
        dxbuf = torch.empty_like(xbuf)
        torch.mm(g, W.t(), out=dxbuf[:, :d])
        if xbuf.shape[1] > d:
            dxbuf[:, d:].zero_()
        return dxbuf, dW, db, None, None


class NarrowLinearFunction(Function):
    """y = x @ W + b for a hidden layer with a narrow output (N <= 64): library forward, own batch-reduced backward."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return torch.addmm(b, x, W)

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        db = g.sum(0)
        K = x.shape[1]
        if K % 4 == 0 and x.is_contiguous() and W.is_contiguous():
            dW = ops.linear_bwd_weight(x, K, K, g)
            dx = ops.linear_bwd_input(g, W, d_ld=K) if ctx.needs_input_grad[0] else None
        else:
            dW = x.t() @ g
            dx = g @ W.t() if ctx.needs_input_grad[0] else None
        return dx, dW, db


# --------------------------------------------------------------------------- #
# fused DeepFM tower: front end + first DNN layer on tcgen05, sorted fused backward + update
# --------------------------------------------------------------------------- #
class TowerFunction(Function):
    """ids -> (y1 [B, N1] = [embeds | dense] @ W1 + b1, fm [B,1], lin [B,1]) in one kernel (rm_tower_fwd).

    ``scal`` is the interleaved k=1 storage [rows + n_dense, 2] = (embedding bias, first-order weight); ``bias_param`` /
    ``W_lin`` are the parameters that view its two columns (sparse gradients are attached to them).
    Backward, two modes:
      * ``fused_opt`` given and (k, N1) = (64, 32): rm_tower_bwd_update - the table rows are updated in place inside
        the backward kernel, dx / the row buffer / the summed gradients never exist in HBM;
      * otherwise: the forward also wrote the row buffer and the backward runs the separate kernels
        (rm_linear_bwd_*, rm_emb_fm_bwd) and attaches ``ops.SparseGrad`` s.
    """

    @staticmethod
    def forward(ctx, table, scal, scal_fwd, bias_param, W_lin, W1, b1, offsets, total_rows, status, ids, dense,
                fused_opt, grad_mode=True, side=None):
        k = table.shape[1]
        N1 = W1.shape[1]
        n_dense = 0 if dense is None else dense.shape[1]
        # grad_mode: torch.is_grad_enabled() of the CALLER (inside forward() autograd is always off, and
        # needs_input_grad ignores an enclosing no_grad())
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        use_bk = need_grad and fused_opt is not None and ops.tower_bwd_supported(k, N1)
        lin_dense = scal_fwd[total_rows:, 1] if n_dense else None
        ctx.plan = None
        if need_grad:
            # the plans need the ids only: forked onto the side stream BEFORE the forward kernel is enqueued (the side
            # stream waits for what the main stream holds at the fork), so that the sort runs under the forward
            ctx.plan = (ops.tower_plan(ids, offsets, total_rows, status=status, side=True) if use_bk
                        else ops.segment_plan(ids, offsets, total_rows, side=True, status=status))
        y1, fm, lin, S, x = ops.tower_fwd(table, scal_fwd[:total_rows], offsets, ids, dense, lin_dense, W1, b1,
                                          want_x=need_grad and not use_bk, status=status)
        ctx.use_bk, ctx.fused_opt = use_bk, fused_opt
        ctx.table, ctx.scal, ctx.bias_param, ctx.W_lin = table, scal, bias_param, W_lin
        ctx.total_rows, ctx.n_dense, ctx.status = total_rows, n_dense, status
        ctx.side = side
        ctx.save_for_backward(x, S, ids, dense, W1)
        ctx.set_materialize_grads(False)
        return y1, fm.reshape(-1, 1), lin.reshape(-1, 1)

    @staticmethod
    def backward(ctx, dy1, dfm, dlin):
        x, S, ids, dense, W1 = ctx.saved_tensors
        B, m = ids.shape
        k = ctx.table.shape[1]
        N1 = W1.shape[1]
        dev = ids.device
        plan, ctx.plan = ctx.plan, None
        g1 = torch.zeros(B, N1, dtype=torch.float32, device=dev) if dy1 is None else dy1.contiguous()
        g_fm = torch.zeros(B, dtype=torch.float32, device=dev) if dfm is None else dfm.reshape(-1).contiguous()
        g_lin = None if dlin is None else dlin.reshape(-1).contiguous()
        # the head kernel of the same step may already hold the gradients that need no embedding rows (TowerSide)
        side_db1, side_dW1d, side_dlind = ctx.side.take() if ctx.side is not None else (None, None, None)
        if dy1 is None or (ctx.n_dense and (side_dW1d is None or (g_lin is not None and side_dlind is None))):
            side_db1 = side_dW1d = side_dlind = None
        db1 = side_db1 if side_db1 is not None else g1.sum(0)
        total = ctx.total_rows
        if g_lin is not None and ctx.n_dense:
            ctx.W_lin.rm_dense_tail = (total, side_dlind if side_db1 is not None else dense.t() @ g_lin)
        if ctx.use_bk:
            kind, lr = ctx.fused_opt[:2]
            variant = ctx.fused_opt[2] if len(ctx.fused_opt) > 2 else 0
            ctx.table.rm_hot_flag = plan.unit_bounds[-1:]  # device flag: does the batch hold hot rows (DeepModel reads it)
            dW1 = torch.empty(W1.shape, dtype=torch.float32, device=dev)
            ops.tower_bwd_update(ctx.table.data, ctx.scal[:total], plan, g1, S, g_fm, g_lin, W1.data, kind, lr,
                                 status=ctx.status, out=dW1[: m * k], variant=variant)
            if ctx.n_dense:
                if side_db1 is not None:
                    dW1[m * k :] = side_dW1d
                else:
                    torch.mm(dense.t(), g1, out=dW1[m * k :])
            return (None, None, None, None, None, dW1, db1) + (None,) * 8
        d = m * k + ctx.n_dense
        ld = x.shape[1]
        if ops.narrow_linear_ok(N1):
            dW1 = ops.linear_bwd_weight(x, ld, d, g1)
            dxbuf = ops.linear_bwd_input(g1, W1.data, d_ld=ld)
        else:
            dW1 = x[:, :d].t() @ g1
            dxbuf = torch.zeros_like(x)
            torch.mm(g1, W1.t(), out=dxbuf[:, :d])
        rows, ob, ol = ops.emb_fm_bwd(dxbuf, x, ld, S, g_fm, g_lin, plan, k, True, True, g_lin is not None)
        attach_sparse_grad(ctx.table, ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique))
        attach_sparse_grad(ctx.bias_param, ops.SparseGrad(plan.uniq_rows, ob, plan.n_unique))
        if g_lin is not None:
            attach_sparse_grad(ctx.W_lin, ops.SparseGrad(plan.uniq_rows, ol, plan.n_unique))
        return (None, None, None, None, None, dW1, db1) + (None,) * 8


class TowerSide:
    """Side channel from the head kernel to the tower's backward (same step): the gradients of the first layer that do
    not need the embedding rows - db1 = sum_b g1, dW1_dense = dense^T g1, dlin_dense = dense^T g - come out of
    rm_deepfm_head's partial sums, so the tower's backward does not launch a reduce, a skinny GEMM and a GEMV for them.
    Filled by HeadFunction, read (and cleared) by TowerFunction / P2PTowerFunction; empty = compute them from g1."""

    __slots__ = ("db1", "dW1_dense", "dlin_dense")

    def __init__(self):
        self.clear()

    def clear(self):
        self.db1 = self.dW1_dense = self.dlin_dense = None

    def take(self):
        out = (self.db1, self.dW1_dense, self.dlin_dense)
        self.clear()
        return out


class HeadFunction(Function):
    """DeepFM head in one kernel (rm_deepfm_head): (y1, fm, lin, w0, W2, b2, w3, b3, labels) -> (loss, logit, pred).

    The kernel computes the loss AND every gradient in the same pass; ``backward`` only scales them by the incoming
    gradient of the loss (a scalar) - or not at all with ``unit_grad`` (the caller guarantees that it backpropagates
    d(loss) = 1, as ``DeepModel._eager_step`` does on one GPU: saves six elementwise launches per step).
    ``dense`` + ``side``: see TowerSide."""

    @staticmethod
    def forward(ctx, y1, fm, lin, w0, W2, b2, w3, b3, labels, act, task, dense=None, side=None, unit_grad=False):
        out = ops.deepfm_head(y1.contiguous(), fm.reshape(-1).contiguous(), lin.reshape(-1).contiguous(), w0, W2, b2,
                              w3.reshape(-1).contiguous(), b3, labels, act, task,
                              dense=dense if side is not None else None)
        ctx.out = out
        ctx.side, ctx.unit_grad = side, bool(unit_grad)
        ctx.w3_shape = w3.shape
        ctx.fm_shape, ctx.lin_shape = fm.shape, lin.shape
        ctx.mark_non_differentiable(out["logit"], out["pred"])
        return out["loss"].reshape(()), out["logit"], out["pred"]

    @staticmethod
    def backward(ctx, gloss, _glogit, _gpred):
        o, ctx.out = ctx.out, None
        side, ctx.side = ctx.side, None
        tail = (None,) * 6
        if ctx.unit_grad:
            if side is not None:
                side.db1, side.dW1_dense, side.dlin_dense = o["db1"], o.get("dW1_dense"), o.get("dlin_dense")
            g = o["g"]
            return (o["g1"], g.reshape(ctx.fm_shape), g.reshape(ctx.lin_shape), o["dscal"], o["dW2"], o["db2"],
                    o["dw3"].reshape(ctx.w3_shape), o["dscal"]) + tail
        if side is not None:
            side.db1 = o["db1"] * gloss
            side.dW1_dense = None if "dW1_dense" not in o else o["dW1_dense"] * gloss
            side.dlin_dense = None if "dlin_dense" not in o else o["dlin_dense"] * gloss
        g = o["g"] * gloss
        ds = o["dscal"] * gloss
        return (o["g1"].mul_(gloss), g.reshape(ctx.fm_shape), g.reshape(ctx.lin_shape), ds, o["dW2"] * gloss,
                o["db2"] * gloss, (o["dw3"] * gloss).reshape(ctx.w3_shape), ds) + tail
