"""torch-facing wrappers over the C ABI (``recman_b200._C``).

PyTorch is plumbing here: it owns device memory and streams; every hot op is a
hand-written sm_100a kernel reached through ``librecman_b200.so``.  There is no
fallback: tensors that are not on a CUDA device raise.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _C

_checked_devices = set()


def _dev_check(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise _C.RecmanB200Error("recman_b200 kernels need CUDA tensors (there is no CPU fallback)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        _C.call("rm_device_check", idx)
        _checked_devices.add(idx)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t


def _rows_view(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """Accept a 2-D/3-D tensor whose rows are dense; returns (tensor, row stride in floats)."""
    _f32c(t, name)
    if t.dim() == 3:
        if t.stride(2) != 1 or t.stride(1) != t.shape[2]:
            t = t.contiguous()
        return t, t.stride(0)
    if t.dim() == 2:
        if t.stride(1) != 1:
            t = t.contiguous()
        return t, t.stride(0)
    raise ValueError(f"{name} must be 2-D or 3-D")


def enable_profile():
    _C.enable_profile()


def disable_profile():
    return _C.disable_profile()


def launch_count() -> int:
    return int(_C.lib.rm_launch_count())


def new_status(device) -> torch.Tensor:
    return torch.zeros(1, dtype=torch.int32, device=device)


# --------------------------------------------------------------------------- #
# K1 gather
# --------------------------------------------------------------------------- #
def gather(table, table_offsets, ids, out=None, status=None):
    """[total_rows,k] table, int64 ids [B,m] -> [B, m, k] (or into ``out`` [B, >=m*k])."""
    _dev_check(table)
    _f32c(table, "table")
    assert ids.dtype == torch.int64 and ids.is_contiguous() and table.is_contiguous()
    assert table_offsets.dtype == torch.int64 and table_offsets.is_cuda
    B, m = ids.shape
    k = table.shape[1]
    if out is None:
        out = torch.empty(B, m, k, dtype=torch.float32, device=table.device)
        stride = m * k
    else:
        assert out.dim() == 2 and out.stride(1) == 1
        stride = out.stride(0)
    _C.call("rm_gather_fwd", _p(table), _p(table_offsets), _p(ids), B, m, k, _p(out), stride, _p(status), _stream())
    return out


def gather_pooled(table, row_offset, table_rows, values, offsets, out=None, status=None):
    """sqrtn-pooled CSR lookup -> [B, k]."""
    _dev_check(table)
    assert values.dtype == torch.int64 and offsets.dtype == torch.int64
    B = offsets.numel() - 1
    k = table.shape[1]
    if out is None:
        out = torch.empty(B, k, dtype=torch.float32, device=table.device)
    assert out.stride(-1) == 1
    _C.call(
        "rm_gather_pooled_fwd", _p(table), int(row_offset), int(table_rows), _p(values), _p(offsets), B, k, _p(out),
        out.stride(0), _p(status), _stream(),
    )
    return out


# --------------------------------------------------------------------------- #
# K3 FM
# --------------------------------------------------------------------------- #
def fm_fwd(embeds, bias=None, want_sum=True):
    _dev_check(embeds)
    e, ld = _rows_view(embeds, "embeds")
    B, m, k = embeds.shape
    out = torch.empty(B, dtype=torch.float32, device=e.device)
    S = torch.empty(B, k, dtype=torch.float32, device=e.device) if want_sum else None
    if bias is not None:
        bias = _f32c(bias, "bias").reshape(B, m).contiguous()
    _C.call("rm_fm_fwd", _p(e), ld, _p(bias), B, m, k, _p(out), _p(S), _stream())
    return out, S


def fm_bwd(embeds, S, gout, d_embeds=None, want_bias=True, accumulate=False):
    _dev_check(embeds)
    e, ld = _rows_view(embeds, "embeds")
    B, m, k = embeds.shape
    if d_embeds is None:
        d_embeds = torch.empty(B, m, k, dtype=torch.float32, device=e.device)
        accumulate = False
    de, d_ld = _rows_view(d_embeds, "d_embeds")
    assert de.data_ptr() == d_embeds.data_ptr(), "d_embeds must have dense rows"
    d_bias = torch.empty(B, m, dtype=torch.float32, device=e.device) if want_bias else None
    gout = gout.reshape(B).contiguous()
    _C.call("rm_fm_bwd", _p(e), ld, _p(S), _p(gout), B, m, k, _p(de), d_ld, _p(d_bias), int(accumulate), _stream())
    return d_embeds, d_bias


def gather_fm_fwd(table, bias_table, lin_table, table_offsets, ids, dense, lin_dense, ld=None, status=None):
    """Fused DeepFM front end.  Returns (x_buf [B, ld], fm [B], lin [B], S [B,k])."""
    _dev_check(table)
    B, m = ids.shape
    k = table.shape[1]
    n_dense = 0 if dense is None else dense.shape[1]
    d = m * k + n_dense
    if ld is None:
        ld = (d + 3) // 4 * 4
    dev = table.device
    x = torch.empty(B, ld, dtype=torch.float32, device=dev)
    if ld > d:
        x[:, d:].zero_()
    fm = torch.empty(B, dtype=torch.float32, device=dev)
    lin = torch.empty(B, dtype=torch.float32, device=dev)
    S = torch.empty(B, k, dtype=torch.float32, device=dev)
    if dense is not None:
        dense = _f32c(dense, "dense").contiguous()
    _C.call(
        "rm_gather_fm_fwd", _p(table), _p(bias_table), _p(lin_table), _p(table_offsets), _p(ids), _p(dense),
        _p(lin_dense), n_dense, B, m, k, _p(x), ld, _p(fm), _p(lin), _p(S), _p(status), _stream(),
    )
    return x, fm, lin, S


# --------------------------------------------------------------------------- #
# K2 deterministic scatter-add
# --------------------------------------------------------------------------- #
@dataclass
class SegmentPlan:
    """sort + run-length-encode of the (table,row) keys of one batch (ids only)."""

    N: int
    m: int
    sorted_pos: torch.Tensor  # int32 [N]
    seg_start: torch.Tensor  # int32 [N+1]
    uniq_rows: torch.Tensor  # int64 [N]
    n_unique: torch.Tensor  # int32 [1] (device)
    workspace: torch.Tensor
    ready: Optional["torch.cuda.Event"] = None  # set when the plan was built on the side stream

    def num_unique(self) -> int:
        self.wait()
        return int(self.n_unique.item())  # host sync: tests / densify only

    def wait(self) -> None:
        """Make the current stream wait for a plan that was built on the side stream (no-op otherwise)."""
        if self.ready is not None:
            torch.cuda.current_stream().wait_event(self.ready)
            self.ready = None


_side_streams = {}


def side_stream() -> "torch.cuda.Stream":
    """The side stream paired with the current stream (plans that depend on the ids alone run on it)."""
    cur = torch.cuda.current_stream()
    key = (cur.device.index, cur.cuda_stream)
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=cur.device)
    return st


def _launch_maybe_side(side: bool, launch, tensors=(), fork: bool = True):
    """Run ``launch()`` (kernel launches only, every buffer already allocated) on the current stream, or - the plan
    depends on the ids alone - on a side stream forked from it, so that it overlaps the forward.  Returns the event to
    wait for, or None.  ``tensors``: every buffer the launch touches; each is recorded on the side stream so that the
    caching allocator does not hand its memory to the main stream while the side stream still uses it (an all-gathered
    id buffer dropped at the end of the caller's forward, a plan dropped without a backward)."""
    if not side:
        launch()
        return None
    cur = torch.cuda.current_stream()
    st = side_stream()
    if fork:  # fork=False: the caller already forked the side stream and queued work on it that `launch` depends on
        st.wait_stream(cur)
    with torch.cuda.stream(st):
        launch()
        ev = torch.cuda.Event()
        ev.record(st)
    for t in tensors:
        if t is not None:
            t.record_stream(st)
    return ev


def segment_plan(ids, table_offsets, total_rows, side: bool = False, status=None) -> SegmentPlan:
    _dev_check(ids)
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    if ids.dim() == 2:
        B, m = ids.shape
    else:
        B, m = ids.numel(), 1
    N = B * m
    dev = ids.device
    ws_bytes = _C.lib.rm_segment_plan_workspace_bytes(N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    sorted_pos = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
    seg_start = torch.empty(N + 1, dtype=torch.int32, device=dev)
    uniq_rows = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
    n_unique = torch.empty(1, dtype=torch.int32, device=dev)
    ev = _launch_maybe_side(side, lambda: _C.call(
        "rm_segment_plan", _p(ids), _p(table_offsets), N, m, int(total_rows), _p(ws), ws_bytes, _p(sorted_pos),
        _p(seg_start), _p(uniq_rows), _p(n_unique), _p(status), _stream(),
    ), (ids, table_offsets, ws, sorted_pos, seg_start, uniq_rows, n_unique, status))
    return SegmentPlan(N, m, sorted_pos, seg_start, uniq_rows, n_unique, ws, ev)


def _reduce_ws(plan: SegmentPlan, k: int):
    """Workspace of the chunked long-segment path (hot rows under skewed ids)."""
    n = _C.lib.rm_segment_reduce_workspace_bytes(max(plan.N, 1), k)
    return torch.empty(n, dtype=torch.uint8, device=plan.sorted_pos.device), n


def segment_reduce(grad, plan: SegmentPlan, k: int, ld: Optional[int] = None, out_rows=None):
    """grad rows addressed as grad[(p//m)*ld + (p%m)*k] -> summed rows [N, k] (first n_unique valid)."""
    _dev_check(grad)
    _f32c(grad, "grad")
    if ld is None:
        g2 = grad.reshape(-1, plan.m * k)
        grad = g2 if g2.is_contiguous() else g2.contiguous()
        ld = plan.m * k
    if out_rows is None:
        out_rows = torch.empty(max(plan.N, 1), k, dtype=torch.float32, device=grad.device)
    plan.wait()
    ws, wsn = _reduce_ws(plan, k)
    _C.call(
        "rm_segment_reduce", _p(grad), ld, plan.m, k, plan.N, _p(plan.sorted_pos), _p(plan.seg_start),
        _p(plan.n_unique), _p(out_rows), _p(ws), wsn, _stream(),
    )
    return out_rows


def emb_fm_bwd(dx, x, ld, S, g_fm, g_lin, plan: SegmentPlan, k, want_rows=True, want_bias=False, want_lin=False):
    dev = plan.sorted_pos.device
    n = max(plan.N, 1)
    out_rows = torch.empty(n, k, dtype=torch.float32, device=dev) if want_rows else None
    out_bias = torch.empty(n, dtype=torch.float32, device=dev) if want_bias else None
    out_lin = torch.empty(n, dtype=torch.float32, device=dev) if want_lin else None
    plan.wait()
    ws, wsn = _reduce_ws(plan, k)
    _C.call(
        "rm_emb_fm_bwd", _p(dx), _p(x), ld, _p(S), _p(g_fm), _p(g_lin), plan.m, k, plan.N, _p(plan.sorted_pos),
        _p(plan.seg_start), _p(plan.n_unique), _p(out_rows), _p(out_bias), _p(out_lin), _p(ws), wsn, _stream(),
    )
    return out_rows, out_bias, out_lin


def segment_reduce_p2p_update(G_ptrs, rows_per_rank, KP, k, plan: SegmentPlan, table, bias_table, lin_table, opt, lr,
                              l2=0.0, gscal=None, m=1):
    plan.wait()
    ws, wsn = _reduce_ws(plan, k)
    _C.call(
        "rm_segment_reduce_p2p_update", _ptr_array(G_ptrs), _p(gscal), int(m), len(G_ptrs), int(rows_per_rank), KP, k,
        plan.N,
        _p(plan.sorted_pos), _p(plan.seg_start), _p(plan.uniq_rows), _p(plan.n_unique), _p(table), _p(bias_table),
        _p(lin_table), opt, float(lr), float(l2), _p(ws), wsn, _stream(),
    )


def emb_fm_bwd_update(dx, x, ld, S, g_fm, g_lin, plan: SegmentPlan, k, table, bias_table, lin_table, opt, lr, l2=0.0):
    """rm_emb_fm_bwd + rm_sparse_opt_step in one pass: the tables are updated in place, nothing is returned."""
    plan.wait()
    ws, wsn = _reduce_ws(plan, k)
    _C.call(
        "rm_emb_fm_bwd_update", _p(dx), _p(x), ld, _p(S), _p(g_fm), _p(g_lin), plan.m, k, plan.N, _p(plan.sorted_pos),
        _p(plan.seg_start), _p(plan.uniq_rows), _p(plan.n_unique), _p(table), _p(bias_table), _p(lin_table), opt,
        float(lr), float(l2), _p(ws), wsn, _stream(),
    )


@dataclass
class SparseGrad:
    """K2's output for one table: rows ``uniq_rows[:n]`` received ``rows[:n]``."""

    uniq_rows: torch.Tensor
    rows: torch.Tensor  # [N, k] or [N] for k == 1 tables
    n_unique: torch.Tensor

    def to_dense(self, total_rows: int) -> torch.Tensor:
        n = int(self.n_unique.item())
        rows = self.rows[:n].reshape(n, -1)
        out = torch.zeros(total_rows, rows.shape[1], dtype=rows.dtype, device=rows.device)
        out.index_copy_(0, self.uniq_rows[:n], rows)
        return out


# --------------------------------------------------------------------------- #
# K4 cross network
# --------------------------------------------------------------------------- #
def cross_fwd(x, w, b, w_out, w0_out):
    _dev_check(x)
    x, ld = _rows_view(x, "x")
    B, d = x.shape
    L = w.shape[0]
    logit = torch.empty(B, dtype=torch.float32, device=x.device)
    dots = torch.empty(B, max(L, 1), dtype=torch.float32, device=x.device)
    _C.call(
        "rm_cross_fwd", _p(x), ld, _p(w.contiguous()), _p(b.contiguous()), _p(w_out.contiguous()), _p(w0_out), B, d, L,
        _p(logit), _p(dots), _stream(),
    )
    return logit, dots


def cross_bwd(x, w, b, w_out, dots, gout, dx=None, accumulate=False):
    _dev_check(x)
    x, ld = _rows_view(x, "x")
    B, d = x.shape
    L = w.shape[0]
    dev = x.device
    if dx is None:
        dx = torch.empty(B, d, dtype=torch.float32, device=dev)
        accumulate = False
    assert dx.stride(1) == 1
    dw = torch.empty(L, d, dtype=torch.float32, device=dev)
    db = torch.empty(L, d, dtype=torch.float32, device=dev)
    dw_out = torch.empty(d, dtype=torch.float32, device=dev)
    dw0 = torch.empty(1, dtype=torch.float32, device=dev)
    ws_bytes = _C.lib.rm_cross_bwd_workspace_bytes(B, d, L)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    gout = gout.reshape(B).contiguous()
    _C.call(
        "rm_cross_bwd", _p(x), ld, _p(w.contiguous()), _p(b.contiguous()), _p(w_out.contiguous()), _p(dots), _p(gout),
        B, d, L, _p(dx), dx.stride(0), int(accumulate), _p(dw), _p(db), _p(dw_out), _p(dw0), _p(ws), ws_bytes,
        _stream(),
    )
    return dx, dw, db, dw_out, dw0


# --------------------------------------------------------------------------- #
# K5 CIN layer
# --------------------------------------------------------------------------- #
def cin_layer_fwd(x0, xk, W, bias, act: int, precision: int, want_pre=True):
    """x0 [B,m,D], xk [B,H,D] (rows dense, batch stride free), W [m*H, N] -> out [B,N,D], pre."""
    _dev_check(x0)
    _f32c(x0, "x0")
    B, m, D = x0.shape
    if x0.stride(2) != 1 or x0.stride(1) != D:
        x0 = x0.contiguous()
    H = xk.shape[1]
    assert xk.stride(2) == 1 and xk.stride(1) == D
    W = _f32c(W, "W").contiguous()
    N = W.shape[1]
    assert W.shape[0] == m * H
    dev = x0.device
    out = torch.empty(B, N, D, dtype=torch.float32, device=dev)
    pre = torch.empty(B, N, D, dtype=torch.float32, device=dev) if want_pre else None
    ws_bytes = _C.lib.rm_cin_layer_workspace_bytes(B, m, H, D, N, precision)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    global _last_cin_ws
    _last_cin_ws = ws if ws_bytes else None
    _C.call(
        "rm_cin_layer_fwd", _p(x0), x0.stride(0), _p(xk), xk.stride(0), _p(W), _p(bias.contiguous()), B, m, H, D, N, act, precision,
        _p(out), _p(pre), _p(ws), ws_bytes, _stream(),
    )
    return out, pre


_last_cin_ws = None


def cin_pool_fwd(out, n0: int):
    """pooled[b, n - n0] = sum_d out[b, n, d] for n >= n0 (layers.py:738-751: the direct half, sum-pooled over D)."""
    _dev_check(out)
    B, N, D = out.shape
    assert out.is_contiguous() and 0 <= n0 < N
    pooled = torch.empty(B, N - n0, dtype=torch.float32, device=out.device)
    _C.call("rm_cin_pool_fwd", _p(out), B, N, D, n0, _p(pooled), _stream())
    return pooled


def cin_pool_bwd(d_next, d_pool, B: int, N: int, D: int, n0: int, device):
    """dout [B,N,D] of a layer from the gradients of its two consumers (either may be None = zero)."""
    if d_next is not None:
        assert d_next.shape == (B, n0, D) and d_next.stride(2) == 1 and d_next.stride(1) == D
    if d_pool is not None:
        d_pool = d_pool.contiguous()
        assert d_pool.shape == (B, N - n0)
    dout = torch.empty(B, N, D, dtype=torch.float32, device=device)
    _C.call("rm_cin_pool_bwd", _p(d_next), 0 if d_next is None else d_next.stride(0), _p(d_pool), B, N, D, n0, _p(dout),
            _stream())
    return dout


def cin_tc_status() -> int:
    """Status word of the last tensor-core CIN launch (0 = ok, 2 = a bounded pipeline wait expired). Host sync."""
    if _last_cin_ws is None:
        return 0
    return int(_last_cin_ws[:4].view(torch.int32).item())


def cin_layer_bwd(x0, xk, W, pre, dout, act: int, precision: int, dx0, dxk):
    """dx0 [B,m,D] is accumulated into; dxk [B,H,D] (batch stride free) is written."""
    _dev_check(x0)
    B, m, D = x0.shape
    if x0.stride(2) != 1 or x0.stride(1) != D:
        x0 = x0.contiguous()
    H = xk.shape[1]
    W = W.contiguous()
    N = W.shape[1]
    dev = x0.device
    dout = dout.contiguous()
    assert xk.stride(2) == 1 and xk.stride(1) == D and dxk.stride(2) == 1 and dxk.stride(1) == D
    assert dx0.is_contiguous()
    dW = torch.empty(m * H, N, dtype=torch.float32, device=dev)
    dbias = torch.empty(N, dtype=torch.float32, device=dev)
    ws_bytes = _C.lib.rm_cin_layer_bwd_workspace_bytes(B, m, H, D, N, precision)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    if precision != 0 and D % 4 == 0 and m <= 32 and N <= 256:  # the tensor-core backward ran (cin.cu)
        global _last_cin_ws
        _last_cin_ws = ws
    _C.call(
        "rm_cin_layer_bwd", _p(x0), x0.stride(0), _p(xk), xk.stride(0), _p(W), _p(pre), _p(dout), B, m, H, D, N, act, precision,
        _p(dW), _p(dbias), _p(dx0), _p(dxk), dxk.stride(0), _p(ws), ws_bytes, _stream(),
    )
    return dW, dbias


# --------------------------------------------------------------------------- #
# N1 optimizer steps
# --------------------------------------------------------------------------- #
def sparse_opt_step(table, sg: SparseGrad, opt: int, lr: float, l2: float = 0.0):
    _dev_check(table)
    k = table.shape[1] if table.dim() == 2 else 1
    if not table.is_contiguous():  # a column of the interleaved [rows, 2] k=1 storage (tower layout)
        assert k == 1 and table.stride(0) > 1
        _C.call(
            "rm_sparse_opt_step_strided", _p(table), 1, table.stride(0), _p(sg.uniq_rows), _p(sg.rows), _p(sg.n_unique),
            sg.uniq_rows.numel(), opt, float(lr), float(l2), _stream(),
        )
        return
    _C.call(
        "rm_sparse_opt_step", _p(table), k, _p(sg.uniq_rows), _p(sg.rows), _p(sg.n_unique), sg.uniq_rows.numel(), opt,
        float(lr), float(l2), _stream(),
    )


def dense_opt_step(p, g, opt: int, lr: float, l2: float = 0.0):
    _dev_check(p)
    assert p.numel() == g.numel()
    g = g.contiguous()
    if not p.is_contiguous():  # strided view (tower layout): update a packed copy, write it back
        tmp = p.contiguous()
        _C.call("rm_dense_opt_step", _p(tmp), _p(g), tmp.numel(), opt, float(lr), float(l2), _stream())
        p.copy_(tmp)
        return
    _C.call("rm_dense_opt_step", _p(p), _p(g), p.numel(), opt, float(lr), float(l2), _stream())


# --------------------------------------------------------------------------- #
# A8' narrow dense layer backward (N <= 64)
# --------------------------------------------------------------------------- #
def narrow_linear_ok(N: int) -> bool:
    return 0 < N <= 64 and N % 4 == 0


def linear_bwd_input(g, W, d_ld: Optional[int] = None, out=None):
    """dx [B, d_ld] = g [B, N] @ W[d, N]^T, columns d..d_ld zero (g, W contiguous fp32, N <= 64, N % 4 == 0)."""
    _dev_check(g)
    B, N = g.shape
    d = W.shape[0]
    assert W.shape[1] == N and g.is_contiguous() and W.is_contiguous()
    d_ld = (d + 3) // 4 * 4 if d_ld is None else int(d_ld)
    if out is None:
        out = torch.empty(B, d_ld, dtype=torch.float32, device=g.device)
    assert out.shape == (B, d_ld) and out.is_contiguous()
    _C.call("rm_linear_bwd_input", _p(g), B, N, _p(W), d, _p(out), d_ld, _stream())
    return out


def linear_bwd_input_fm(g, W, m: int, k: int, x, S, g_fm, out=None):
    """G [B, m*k] = g @ W[:m*k]^T + g_fm * (S - x): the complete embedding-row gradients of DeepFM (MLP input gradient +
    FM backward) in one pass; ``x`` is the row buffer [B, ld], ``S`` [B, k] the field sums, ``g_fm`` [B]."""
    _dev_check(g)
    B, N = g.shape
    assert W.shape[0] >= m * k and W.shape[1] == N and g.is_contiguous() and W.is_contiguous()
    assert x.is_contiguous() and S.is_contiguous() and g_fm.is_contiguous() and x.shape[0] == B
    if out is None:
        out = torch.empty(B, m * k, dtype=torch.float32, device=g.device)
    assert out.is_contiguous() and out.numel() == B * m * k
    _C.call("rm_linear_bwd_input_fm", _p(g), B, N, _p(W), m, k, _p(x), x.shape[1], _p(S), _p(g_fm), _p(out), _stream())
    return out


def linear_bwd_weight(x, ld: int, K: int, g):
    """dW [K, N] = x[:, :K]^T @ g, x rows ld floats apart (ld % 4 == 0), deterministic slabbed batch reduction."""
    _dev_check(g)
    B, N = g.shape
    assert g.is_contiguous() and x.dtype == torch.float32
    ws_bytes = _C.lib.rm_linear_bwd_weight_workspace_bytes(B, K, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
    dW = torch.empty(K, N, dtype=torch.float32, device=g.device)
    _C.call("rm_linear_bwd_weight", _p(x), int(ld), _p(g), B, K, N, _p(dW), _p(ws), ws_bytes, _stream())
    return dW


# --------------------------------------------------------------------------- #
# (e) a2a staging for row-sharded tables
# --------------------------------------------------------------------------- #
def dense_opt_step_multi(pairs, opt: int, lr: float, l2: float = 0.0):
    """One launch for a list of (parameter, gradient) tensor pairs (contiguous fp32, same device)."""
    import ctypes

    if not pairs:
        return
    n = len(pairs)
    ps = (ctypes.c_void_p * n)(*[p.data_ptr() for p, _ in pairs])
    gs = (ctypes.c_void_p * n)(*[g.data_ptr() for _, g in pairs])
    ns = (ctypes.c_int64 * n)(*[p.numel() for p, _ in pairs])
    for p, g in pairs:
        _dev_check(p)
        assert p.is_contiguous() and g.is_contiguous() and p.dtype == torch.float32 and g.dtype == torch.float32
        assert g.numel() == p.numel()
    _C.call("rm_dense_opt_step_multi", ps, gs, ns, n, opt, float(lr), float(l2), _stream())


def unpack_rows(recv, pos, m, k, x, bias_out=None, lin_out=None):
    _dev_check(recv)
    n, KP = recv.shape
    assert recv.is_contiguous() and pos.dtype == torch.int32 and x.stride(1) == 1
    _C.call("rm_unpack_rows", _p(recv), n, KP, _p(pos), m, k, _p(x), x.stride(0), _p(bias_out), _p(lin_out), _stream())


def pack_grad_rows(dx, x, ld, S, g_fm, g_lin, pos, m, k, KP, out=None, n=None):
    """Gradient rows [n, KP] = [dx + g_fm*(S - x) | g_fm | g_lin | 0 | 0] (KP = k: no scalar columns); ``pos`` None
    keeps position order."""
    if n is None:
        n = pos.numel()
    send = out if out is not None else torch.empty(n, KP, dtype=torch.float32, device=x.device)
    assert send.is_contiguous() and send.shape == (n, KP)
    _C.call("rm_pack_grad_rows", _p(dx), _p(x), ld, _p(S), _p(g_fm), _p(g_lin), n, KP, _p(pos), m, k, _p(send),
            _stream())
    return send


# --------------------------------------------------------------------------- #
# (e') row-sharded tables over NVLink peer memory
# --------------------------------------------------------------------------- #
def _ptr_array(ptrs):
    """Host array of device pointers (``const float* const*`` in the C ABI)."""
    import ctypes

    return (ctypes.c_void_p * len(ptrs))(*[int(q) for q in ptrs])


def gather_fm_fwd_p2p(tab_ptrs, bias_ptrs, lin_ptrs, k, feat_sizes, local_offs, ids, dense, lin_dense, status=None):
    """Fused front end with every row read from its owner's table (peer memory).  Returns (x, fm, lin, S)."""
    _dev_check(ids)
    B, m = ids.shape
    W = len(tab_ptrs)
    n_dense = 0 if dense is None else dense.shape[1]
    d = m * k + n_dense
    ld = (d + 3) // 4 * 4
    dev = ids.device
    x = torch.empty(B, ld, dtype=torch.float32, device=dev)
    if ld > d:
        x[:, d:].zero_()
    fm = torch.empty(B, dtype=torch.float32, device=dev)
    lin = torch.empty(B, dtype=torch.float32, device=dev)
    S = torch.empty(B, k, dtype=torch.float32, device=dev)
    if dense is not None:
        dense = _f32c(dense, "dense").contiguous()
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    _C.call(
        "rm_gather_fm_fwd_p2p", _ptr_array(tab_ptrs), None if bias_ptrs is None else _ptr_array(bias_ptrs),
        None if lin_ptrs is None else _ptr_array(lin_ptrs), W, _p(feat_sizes), _p(local_offs), _p(ids), _p(dense),
        _p(lin_dense), n_dense, B, m, k, _p(x), ld, _p(fm), _p(lin), _p(S), _p(status), _stream(),
    )
    return x, fm, lin, S


def shard_plan(gids, W, rank, feat_sizes, local_offs, total_local, n_cap, status=None, side: bool = False):
    """Owner-side K2 plan over the ids of all ranks (gids [W*b, m]) -> SegmentPlan of global positions + n_own."""
    _dev_check(gids)
    assert gids.dtype == torch.int64 and gids.is_contiguous() and gids.dim() == 2
    Ntot, m = gids.numel(), gids.shape[1]
    n_cap = int(min(n_cap, Ntot))
    dev = gids.device
    ws_bytes = _C.lib.rm_shard_plan_workspace_bytes(Ntot, n_cap)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    sorted_gpos = torch.empty(n_cap, dtype=torch.int32, device=dev)
    seg_start = torch.empty(n_cap + 1, dtype=torch.int32, device=dev)
    uniq_rows = torch.empty(n_cap, dtype=torch.int64, device=dev)
    n_unique = torch.empty(1, dtype=torch.int32, device=dev)
    n_own = torch.empty(1, dtype=torch.int32, device=dev)
    ev = _launch_maybe_side(side, lambda: _C.call(
        "rm_shard_plan", _p(gids), Ntot, m, W, rank, _p(feat_sizes), _p(local_offs), int(total_local), n_cap, _p(ws),
        ws_bytes, _p(sorted_gpos), _p(seg_start), _p(uniq_rows), _p(n_unique), _p(n_own), _p(status), _stream(),
    ), (gids, feat_sizes, local_offs, ws, sorted_gpos, seg_start, uniq_rows, n_unique, n_own, status))
    plan = SegmentPlan(n_cap, 1, sorted_gpos, seg_start, uniq_rows, n_unique, ws, ev)
    plan.n_own = n_own
    return plan


def segment_reduce_p2p(G_ptrs, rows_per_rank, KP, k, plan: SegmentPlan, want_bias=True, want_lin=True, gscal=None,
                       m=1):
    dev = plan.sorted_pos.device
    n = max(plan.N, 1)
    out_rows = torch.empty(n, k, dtype=torch.float32, device=dev)
    out_bias = torch.empty(n, dtype=torch.float32, device=dev) if want_bias else None
    out_lin = torch.empty(n, dtype=torch.float32, device=dev) if want_lin else None
    plan.wait()
    ws, wsn = _reduce_ws(plan, k)
    _C.call(
        "rm_segment_reduce_p2p", _ptr_array(G_ptrs), _p(gscal), int(m), len(G_ptrs), int(rows_per_rank), KP, k, plan.N,
        _p(plan.sorted_pos), _p(plan.seg_start), _p(plan.n_unique), _p(out_rows), _p(out_bias), _p(out_lin), _p(ws),
        wsn, _stream(),
    )
    return out_rows, out_bias, out_lin


# --------------------------------------------------------------------------- #
# T  fused DeepFM tower (front end + first DNN layer on tcgen05; sorted fused backward + update)
# --------------------------------------------------------------------------- #
TOWER_UNIT = 2048  # sorted positions per work unit of the fused backward (16 tiles)


def tower_supported(m: int, k: int, n_dense: int, N1: int) -> bool:
    return bool(_C.lib.rm_tower_supported(int(m), int(k), int(n_dense), int(N1)))


def tower_bwd_supported(k: int, N1: int) -> bool:
    return k == 64 and N1 == 32


def tower_fwd(table, scal, table_offsets, ids, dense, lin_dense, W1, b1, want_x=False, status=None):
    """Fused front end + first DNN layer.  ``scal`` [rows, 2] = (bias, first-order weight) interleaved (or None),
    ``lin_dense`` a 1-D (possibly strided) view of the dense first-order weights (or None), ``W1`` [m*k+n_dense, N1].
    Returns (y1 [B,N1] pre-activation, fm [B], lin [B], S [B,k], x [B,ld] | None)."""
    _dev_check(table)
    _f32c(table, "table")
    assert ids.dtype == torch.int64 and ids.is_contiguous() and table.is_contiguous()
    B, m = ids.shape
    k = table.shape[1]
    n_dense = 0 if dense is None else dense.shape[1]
    d = m * k + n_dense
    N1 = W1.shape[1]
    assert W1.shape[0] == d and W1.is_contiguous() and b1.is_contiguous() and b1.numel() == N1
    dev = table.device
    x = None
    ld = (d + 3) // 4 * 4
    if want_x:
        x = torch.empty(B, ld, dtype=torch.float32, device=dev)
        if ld > d:
            x[:, d:].zero_()
    y1 = torch.empty(B, N1, dtype=torch.float32, device=dev)
    fm = torch.empty(B, dtype=torch.float32, device=dev)
    lin = torch.empty(B, dtype=torch.float32, device=dev)
    S = torch.empty(B, k, dtype=torch.float32, device=dev)
    if dense is not None:
        dense = _f32c(dense, "dense").contiguous()
    if scal is not None:
        assert scal.dim() == 2 and scal.shape[1] == 2 and scal.is_contiguous()
    ld_stride = 1
    if lin_dense is not None:
        assert lin_dense.dim() == 1 and lin_dense.numel() == n_dense
        ld_stride = lin_dense.stride(0) if n_dense > 1 else 1
    ws_bytes = _C.lib.rm_tower_fwd_workspace_bytes(m, k, N1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _C.call(
        "rm_tower_fwd", _p(table), _p(scal), _p(table_offsets), _p(ids), _p(dense), _p(lin_dense), ld_stride, n_dense,
        _p(W1), _p(b1), N1, B, m, k, _p(x), ld, _p(y1), _p(fm), _p(lin), _p(S), _p(status), _p(ws), ws_bytes, _stream(),
    )
    return y1, fm, lin, S, x


@dataclass
class TowerPlan:
    """Sorted (table row, position) pairs of one batch + the work units of the fused backward (ids only)."""

    B: int
    m: int
    unit: int
    sorted_keys: torch.Tensor  # uint32 as int32 storage [N]
    sorted_pos: torch.Tensor  # int32 [N]
    field_bounds: torch.Tensor  # int32 [m+1]
    unit_bounds: torch.Tensor  # int32 [m*(upf+1)]
    workspace: torch.Tensor
    ready: Optional["torch.cuda.Event"] = None

    def wait(self) -> None:
        if self.ready is not None:
            torch.cuda.current_stream().wait_event(self.ready)
            self.ready = None


def tower_plan(ids, table_offsets, total_rows, unit: int = TOWER_UNIT, status=None, side: bool = False) -> TowerPlan:
    _dev_check(ids)
    assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.dim() == 2
    B, m = ids.shape
    N = B * m
    dev = ids.device
    upf = _C.lib.rm_tower_units_per_field(B, unit)
    ws_bytes = _C.lib.rm_tower_plan_workspace_bytes(N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    keys = torch.empty(N, dtype=torch.int32, device=dev)
    pos = torch.empty(N, dtype=torch.int32, device=dev)
    fb = torch.empty(m + 1, dtype=torch.int32, device=dev)
    ub = torch.empty(m * (upf + 1) + 1, dtype=torch.int32, device=dev)  # + the hot-row flag
    ev = _launch_maybe_side(side, lambda: _C.call(
        "rm_tower_plan", _p(ids), _p(table_offsets), B, m, int(total_rows), unit, _p(ws), ws_bytes, _p(keys), _p(pos),
        _p(fb), _p(ub), _p(status), _stream(),
    ), (ids, table_offsets, ws, keys, pos, fb, ub, status))
    return TowerPlan(B, m, unit, keys, pos, fb, ub, ws, ev)


def tower_bwd_update(table, scal, plan: TowerPlan, g1, S, g_fm, g_lin, W1, opt: int, lr: float, l2: float = 0.0,
                     update: bool = True, debug: bool = False, status=None, out=None, variant: int = 0):
    """Fused sparse backward + optimizer update (in place on ``table`` / ``scal``).  ``variant``: 0 = both kernel variants
    launched, the plan's hot-row flag picks one on the device; 1 = plain only; 2 = hot-row variant only (same results).
    Returns dW1[:m*k] [m*k, N1]
    (and, with ``debug``, the summed gradient rows / k=1 gradients at the sorted position closing each segment)."""
    _dev_check(g1)
    B, m = plan.B, plan.m
    k = S.shape[1]
    N1 = g1.shape[1]
    assert g1.is_contiguous() and S.is_contiguous() and g_fm.is_contiguous() and W1.is_contiguous()
    # single GPU: one row per sample of the batch; row-sharded: the all-gathered operands of all W*b samples
    n_s = g1.shape[0]
    assert (n_s == B or hasattr(plan, "n_own")) and S.shape[0] == n_s and g_fm.numel() == n_s
    assert W1.shape[0] >= m * k and W1.shape[1] == N1
    assert g_lin is None or (g_lin.is_contiguous() and g_lin.numel() == n_s)
    dev = g1.device
    if out is not None:  # write dW1[:m*k] straight into the caller's (contiguous) buffer
        assert out.shape == (m * k, N1) and out.is_contiguous() and out.dtype == torch.float32
        dW1 = out
    else:
        dW1 = torch.empty(m * k, N1, dtype=torch.float32, device=dev)
    out_rows = out_scal = None
    if debug:
        out_rows = torch.zeros(plan.sorted_pos.numel(), k, dtype=torch.float32, device=dev)
        out_scal = torch.zeros(plan.sorted_pos.numel(), 2, dtype=torch.float32, device=dev)
    ws_bytes = _C.lib.rm_tower_bwd_workspace_bytes(B, m, plan.unit)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    plan.wait()
    _C.call(
        "rm_tower_bwd_update", _p(table), _p(scal), _p(plan.sorted_keys),
        _p(plan.sorted_pos), _p(plan.unit_bounds), _p(g1), _p(S), _p(g_fm), _p(g_lin), _p(W1), B, m, k, N1, plan.unit,
        opt if update else -1, float(lr), float(l2), int(variant), _p(dW1), _p(out_rows), _p(out_scal), _p(status), _p(ws),
        ws_bytes,
        _stream(),
    )
    if debug:
        return dW1, out_rows, out_scal
    return dW1


def umma_probe(At, Bt, variant: int = 0):
    _dev_check(At)
    K = At.shape[0]
    assert At.shape == (K, 128) and Bt.shape == (K, 32) and At.is_contiguous() and Bt.is_contiguous()
    D = torch.empty(128, 32, dtype=torch.float32, device=At.device)
    st = new_status(At.device)
    _C.call("rm_umma_probe", _p(At), _p(Bt), K, variant, _p(D), _p(st), _stream())
    return D, st


# --------------------------------------------------------------------------- #
# H  fused DeepFM head (second DNN layer .. loss, forward + backward in one kernel)
# --------------------------------------------------------------------------- #
def deepfm_head_supported(N1: int, N2: int) -> bool:
    return bool(_C.lib.rm_deepfm_head_supported(int(N1), int(N2)))


def deepfm_head(y1, fm, lin, w0, W2, b2, w3, b3, labels, act: int, task: int, grad_scale: float = 1.0, dense=None):
    """labels None: forward only -> (logit, pred).  Otherwise -> dict with logit, pred, loss [1] and the gradients
    g1 [B,N1], g [B], dW2, db2, dw3, dscal [1] (= db3 = dw0), db1 - all scaled by ``grad_scale`` - and, with the
    samples' ``dense`` features [B, n_dense], dW1_dense [n_dense, N1] = dense^T g1 and dlin_dense [n_dense] = dense^T g."""
    _dev_check(y1)
    B, N1 = y1.shape
    N2 = W2.shape[1]
    dev = y1.device
    f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
    for t in (y1, fm, lin, W2, b2, w3, b3):
        assert t.is_contiguous() and t.dtype == torch.float32
    logit, pred = f(B), f(B)
    if labels is None:
        _C.call("rm_deepfm_head", _p(y1), _p(fm), _p(lin), _p(w0), _p(W2), _p(b2), _p(w3), _p(b3), None, None, 0, B, N1, N2,
                act, task, 1.0, _p(logit), _p(pred), None, None, None, None, None, None, None, None, None, None, None, 0,
                _stream())
        return logit, pred
    labels = labels.to(torch.float32).contiguous()
    nd = 0
    if dense is not None:
        dense = _f32c(dense, "dense").contiguous()
        assert dense.dim() == 2 and dense.shape[0] == B
        nd = dense.shape[1]
    out = dict(logit=logit, pred=pred, loss=f(1), g1=f(B, N1), g=f(B), dW2=f(N1, N2), db2=f(N2), dw3=f(N2), dscal=f(1),
               db1=f(N1))
    if nd:
        out["dW1_dense"], out["dlin_dense"] = f(nd, N1), f(nd)
    ws_bytes = _C.lib.rm_deepfm_head_workspace_bytes(B, nd)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _C.call("rm_deepfm_head", _p(y1), _p(fm), _p(lin), _p(w0), _p(W2), _p(b2), _p(w3), _p(b3), _p(labels),
            _p(dense) if nd else None, nd, B, N1, N2, act, task, float(grad_scale), _p(logit), _p(pred), _p(out["loss"]),
            _p(out["g1"]), _p(out["g"]), _p(out["dW2"]), _p(out["db2"]), _p(out["dw3"]), _p(out["dscal"]), _p(out["db1"]),
            _p(out.get("dW1_dense")), _p(out.get("dlin_dense")), _p(ws), ws_bytes, _stream())
    return out


# --------------------------------------------------------------------------- #
# T'  fused tower over row-sharded tables (NVLink peer memory)
# --------------------------------------------------------------------------- #
def tower_fwd_p2p(tab_ptrs, scal_ptrs, k, feat_sizes, local_offs, ids, dense, lin_dense, W1, b1, status=None):
    """rm_tower_fwd with every row read from its owner's shard.  Returns (y1, fm, lin, S)."""
    _dev_check(ids)
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    B, m = ids.shape
    W = len(tab_ptrs)
    n_dense = 0 if dense is None else dense.shape[1]
    N1 = W1.shape[1]
    assert W1.shape[0] == m * k + n_dense and W1.is_contiguous() and b1.is_contiguous()
    dev = ids.device
    y1 = torch.empty(B, N1, dtype=torch.float32, device=dev)
    fm = torch.empty(B, dtype=torch.float32, device=dev)
    lin = torch.empty(B, dtype=torch.float32, device=dev)
    S = torch.empty(B, k, dtype=torch.float32, device=dev)
    if dense is not None:
        dense = _f32c(dense, "dense").contiguous()
    ld_stride = 1
    if lin_dense is not None:
        assert lin_dense.dim() == 1 and lin_dense.numel() == n_dense
        ld_stride = lin_dense.stride(0) if n_dense > 1 else 1
    ws_bytes = _C.lib.rm_tower_fwd_workspace_bytes(m, k, N1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _C.call(
        "rm_tower_fwd_p2p", _ptr_array(tab_ptrs), None if scal_ptrs is None else _ptr_array(scal_ptrs), W,
        _p(feat_sizes), _p(local_offs), _p(ids), _p(dense), _p(lin_dense), ld_stride, n_dense, _p(W1), _p(b1), N1, B, m, k,
        _p(y1), _p(fm), _p(lin), _p(S), _p(status), _p(ws), ws_bytes, _stream(),
    )
    return y1, fm, lin, S


def tower_shard_plan(gids, W, rank, feat_sizes, local_offs, total_local, n_cap, b_cap, unit: int = TOWER_UNIT,
                     status=None, side: bool = False, fork: bool = True) -> TowerPlan:
    """Owner-side plan of the fused backward over the ids of all ranks (gids [W*b, m])."""
    _dev_check(gids)
    assert gids.dtype == torch.int32 and gids.is_contiguous() and gids.dim() == 2
    Ntot, m = gids.numel(), gids.shape[1]
    n_cap = int(min(n_cap, Ntot))
    dev = gids.device
    upf = _C.lib.rm_tower_units_per_field(int(b_cap), unit)
    ws_bytes = _C.lib.rm_tower_shard_plan_workspace_bytes(Ntot, n_cap)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    keys = torch.empty(n_cap, dtype=torch.int32, device=dev)
    pos = torch.empty(n_cap, dtype=torch.int32, device=dev)
    fb = torch.empty(m + 1, dtype=torch.int32, device=dev)
    ub = torch.empty(m * (upf + 1) + 1, dtype=torch.int32, device=dev)  # + the hot-row flag
    n_own = torch.empty(1, dtype=torch.int32, device=dev)
    ev = _launch_maybe_side(side, lambda: _C.call(
        "rm_tower_shard_plan", _p(gids), Ntot, m, W, rank, _p(feat_sizes), _p(local_offs), int(total_local), n_cap,
        int(b_cap), unit, _p(ws), ws_bytes, _p(keys), _p(pos), _p(fb), _p(ub), _p(n_own), _p(status), _stream(),
    ), (gids, feat_sizes, local_offs, ws, keys, pos, fb, ub, n_own, status), fork)
    plan = TowerPlan(int(b_cap), m, unit, keys, pos, fb, ub, ws, ev)
    plan.n_own = n_own
    return plan
