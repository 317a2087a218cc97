"""Row-sharded embedding tables + data-parallel dense layers over one NVSwitch box.

Nothing in the single-process reference corresponds to this module.  Layout: global row ``r`` of table ``f``
lives on rank ``r mod W`` at local row ``r div W`` (cyclic: immune to id-range skew); every rank keeps
``ceil(V_f / W)`` rows per table in one concatenated local table.  Per step and rank:

    forward   a2a #1  ids -> owners (local row keys, destination-sorted)
              owners: K1 gather straight into the return buffer [n_recv, k+4] (row | bias | lin | pad)
              a2a #2  vectors back;  rm_unpack_rows -> DNN row buffer; FM on the rows (K3)
    backward  rm_pack_grad_rows (d(x) + FM backward, destination-sorted) ; a2a #3 gradient rows -> owners
              owners: K2 (rm_segment_plan / rm_segment_reduce) directly on the receive buffer
    dense     one flat-bucket all-reduce of the replicated parameters' gradients (they are tiny)

The routing arithmetic (``ShardPlan``, ``build_exchange``) is pure torch and device agnostic so that it is
tested on CPU with the gloo backend (tests/test_dist_gloo.py); the row movement is CUDA kernels + NCCL.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch.autograd import Function

__all__ = ["ShardPlan", "Exchange", "build_exchange", "shard_model", "ShardedFrontEndFunction", "allreduce_dense"]


class ShardPlan:
    """Cyclic row sharding of m tables over ``world`` ranks."""

    def __init__(self, feat_sizes: Sequence[int], world: int, rank: int, group=None):
        self.feat_sizes = [int(v) for v in feat_sizes]
        self.world, self.rank, self.group = int(world), int(rank), group
        self.local_sizes = [(v + world - 1) // world for v in self.feat_sizes]
        offs = [0]
        for s in self.local_sizes:
            offs.append(offs[-1] + s)
        self.local_offsets = offs
        self.total_local = offs[-1]
        self._offs_dev = {}

    def offsets_on(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._offs_dev:
            self._offs_dev[key] = torch.tensor(self.local_offsets[:-1], dtype=torch.int64, device=device)
        return self._offs_dev[key]

    def route(self, ids: torch.Tensor):
        """ids [b, m] -> (dest [b*m] int64 in [0, W), key [b*m] int64 = owner-local concatenated row)."""
        W = self.world
        dest = torch.remainder(ids, W)
        key = torch.div(ids, W, rounding_mode="floor") + self.offsets_on(ids.device)[None, :]
        return dest.reshape(-1), key.reshape(-1)

    def local_rows_of(self, f: int) -> torch.Tensor:
        """Global row ids of table f held by this rank, in local order (for building shards from a full table)."""
        if self.rank >= self.feat_sizes[f]:
            return torch.zeros(0, dtype=torch.int64)  # table with fewer rows than ranks
        return torch.arange(self.rank, self.feat_sizes[f], self.world)


@dataclass
class Exchange:
    sorted_pos: torch.Tensor  # int32 [n]: routed row j came from position p = b*m + f
    send_splits: List[int]
    recv_splits: List[int]
    recv_keys: torch.Tensor  # int64 [n_recv] owner-local rows requested from this rank


def _a2a(out, inp, out_splits, in_splits, group):
    dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)


def build_exchange(plan: ShardPlan, ids: torch.Tensor, sort_fn: Callable[[torch.Tensor], torch.Tensor]) -> Exchange:
    """Route one batch: destination-sort the positions, exchange counts then keys (a2a #1)."""
    W = plan.world
    dest, key = plan.route(ids)
    sorted_pos = sort_fn(dest)  # stable: ascending position inside every destination bucket
    send_counts = torch.bincount(dest, minlength=W)
    recv_counts = torch.empty_like(send_counts)
    _a2a(recv_counts, send_counts, None, None, plan.group)
    both = torch.stack([send_counts, recv_counts]).cpu()  # the step's one host sync (split sizes)
    send_splits, recv_splits = both[0].tolist(), both[1].tolist()
    send_keys = key[sorted_pos.long()]
    recv_keys = torch.empty(sum(recv_splits), dtype=torch.int64, device=ids.device)
    _a2a(recv_keys, send_keys, recv_splits, send_splits, plan.group)
    return Exchange(sorted_pos, send_splits, recv_splits, recv_keys)


def _cuda_sort_fn(world: int):
    def fn(dest: torch.Tensor) -> torch.Tensor:
        from .. import ops

        # K2's plan kernel as a stable bucket sort: keys in [0, W) -> one 3-bit radix pass at W = 8
        return ops.segment_plan(dest.contiguous(), None, max(world, 2)).sorted_pos[: dest.numel()]

    return fn


class ShardedFrontEndFunction(Function):
    """Sharded version of autograd.FrontEndFunction: same outputs (xbuf, fm, lin), rows fetched over NVLink."""

    @staticmethod
    def forward(ctx, table, bias_table, W_lin, lin_table, lin_dense, plan: ShardPlan, status, ids, dense):
        from .. import ops

        b, m = ids.shape
        k = table.shape[1]
        KP = k + 4
        dev = table.device
        ex = build_exchange(plan, ids, _cuda_sort_fn(plan.world))
        n_recv = ex.recv_keys.numel()
        rows = torch.zeros(n_recv, KP, dtype=torch.float32, device=dev)
        flat_offs = torch.tensor([0, plan.total_local], dtype=torch.int64, device=dev)
        keys2d = ex.recv_keys.view(-1, 1)
        if n_recv:
            ops.gather(table, flat_offs, keys2d, out=rows, status=status)
            if bias_table is not None:
                ops.gather(bias_table.reshape(-1, 1), flat_offs, keys2d, out=rows[:, k:], status=status)
            if lin_table is not None:
                ops.gather(lin_table.reshape(-1, 1), flat_offs, keys2d, out=rows[:, k + 1 :], status=status)
        back = torch.empty(b * m, KP, dtype=torch.float32, device=dev)
        _a2a(back, rows, ex.send_splits, ex.recv_splits, plan.group)  # a2a #2
        n_dense = 0 if dense is None else dense.shape[1]
        d = m * k + n_dense
        ld = (d + 3) // 4 * 4
        x = torch.empty(b, ld, dtype=torch.float32, device=dev)
        if ld > m * k:
            x[:, m * k :].zero_()
        bias_pos = torch.empty(b, m, dtype=torch.float32, device=dev)
        lin_pos = torch.empty(b, m, dtype=torch.float32, device=dev)
        ops.unpack_rows(back, ex.sorted_pos, m, k, x, bias_pos, lin_pos)
        if n_dense:
            x[:, m * k : d] = dense
        fm, S = ops.fm_fwd(x[:, : m * k].unflatten(1, (m, k)), bias_pos if bias_table is not None else None)
        lin = lin_pos.sum(dim=1)
        if lin_dense is not None and n_dense:
            lin = lin + dense @ lin_dense
        ctx.plan, ctx.ex = plan, ex
        ctx.table, ctx.bias_table, ctx.W_lin = table, bias_table, W_lin
        ctx.has_lin = lin_table is not None
        ctx.has_lin_dense = lin_dense is not None and n_dense > 0
        ctx.save_for_backward(x, S, dense)
        ctx.m, ctx.k, ctx.KP = m, k, KP
        ctx.set_materialize_grads(False)
        return x, fm.reshape(-1, 1), lin.reshape(-1, 1)

    @staticmethod
    def backward(ctx, dx, dfm, dlin):
        from .. import ops
        from ..autograd import attach_sparse_grad

        x, S, dense = ctx.saved_tensors
        plan, ex = ctx.plan, ctx.ex
        m, k, KP = ctx.m, ctx.k, ctx.KP
        ld = x.shape[1]
        if dx is not None and (dx.stride(1) != 1 or dx.stride(0) != ld):
            dx = dx.contiguous()
        g_fm = None if dfm is None else dfm.reshape(-1).contiguous()
        g_lin = None if dlin is None else dlin.reshape(-1).contiguous()
        send = ops.pack_grad_rows(dx, x, ld, S, g_fm, g_lin, ex.sorted_pos, m, k, KP)
        n_recv = ex.recv_keys.numel()
        recv = torch.empty(n_recv, KP, dtype=torch.float32, device=x.device)
        _a2a(recv, send, ex.recv_splits, ex.send_splits, plan.group)  # a2a #3
        if n_recv:
            sp = ops.segment_plan(ex.recv_keys, None, plan.total_local)
            rows = ops.segment_reduce(recv, sp, KP, ld=KP)
            attach_sparse_grad(ctx.table, ops.SparseGrad(sp.uniq_rows, rows[:, :k].contiguous(), sp.n_unique))
            if ctx.bias_table is not None and g_fm is not None:
                attach_sparse_grad(ctx.bias_table, ops.SparseGrad(sp.uniq_rows, rows[:, k].contiguous(), sp.n_unique))
            if ctx.has_lin and ctx.W_lin is not None and g_lin is not None:
                attach_sparse_grad(ctx.W_lin, ops.SparseGrad(sp.uniq_rows, rows[:, k + 1].contiguous(), sp.n_unique))
        if ctx.has_lin_dense and ctx.W_lin is not None and g_lin is not None:
            ctx.W_lin.rm_dense_tail = (plan.total_local, dense.t() @ g_lin)  # replicated: all-reduced by the optimizer
        return (None,) * 10


def allreduce_dense(grads: List[torch.Tensor], group=None) -> None:
    """One flat bucket: the replicated parameters are ~1.5 M floats at most, latency bound."""
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off : off + n].view_as(g))
        off += n


def shard_model(model, world: int, rank: int, group=None):
    """Switch a recman.th model to row-sharded tables + DP dense layers (before its variables are created)."""
    if model.variables:
        raise RuntimeError("shard_model must be called before the first forward creates the variables")
    sizes = [f.feat_size for f in model.feat_dict.embedding_feats]
    model.shard = ShardPlan(sizes, world, rank, group)
    return model
