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

Peer-memory mode (default for any world size <= 8, i.e. one NVSwitch box; powers of two use mask / shift for the
owner / local-row arithmetic, other sizes one integer division per id; ``PeerMemory``,
``P2PFrontEndFunction``): the tables and one gradient-row buffer per rank are cudaIpc-mapped into every rank, and

    forward   all-gather of the ids (13.6 MB per rank at C5; also the step's cross-rank barrier)
              ONE fused front-end kernel per rank that reads every row from its owner: local HBM or NVLink loads
    backward  rm_pack_grad_rows (position order, into this rank's buffer G) ; 4-byte all-reduce (= "all G written")
              owner-side plan over the gathered ids + deterministic segmented reduce that pulls its rows from the
              peers' G over NVLink, summed in ascending GLOBAL position (the single-GPU order of the global batch)

There is no all-to-all, no pack/unpack of the forward rows, no split-size host sync, so the step is CUDA-graph
capturable; NVLink traffic is exactly one row per remote id each way.
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch.autograd import Function

__all__ = ["ShardPlan", "Exchange", "build_exchange", "shard_model", "ShardedFrontEndFunction", "allreduce_dense",
           "PeerMemory", "P2PFrontEndFunction", "P2PTowerFunction"]


class _DevicePointer:
    """Raw device allocation exposed through __cuda_array_interface__ so that torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerMemory:
    """cudaIpc-shared device allocations of one process group: ``alloc`` is collective (every rank, same order)."""

    def __init__(self, world: int, rank: int, group=None):
        self.world, self.rank, self.group = world, rank, group
        self._ptrs = {}  # local base pointer -> [pointer of that allocation in rank 0..W-1's address space mapping]
        self._keep = []

    def alloc(self, shape, dtype=torch.float32, zero=True) -> torch.Tensor:
        import ctypes

        from .. import _C

        n = 1
        for v in shape:
            n *= int(v)
        nbytes = max(n * torch.empty((), dtype=dtype).element_size(), 256)
        # whole 2 MiB pages: an allocation with a ragged tail is imported by the peers with small pages, and random
        # NVLink reads over a multi-GB small-page mapping fall off a TLB cliff (scripts/p2p_bench3.py)
        gran = int(os.environ.get("RM_P2P_ALLOC_GRAN", str(2 << 20)))
        nbytes = (nbytes + gran - 1) // gran * gran
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        _C.check(_C.lib.rm_p2p_alloc(nbytes, ctypes.byref(ptr), handle), "rm_p2p_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(int(ptr.value))
                continue
            q = ctypes.c_void_p()
            hb = (ctypes.c_uint8 * 64).from_buffer_copy(h)
            _C.check(_C.lib.rm_p2p_open(hb, ctypes.byref(q)), "rm_p2p_open")
            ptrs.append(int(q.value))
        dev = torch.device("cuda", torch.cuda.current_device())
        raw = torch.as_tensor(_DevicePointer(ptr.value, nbytes), device=dev)
        t = raw[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)
        if zero:
            t.zero_()
        self._ptrs[int(ptr.value)] = ptrs
        self._keep.append(raw)
        return t

    def ptrs_of(self, t: torch.Tensor):
        """Pointers of ``t``'s allocation in every rank (``t`` must start at the allocation's base)."""
        return self._ptrs[int(t.data_ptr())]


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
        self.peer: Optional[PeerMemory] = None  # set by enable_peer_memory(): tables live in cudaIpc-shared memory
        self._bufs = {}
        self.slack = 1.0  # owner-side plan capacity = 2x the uniform share (skewed ids load the owners of hot rows), see capacity()

    # ---- peer-memory mode -------------------------------------------------------------------------------
    def enable_peer_memory(self):
        if self.world > 8:
            raise ValueError("peer-memory sharding needs a world size <= 8 (one NVSwitch box)")
        self.peer = PeerMemory(self.world, self.rank, self.group)
        return self

    def alloc(self, shape, zero=True) -> torch.Tensor:
        """Parameter storage: cudaIpc-shared in peer-memory mode, plain torch memory otherwise."""
        if self.peer is not None:
            return self.peer.alloc(shape, zero=zero)
        return (torch.zeros if zero else torch.empty)(*shape, dtype=torch.float32, device="cuda")

    def feat_sizes_on(self, device) -> torch.Tensor:
        key = ("fs", str(device))
        if key not in self._offs_dev:
            self._offs_dev[key] = torch.tensor(self.feat_sizes, dtype=torch.int64, device=device)
        return self._offs_dev[key]

    def grad_buffer(self, n: int, KP: int) -> torch.Tensor:
        key = ("G", n, KP)
        if key not in self._bufs:
            self._bufs[key] = self.peer.alloc((n, KP), zero=True)
        return self._bufs[key]

    def flag(self, device) -> torch.Tensor:
        key = ("flag", str(device))
        if key not in self._bufs:
            self._bufs[key] = torch.zeros(1, dtype=torch.float32, device=device)
        return self._bufs[key]

    def capacity(self, b: int) -> int:
        """Entries of the owner-side plan for a local batch of ``b`` samples: the largest share any rank owns under
        uniform ids (tables with fewer rows than ranks load the low ranks), plus ``slack``; exceeding it sets status
        bit 2 (``DeepModel.check_ids`` raises)."""
        W, m = self.world, len(self.feat_sizes)
        share = max(sum(((v - r + W - 1) // W) / v for v in self.feat_sizes if v > 0) for r in range(W))
        return min(W * b * m, int(W * b * share * (1.0 + self.slack)) + 1024)

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
    def forward(ctx, table, bias_table, W_lin, lin_table, lin_dense, plan: ShardPlan, status, ids, dense,
                fused_opt=None, fm_back=None):
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
        return (None,) * 11


class P2PFrontEndFunction(Function):
    """Peer-memory version of autograd.FrontEndFunction: same outputs (xbuf, fm, lin); rows come from their owners."""

    @staticmethod
    def forward(ctx, table, bias_table, W_lin, lin_table, lin_dense, plan: ShardPlan, status, ids, dense,
                fused_opt=None, fm_back=None):
        from .. import ops

        b, m = ids.shape
        k = table.shape[1]
        W, dev, peer = plan.world, table.device, plan.peer
        ctx.fused_opt, ctx.lin_table = fused_opt, lin_table
        ctx.fm_back = fm_back
        ids = ids.contiguous()
        gids = torch.empty(W * b, m, dtype=torch.int64, device=dev)
        # every rank has finished the previous step's table update once this returns (it is also the barrier that
        # protects the peers' tables and gradient buffers)
        dist.all_gather_into_tensor(gids, ids, group=plan.group)
        x, fm, lin, S = ops.gather_fm_fwd_p2p(
            peer.ptrs_of(table), None if bias_table is None else peer.ptrs_of(bias_table),
            None if lin_table is None else peer.ptrs_of(lin_table), k, plan.feat_sizes_on(dev), plan.offsets_on(dev),
            ids, dense, lin_dense, status=status)
        ctx.plan, ctx.status = plan, status
        # the owner-side K2 plan needs the gathered ids only: build it on the side stream, under the NVLink-bound
        # front-end kernel and the MLP forward
        ctx.sp = None
        if any(ctx.needs_input_grad):
            ctx.sp = ops.shard_plan(gids, W, plan.rank, plan.feat_sizes_on(dev), plan.offsets_on(dev),
                                    plan.total_local, plan.capacity(b), status, side=True)
        ctx.table, ctx.bias_table, ctx.W_lin = table, bias_table, W_lin
        ctx.has_lin = lin_table is not None
        ctx.has_lin_dense = lin_dense is not None and dense is not None and dense.shape[1] > 0
        if fm_back is not None:
            fm_back.x_ptr, fm_back.S, fm_back.m, fm_back.k = x.data_ptr(), S, m, k
            fm_back.g_fm = fm_back.G = None
            fm_back.alloc = lambda: plan.grad_buffer(b * m, k)
        ctx.save_for_backward(x, S, dense, gids)
        ctx.b, ctx.m, ctx.k = b, m, k
        ctx.set_materialize_grads(False)
        return x, fm.reshape(-1, 1), lin.reshape(-1, 1)

    @staticmethod
    def backward(ctx, dx, dfm, dlin):
        from .. import ops
        from ..autograd import attach_sparse_grad

        x, S, dense, gids = ctx.saved_tensors
        plan: ShardPlan = ctx.plan
        b, m, k = ctx.b, ctx.m, ctx.k
        KP = k + 4
        ld = x.shape[1]
        dev = x.device
        if dx is not None and (dx.stride(1) != 1 or dx.stride(0) != ld):
            dx = dx.contiguous()
        g_fm = None if dfm is None else dfm.reshape(-1).contiguous()
        g_lin = None if dlin is None else dlin.reshape(-1).contiguous()
        sp, ctx.sp = ctx.sp, None
        if sp is None:
            sp = ops.shard_plan(gids, plan.world, plan.rank, plan.feat_sizes_on(dev), plan.offsets_on(dev),
                                plan.total_local, plan.capacity(b), ctx.status)
        # gradient rows are exactly k floats (256-byte rows at k = 64: one NVLink request fewer per row than k+4); the two
        # k=1 gradients are per-sample values and travel once, in the all-gather that also says "every G is complete"
        fb = ctx.fm_back
        if fb is not None and fb.G is not None:
            G, fb.G = fb.G, None  # the first MLP layer already wrote dx + FM backward into this rank's peer-shared buffer
        else:
            G = plan.grad_buffer(b * m, k)
            ops.pack_grad_rows(dx, x, ld, S, g_fm, g_lin, None, m, k, k, out=G, n=b * m)
        zeros = None
        if g_fm is None or g_lin is None:
            zeros = torch.zeros(b, dtype=torch.float32, device=dev)
        gs = torch.stack([g_fm if g_fm is not None else zeros, g_lin if g_lin is not None else zeros], dim=1)
        gscal = torch.empty(plan.world * b, 2, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gscal, gs.contiguous(), group=plan.group)
        want_bias = ctx.bias_table is not None and g_fm is not None
        want_lin = ctx.has_lin and ctx.W_lin is not None and g_lin is not None
        if ctx.has_lin_dense and ctx.W_lin is not None and g_lin is not None:
            ctx.W_lin.rm_dense_tail = (plan.total_local, dense.t() @ g_lin)
        if ctx.fused_opt is not None:
            kind, lr = ctx.fused_opt
            ops.segment_reduce_p2p_update(plan.peer.ptrs_of(G), b * m, k, k, sp, ctx.table.data,
                                          ctx.bias_table.data if want_bias else None,
                                          ctx.lin_table if want_lin else None, kind, lr, gscal=gscal, m=m)
            return (None,) * 11
        rows, ob, ol = ops.segment_reduce_p2p(plan.peer.ptrs_of(G), b * m, k, k, sp, want_bias, want_lin, gscal=gscal, m=m)
        attach_sparse_grad(ctx.table, ops.SparseGrad(sp.uniq_rows, rows, sp.n_unique))
        if want_bias:
            attach_sparse_grad(ctx.bias_table, ops.SparseGrad(sp.uniq_rows, ob, sp.n_unique))
        if want_lin:
            attach_sparse_grad(ctx.W_lin, ops.SparseGrad(sp.uniq_rows, ol, sp.n_unique))
        return (None,) * 11


def _all_gather_many(pairs, group=None) -> None:
    """Several all_gather_into_tensor calls as ONE grouped NCCL launch (ncclGroupStart/End) where torch offers it."""
    cm = getattr(dist, "_coalescing_manager", None)
    if cm is None:
        for out, inp in pairs:
            dist.all_gather_into_tensor(out, inp, group=group)
        return
    with cm(group=group):
        for out, inp in pairs:
            dist.all_gather_into_tensor(out, inp, group=group)


class P2PTowerFunction(Function):
    """Row-sharded version of autograd.TowerFunction (fused DeepFM tower): ids -> (y1, fm, lin).

    forward   all-gather of the ids (the step's cross-rank barrier) ; ONE kernel per rank (rm_tower_fwd_p2p) reads every
              row from its owner - local HBM or NVLink - and runs FM + first-order + the first DNN layer on it
    backward  all-gathers of the per-sample operands (g1 [b,32], S [b,64], g_fm, g_lin: 392 B per sample instead of
              m gradient rows of 256 B) ; the owner forms the gradient rows of its own positions itself and applies the
              update in place (rm_tower_bwd_update over rm_tower_shard_plan) ; dW1 partials join the dense all-reduce.
    Training only with the in-kernel update (k = 64, first hidden layer 32)."""

    @staticmethod
    def forward(ctx, table, scal, bias_param, W_lin, W1, b1, plan: ShardPlan, status, ids, dense, fused_opt,
                grad_mode=True, side=None):
        from .. import ops

        b, m = ids.shape
        k = table.shape[1]
        W, dev, peer = plan.world, table.device, plan.peer
        total = plan.total_local
        n_dense = 0 if dense is None else dense.shape[1]
        ids = ids.contiguous()
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        if need_grad and (fused_opt is None or not ops.tower_bwd_supported(k, W1.shape[1])):
            raise NotImplementedError("row-sharded fused tower: training needs the in-kernel sparse update "
                                      "(fit_on_batch; no L2 on the tables) with k = 64 and a first hidden layer of 32")
        ctx.tp = None
        if need_grad:
            # Owner-side plan, entirely on the side stream and forked BEFORE the forward kernel is enqueued: the ids
            # all-gather (int32: half the bytes of the reference's int64 ids; a table has < 2^31 rows, an id outside
            # int32 is outside its table anyway and becomes -1 = "invalid") and the sort run under the forward's NVLink
            # reads.  The forward needs no barrier of its own: the peers' rows it reads were last written by the
            # previous step's update, and every rank's dense all-reduce of that step was enqueued behind its update -
            # this rank's all-reduce could not complete before every rank had entered it.
            ids32 = torch.where((ids >= 0) & (ids < 2 ** 31), ids, torch.full_like(ids, -1)).to(torch.int32)
            gids = torch.empty(W * b, m, dtype=torch.int32, device=dev)
            sst = ops.side_stream()
            sst.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(sst):
                dist.all_gather_into_tensor(gids, ids32, group=plan.group)
            ids32.record_stream(sst)
            gids.record_stream(sst)
            n_cap = plan.capacity(b)
            ctx.tp = ops.tower_shard_plan(gids, W, plan.rank, plan.feat_sizes_on(dev), plan.offsets_on(dev), total, n_cap,
                                          (n_cap + m - 1) // m, status=status, side=True, fork=False)
        lin_dense = scal[total:, 1] if n_dense else None
        y1, fm, lin, S = ops.tower_fwd_p2p(peer.ptrs_of(table), peer.ptrs_of(scal), k, plan.feat_sizes_on(dev),
                                           plan.offsets_on(dev), ids, dense, lin_dense, W1, b1, status=status)
        ctx.plan, ctx.status, ctx.fused_opt = plan, status, fused_opt
        ctx.table, ctx.scal, ctx.W_lin = table, scal, W_lin
        ctx.n_dense = n_dense
        ctx.side = side
        ctx.save_for_backward(S, dense, W1)
        ctx.b, ctx.m, ctx.k = b, m, k
        ctx.set_materialize_grads(False)
        return y1, fm.reshape(-1, 1), lin.reshape(-1, 1)

    @staticmethod
    def backward(ctx, dy1, dfm, dlin):
        from .. import ops

        S, dense, W1 = ctx.saved_tensors
        plan: ShardPlan = ctx.plan
        b, m, k = ctx.b, ctx.m, ctx.k
        W, dev = plan.world, S.device
        N1 = W1.shape[1]
        g1 = torch.zeros(b, N1, dtype=torch.float32, device=dev) if dy1 is None else dy1.contiguous()
        g_fm = torch.zeros(b, dtype=torch.float32, device=dev) if dfm is None else dfm.reshape(-1).contiguous()
        g_lin = torch.zeros(b, dtype=torch.float32, device=dev) if dlin is None else dlin.reshape(-1).contiguous()
        # per-sample operands of every rank: the owners form the gradient rows themselves
        g1_all = torch.empty(W * b, N1, dtype=torch.float32, device=dev)
        S_all = torch.empty(W * b, k, dtype=torch.float32, device=dev)
        gfm_all = torch.empty(W * b, dtype=torch.float32, device=dev)
        glin_all = torch.empty(W * b, dtype=torch.float32, device=dev)
        _all_gather_many([(S_all, S), (g1_all, g1), (gfm_all, g_fm), (glin_all, g_lin)], plan.group)
        total = plan.total_local
        # gradients that need no embedding rows: from the head kernel of the same step when it left them (TowerSide)
        side_db1, side_dW1d, side_dlind = ctx.side.take() if ctx.side is not None else (None, None, None)
        if dy1 is None or dlin is None or (ctx.n_dense and (side_dW1d is None or side_dlind is None)):
            side_db1 = None
        if ctx.n_dense:  # replicated: all-reduced by the optimizer
            ctx.W_lin.rm_dense_tail = (total, side_dlind if side_db1 is not None else dense.t() @ g_lin)
        kind, lr = ctx.fused_opt[:2]
        variant = ctx.fused_opt[2] if len(ctx.fused_opt) > 2 else 0
        tp, ctx.tp = ctx.tp, None
        ctx.table.rm_hot_flag = tp.unit_bounds[-1:]
        dW1 = torch.empty(W1.shape, dtype=torch.float32, device=dev)
        ops.tower_bwd_update(ctx.table.data, ctx.scal[:total], tp, g1_all, S_all, gfm_all, glin_all, W1.data, kind, lr,
                             status=ctx.status, out=dW1[: m * k], variant=variant)
        if ctx.n_dense:
            if side_db1 is not None:
                dW1[m * k :] = side_dW1d
            else:
                torch.mm(dense.t(), g1, out=dW1[m * k :])
        return (None, None, None, None, dW1, side_db1 if side_db1 is not None else g1.sum(0)) + (None,) * 7


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


def shard_model(model, world: int, rank: int, group=None, mode: str = "auto"):
    """Switch a recman.th model to row-sharded tables + DP dense layers (before its variables are created).

    ``mode``: "p2p" (NVLink peer memory, world <= 8), "a2a" (NCCL all-to-all exchange, any world) or "auto"."""
    if model.variables:
        raise RuntimeError("shard_model must be called before the first forward creates the variables")
    sizes = [f.feat_size for f in model.feat_dict.embedding_feats]
    model.shard = ShardPlan(sizes, world, rank, group)
    if mode == "auto":
        mode = "p2p" if world <= 8 and torch.cuda.is_available() else "a2a"
    if mode == "p2p":
        model.shard.enable_peer_memory()
    return model
