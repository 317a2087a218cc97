"""(e') the peer-memory kernels on ONE GPU: every "peer" buffer lives on the same device, so the sharding arithmetic,
the owner-side plan and the ordered reduction are checked against the single-table kernels and the numpy oracle
without needing several GPUs (the NVLink run is scripts/dist_check.py / tests/test_dist_nccl_gpu.py)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

SIZES = [50, 7, 1000, 3, 200, 31, 2, 90, 1]


def _shard(full, sizes, W):
    """global concatenated [sum V, ...] table -> W local tables (row r of table f -> rank r % W, local r // W)."""
    offs = np.concatenate([[0], np.cumsum(sizes)])
    local_sizes = [(v + W - 1) // W for v in sizes]
    loffs = np.concatenate([[0], np.cumsum(local_sizes)])
    shards = []
    for r in range(W):
        t = torch.zeros((int(loffs[-1]),) + tuple(full.shape[1:]), dtype=full.dtype)
        for f, v in enumerate(sizes):
            rows = np.arange(r, v, W)
            t[loffs[f] : loffs[f] + len(rows)] = full[offs[f] + rows]
        shards.append(t.cuda())
    return shards, loffs, offs


SIZES26 = [50, 7, 1000, 3, 200, 31, 2, 90, 1, 17, 400, 5, 64, 9, 300, 12, 77, 4, 128, 33, 6, 250, 19, 8, 500, 41]


@pytest.mark.parametrize("SIZES", [SIZES, SIZES26], ids=["m9", "m26"])
@pytest.mark.parametrize("W", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("k", [16, 64])
def test_p2p_front_end_equals_single_table_front_end(W, k, SIZES):
    from recman_b200 import ops

    g = torch.Generator().manual_seed(W * 100 + k)
    total = sum(SIZES)
    m, B, n_dense = len(SIZES), 333, 5
    table = torch.randn(total, k, generator=g) * 0.1
    bias = torch.randn(total, generator=g) * 0.1
    lin = torch.randn(total, generator=g) * 0.1
    lin_dense = (torch.randn(n_dense, generator=g) * 0.1).cuda()
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in SIZES], 1).contiguous().cuda()
    dense = torch.randn(B, n_dense, generator=g).cuda()
    offs = torch.tensor(np.concatenate([[0], np.cumsum(SIZES)]), dtype=torch.int64).cuda()
    x0, fm0, lin0, S0 = ops.gather_fm_fwd(table.cuda(), bias.cuda(), lin.cuda(), offs, ids, dense, lin_dense)
    tabs, loffs, _ = _shard(table, SIZES, W)
    biases, _, _ = _shard(bias, SIZES, W)
    lins, _, _ = _shard(lin, SIZES, W)
    st = ops.new_status("cuda")
    x, fm, ln, S = ops.gather_fm_fwd_p2p(
        [t.data_ptr() for t in tabs], [t.data_ptr() for t in biases], [t.data_ptr() for t in lins], k,
        torch.tensor(SIZES, dtype=torch.int64).cuda(), torch.tensor(loffs[:-1], dtype=torch.int64).cuda(), ids, dense,
        lin_dense, status=st)
    assert int(st.item()) == 0
    assert torch.equal(x, x0) and torch.equal(fm, fm0) and torch.equal(ln, lin0) and torch.equal(S, S0)
    # an id outside its table: flagged, row zero-filled
    bad = ids.clone()
    bad[5, 2] = SIZES[2]
    x2, *_ = ops.gather_fm_fwd_p2p([t.data_ptr() for t in tabs], None, None, k,
                                   torch.tensor(SIZES, dtype=torch.int64).cuda(),
                                   torch.tensor(loffs[:-1], dtype=torch.int64).cuda(), bad, dense, None, status=st)
    assert int(st.item()) == 1 and torch.all(x2[5, 2 * k : 3 * k] == 0)


@pytest.mark.parametrize("W", [2, 3, 4, 8])
def test_shard_plan_and_peer_reduce_match_numpy(W):
    """Owner-side plan over the gathered ids + reduction that pulls rows from the per-rank gradient buffers:
    unique rows / order bit-exact, sums bit-identical to the sequential fp32 oracle in ascending global position."""
    from recman_b200 import ops

    k, KP, b = 16, 20, 97
    m = len(SIZES)
    rng = np.random.RandomState(W)
    gids = np.stack([rng.randint(0, min(v, 40), size=W * b) for v in SIZES], 1).astype(np.int64)  # many duplicates
    G = [torch.from_numpy(rng.randn(b * m, KP).astype(np.float32)).cuda() for _ in range(W)]
    local_sizes = [(v + W - 1) // W for v in SIZES]
    loffs = np.concatenate([[0], np.cumsum(local_sizes)])
    fs = torch.tensor(SIZES, dtype=torch.int64).cuda()
    lo = torch.tensor(loffs[:-1], dtype=torch.int64).cuda()
    gids_d = torch.from_numpy(gids).cuda()
    Gall = np.concatenate([g.cpu().numpy() for g in G], 0)  # row gp = src*(b*m) + p
    for rank in range(W):
        st = ops.new_status("cuda")
        plan = ops.shard_plan(gids_d, W, rank, fs, lo, int(loffs[-1]), W * b * m, st)
        n_own, nu = int(plan.n_own.item()), int(plan.n_unique.item())
        flat = gids.reshape(-1)
        own = np.flatnonzero(flat % W == rank)
        keys = loffs[own % m] + flat[own] // W
        assert n_own == len(own) and int(st.item()) == 0
        uniq, sums, order, seg = oracle.segment_sum_sorted(keys, Gall[own])
        assert nu == len(uniq)
        assert np.array_equal(plan.uniq_rows[:nu].cpu().numpy(), uniq)
        assert np.array_equal(plan.sorted_pos[:n_own].cpu().numpy(), own[order])
        assert np.array_equal(plan.seg_start[: nu + 1].cpu().numpy(), seg)
        rows, ob, ol = ops.segment_reduce_p2p([g.data_ptr() for g in G], b * m, KP, k, plan)
        assert np.array_equal(rows[:nu].cpu().numpy(), sums[:, :k])
        assert np.array_equal(ob[:nu].cpu().numpy(), sums[:, k]) and np.array_equal(ol[:nu].cpu().numpy(), sums[:, k + 1])
        # fused update == reduce then rm_sparse_opt_step
        T = int(loffs[-1])
        tab = torch.from_numpy(rng.randn(T, k).astype(np.float32)).cuda()
        bt = torch.from_numpy(rng.randn(T).astype(np.float32)).cuda()
        lt = torch.from_numpy(rng.randn(T).astype(np.float32)).cuda()
        tab2, bt2, lt2 = tab.clone(), bt.clone(), lt.clone()
        ops.sparse_opt_step(tab, ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique), 0, 0.01)
        ops.sparse_opt_step(bt, ops.SparseGrad(plan.uniq_rows, ob, plan.n_unique), 0, 0.01)
        ops.sparse_opt_step(lt, ops.SparseGrad(plan.uniq_rows, ol, plan.n_unique), 0, 0.01)
        ops.segment_reduce_p2p_update([g.data_ptr() for g in G], b * m, KP, k, plan, tab2, bt2, lt2, 0, 0.01)
        assert torch.equal(tab, tab2) and torch.equal(bt, bt2) and torch.equal(lt, lt2)


def test_shard_plan_capacity_overflow_is_flagged():
    from recman_b200 import ops

    W, b, m = 2, 64, 3
    sizes = [10, 10, 10]
    gids = torch.zeros(W * b, m, dtype=torch.int64, device="cuda")  # every id -> rank 0
    fs = torch.tensor(sizes, dtype=torch.int64).cuda()
    lo = torch.tensor([0, 5, 10], dtype=torch.int64).cuda()
    st = ops.new_status("cuda")
    plan = ops.shard_plan(gids, W, 0, fs, lo, 15, 200, st)  # 384 owned entries > capacity 200
    torch.cuda.synchronize()
    assert int(plan.n_own.item()) == W * b * m and int(st.item()) & 4
    assert int(plan.seg_start[int(plan.n_unique.item())].item()) == 200  # clamped, nothing out of bounds
    st.zero_()
    plan1 = ops.shard_plan(gids, W, 1, fs, lo, 15, 200, st)  # rank 1 owns nothing
    assert int(plan1.n_own.item()) == 0 and int(plan1.n_unique.item()) == 0 and int(st.item()) == 0


@pytest.mark.parametrize("W", [2, 5, 8])
def test_peer_reduce_with_gathered_per_sample_scalars(W):
    """k-wide gradient rows + all-gathered per-sample (g_bias, g_lin) == (k+4)-wide rows that carry them, bit for bit."""
    from recman_b200 import ops

    k, b = 16, 61
    m = len(SIZES)
    rng = np.random.RandomState(10 + W)
    gids = np.stack([rng.randint(0, min(v, 30), size=W * b) for v in SIZES], 1).astype(np.int64)
    gs = rng.randn(W * b, 2).astype(np.float32)  # per global sample
    G20, Gk = [], []
    for r in range(W):
        rows = rng.randn(b * m, k + 4).astype(np.float32)
        rows[:, k:] = 0
        rows[:, k : k + 2] = np.repeat(gs[r * b : (r + 1) * b], m, axis=0)
        G20.append(torch.from_numpy(rows).cuda())
        Gk.append(torch.from_numpy(np.ascontiguousarray(rows[:, :k])).cuda())
    local_sizes = [(v + W - 1) // W for v in SIZES]
    loffs = np.concatenate([[0], np.cumsum(local_sizes)])
    fs = torch.tensor(SIZES, dtype=torch.int64).cuda()
    lo = torch.tensor(loffs[:-1], dtype=torch.int64).cuda()
    gids_d = torch.from_numpy(gids).cuda()
    gs_d = torch.from_numpy(gs).cuda()
    for rank in (0, W - 1):
        plan = ops.shard_plan(gids_d, W, rank, fs, lo, int(loffs[-1]), W * b * m, ops.new_status("cuda"))
        nu = int(plan.n_unique.item())
        a = ops.segment_reduce_p2p([g.data_ptr() for g in G20], b * m, k + 4, k, plan)
        c = ops.segment_reduce_p2p([g.data_ptr() for g in Gk], b * m, k, k, plan, gscal=gs_d, m=m)
        for u, v in zip(a, c):
            assert torch.equal(u[:nu], v[:nu])
