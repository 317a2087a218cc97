"""(e) multi-GPU routing on CPU: world_size-2 gloo processes exercise the sharding / all-to-all index arithmetic
(ShardPlan.route, build_exchange, return trip, gradient routing) against the single-process oracle lookup."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_sort(dest):
    return torch.argsort(dest, stable=True).to(torch.int32)


def _worker(rank, world, port, sizes, k, b, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recman_b200.th.dist import ShardPlan, allreduce_dense, build_exchange

    try:
        m = len(sizes)
        g = torch.Generator().manual_seed(0)  # same on every rank: the "global" tables
        tables = [torch.randn(v, k, generator=g) for v in sizes]
        plan = ShardPlan(sizes, world, rank)
        # this rank's shard: row r of table f -> rank r % W at local row r // W
        local = torch.zeros(plan.total_local, k)
        for f in range(m):
            rows = plan.local_rows_of(f)
            local[plan.local_offsets[f] : plan.local_offsets[f] + rows.numel()] = tables[f][rows]
        gi = torch.Generator().manual_seed(100 + rank)  # every rank has its own local batch
        ids = torch.stack([torch.randint(0, v, (b,), generator=gi) for v in sizes], 1)
        ids[0] = 0
        ids[-1] = torch.tensor([v - 1 for v in sizes])
        owner, local_row = oracle.shard_route(ids.numpy(), world)
        dest, key = plan.route(ids)
        assert np.array_equal(dest.numpy().reshape(b, m), owner)
        assert np.array_equal(key.numpy().reshape(b, m), local_row + np.asarray(plan.local_offsets[:-1])[None, :])

        ex = build_exchange(plan, ids, _cpu_sort)
        assert sum(ex.send_splits) == b * m
        # owner side: serve the requested rows (torch indexing stands in for the CUDA gather in this CPU test)
        rows = local[ex.recv_keys]
        back = torch.empty(b * m, k)
        dist.all_to_all_single(back, rows, ex.send_splits, ex.recv_splits)
        got = torch.empty(b * m, k)
        got[ex.sorted_pos.long()] = back
        exp, _ = oracle.feat_embedding_layer(tables, [ids[:, f] for f in range(m)])
        assert torch.equal(got.reshape(b, m, k), exp), "sharded lookup must be bit-identical to the single-process one"

        # backward: gradient rows travel to the owners in the same routed order and are segment-summed there
        gg = torch.Generator().manual_seed(200 + rank)
        grad = torch.randn(b * m, k, generator=gg)
        send = grad[ex.sorted_pos.long()]
        recv = torch.empty(len(ex.recv_keys), k)
        dist.all_to_all_single(recv, send, ex.recv_splits, ex.send_splits)
        uniq, sums, _, _ = oracle.segment_sum_sorted(ex.recv_keys.numpy(), recv.numpy())
        shard_grad = np.zeros((plan.total_local, k), dtype=np.float64)
        shard_grad[uniq] = sums
        # reference: gather every rank's (ids, grad), build the global dense gradient, cut out this rank's shard
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        all_grad = [torch.empty_like(grad) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        dist.all_gather(all_grad, grad)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        dense = np.zeros((int(offs[-1]), k))
        for i_, g_ in zip(all_ids, all_grad):
            dense += oracle.dense_table_grad(oracle.global_rows(i_.numpy(), offs).reshape(-1), g_.numpy(), int(offs[-1]))
        for f in range(m):
            rows_f = plan.local_rows_of(f).numpy()
            np.testing.assert_allclose(shard_grad[plan.local_offsets[f] : plan.local_offsets[f] + len(rows_f)],
                                       dense[offs[f] + rows_f], rtol=1e-5, atol=1e-5)

        # dense-gradient all-reduce bucket
        a, c = torch.full((3, 2), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))
        allreduce_dense([a, c])
        tot = sum(range(1, world + 1))
        assert torch.all(a == tot) and torch.all(c == 10.0 * tot)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _tower_worker(rank, world, port, sizes, k, N1, b, out_dir):
    """The data flow of the row-sharded fused tower (th/dist.py:P2PTowerFunction) restated on CPU: ids and the
    per-SAMPLE operands (g1, S, g_fm) are all-gathered, every owner forms the gradient rows of the positions it owns
    itself (dx = g1 . W1_f^T, FM term from S and its own row) and sums duplicates - no gradient row crosses ranks.
    Must equal this rank's shard of the dense gradient of the concatenated global batch."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recman_b200.th.dist import ShardPlan, _all_gather_many

    try:
        m = len(sizes)
        g = torch.Generator().manual_seed(0)  # same on every rank
        tables = [torch.randn(v, k, generator=g, dtype=torch.float64) for v in sizes]
        W1 = torch.randn(m * k, N1, generator=g, dtype=torch.float64) * 0.1
        plan = ShardPlan(sizes, world, rank)
        local = torch.zeros(plan.total_local, k, dtype=torch.float64)
        for f in range(m):
            rows = plan.local_rows_of(f)
            local[plan.local_offsets[f] : plan.local_offsets[f] + rows.numel()] = tables[f][rows]
        gi = torch.Generator().manual_seed(100 + rank)
        ids = torch.stack([torch.randint(0, v, (b,), generator=gi) for v in sizes], 1)
        ids[:3] = 0  # duplicates inside a batch and across ranks
        x = torch.stack([tables[f][ids[:, f]] for f in range(m)], 1)  # [b, m, k] (the forward reads rows from their owners)
        S = x.sum(1)
        g1 = torch.randn(b, N1, generator=gi, dtype=torch.float64)
        g_fm = torch.randn(b, generator=gi, dtype=torch.float64)
        # the step's collectives: ids as int32 + one grouped all-gather of the per-sample operands
        ids32 = ids.to(torch.int32)
        gids = torch.empty(world * b, m, dtype=torch.int32)
        S_all, g1_all, gfm_all = (torch.empty(world * b, k, dtype=torch.float64), torch.empty(world * b, N1, dtype=torch.float64),
                                  torch.empty(world * b, dtype=torch.float64))
        dist.all_gather_into_tensor(gids, ids32)
        _all_gather_many([(S_all, S), (g1_all, g1), (gfm_all, g_fm)])
        assert torch.equal(gids[rank * b : (rank + 1) * b], ids32) and torch.equal(S_all[rank * b : (rank + 1) * b], S)
        # owner side: positions in ascending global position gp = sample * m + f (rank-major samples)
        shard_grad = np.zeros((plan.total_local, k))
        dW1_part = np.zeros((m * k, N1))
        gid = gids.numpy().astype(np.int64)
        for sample in range(world * b):
            for f in range(m):
                i = gid[sample, f]
                if i % world != rank:
                    continue
                lr = plan.local_offsets[f] + i // world
                row = local[lr].numpy()
                dx = g1_all[sample].numpy() @ W1[f * k : (f + 1) * k].numpy().T
                shard_grad[lr] += dx + gfm_all[sample].item() * (S_all[sample].numpy() - row)
                dW1_part[f * k : (f + 1) * k] += np.outer(row, g1_all[sample].numpy())
        # reference: dense gradient of the global tables over the concatenated batch
        offs = np.concatenate([[0], np.cumsum(sizes)])
        dense = np.zeros((int(offs[-1]), k))
        dW1_ref = np.zeros((m * k, N1))
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            ids_r = torch.stack([torch.randint(0, v, (b,), generator=gr) for v in sizes], 1)
            ids_r[:3] = 0
            x_r = torch.stack([tables[f][ids_r[:, f]] for f in range(m)], 1)
            S_r = x_r.sum(1)
            g1_r = torch.randn(b, N1, generator=gr, dtype=torch.float64)
            gfm_r = torch.randn(b, generator=gr, dtype=torch.float64)
            dx_r = (g1_r @ W1.T).reshape(b, m, k) + gfm_r[:, None, None] * (S_r[:, None, :] - x_r)
            for f in range(m):
                np.add.at(dense, offs[f] + ids_r[:, f].numpy(), dx_r[:, f].numpy())
            dW1_ref += x_r.reshape(b, m * k).numpy().T @ g1_r.numpy()
        for f in range(m):
            rows_f = plan.local_rows_of(f).numpy()
            np.testing.assert_allclose(shard_grad[plan.local_offsets[f] : plan.local_offsets[f] + len(rows_f)],
                                       dense[offs[f] + rows_f], rtol=1e-10, atol=1e-10)
        # the owners' dW1 partials add up to the global dW1 (they join the dense all-reduce)
        t = torch.from_numpy(dW1_part)
        dist.all_reduce(t)
        np.testing.assert_allclose(t.numpy(), dW1_ref, rtol=1e-10, atol=1e-10)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_tower_dataflow_gloo(tmp_path, world):
    sizes = [11, 1, 40, 3]
    mp.spawn(_tower_worker, args=(world, _free_port(), sizes, 4, 3, 9, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_routing_gloo(tmp_path, world):
    sizes = [11, 1, 40, 7]
    mp.spawn(_worker, args=(world, _free_port(), sizes, 4, 33, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_plan_layout():
    from recman_b200.th.dist import ShardPlan

    p = ShardPlan([10, 3, 8], world=4, rank=1)
    assert p.local_sizes == [3, 1, 2] and p.local_offsets == [0, 3, 4, 6] and p.total_local == 6
    assert p.local_rows_of(0).tolist() == [1, 5, 9] and p.local_rows_of(1).tolist() == [1]
    dest, key = p.route(torch.tensor([[9, 2, 7], [0, 0, 0]]))
    assert dest.tolist() == [1, 2, 3, 0, 0, 0] and key.tolist() == [2, 3, 5, 0, 3, 4]


def test_shard_plan_capacity_covers_skewed_and_tiny_tables():
    """Owner-side plan capacity: tables with fewer rows than ranks load the low ranks, Zipf ids load the owners of the
    hot rows (id 0 of every table lives on rank 0) - the default capacity has to hold both."""
    from recman_b200.th.dist import ShardPlan

    W, b = 8, 300
    sizes = [50, 7, 1000, 3, 200, 31, 2, 90, 1]
    plan = ShardPlan(sizes, W, 0)
    rng = np.random.RandomState(0)
    ids = np.stack([rng.randint(0, v, size=W * b) for v in sizes], 1)
    owned = np.bincount((ids % W).reshape(-1), minlength=W)
    assert owned.max() <= plan.capacity(b) <= W * b * len(sizes)
    # Criteo-shaped, Zipf(1.05): rank 0 owns ~1.3x the uniform share at W = 8
    m, rows, b = 26, 10_000_000, 8192
    plan = ShardPlan([rows] * m, W, 0)
    z = (np.random.RandomState(1).zipf(1.05, size=(W * b, m)) - 1) % rows
    owned = np.bincount((z % W).reshape(-1), minlength=W)
    assert owned.max() > 1.2 * b * m and owned.max() <= plan.capacity(b)
