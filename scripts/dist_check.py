"""Multi-GPU parity check (run under torchrun, one rank per GPU):
row-sharded DeepFM over W ranks == single-GPU DeepFM on the concatenated global batch.
Gathered rows / logits per sample and every gradient are compared; then one optimizer step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from recman_b200.autograd import dense_table_grad
    from recman_b200.th import DeepFM
    from recman_b200.th import dist as rdist
    from recman_b200.th.input import DataInputs
    from tests import parity_util as pu

    k = int(os.environ.get("DIST_K", "16"))
    sizes = [50, 7, 1000, 3, 200, 31, 2, 90, 1]
    b = 300
    fd = pu.make_feat_dict(sizes, n_dense=5)
    kw = dict(embedding_size=k, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=b, embedding_l2_reg=0.0,
              linear_l2_reg=0.0, deep_l2_reg=1e-5, learning_rate=0.01)
    Xr, yr = pu.synth_batch(fd, b, seed=100 + rank)

    # ---- single-GPU reference on the global batch (same on every rank) ----
    ref = DeepFM(fd, **kw)
    ref.hparams["tower"] = False  # gradient-level comparison: the separate kernels emit the sparse gradients
    batches = [pu.synth_batch(fd, b, seed=100 + r) for r in range(world)]
    Xg = {n: np.concatenate([bt[0][n] for bt in batches]) for n in Xr}
    yg = np.concatenate([bt[1] for bt in batches])
    with torch.no_grad():
        ref._out(DataInputs("cuda").load(fd, Xg, yg))
    pu.randomize_variables(ref, seed=1)
    ref_logit, ref_loss, ref_grads = pu.run_model_step(ref, Xg, yg)

    # ---- sharded model ----
    model = DeepFM(fd, **kw)
    model.hparams["tower"] = False
    mode = os.environ.get("DIST_MODE", "auto")
    rdist.shard_model(model, world, rank, mode=mode)
    with torch.no_grad():
        model._out(DataInputs("cuda").load(fd, Xr, yr))
    plan = model.shard
    offs = np.concatenate([[0], np.cumsum(sizes)])
    total = int(offs[-1])
    rows_of = [plan.local_rows_of(f) + int(offs[f]) for f in range(len(sizes))]  # global rows held by this rank
    for name, p in model.variables.items():
        g = ref.variables[name].data
        if name in ("feat_embed_table", "feat_bias_table"):
            for f, rows in enumerate(rows_of):
                lo = plan.local_offsets[f]
                p.data[lo : lo + rows.numel()] = g[rows.cuda()]
        elif name == "linear_w":
            for f, rows in enumerate(rows_of):
                lo = plan.local_offsets[f]
                p.data[lo : lo + rows.numel()] = g[rows.cuda()]
            p.data[plan.total_local :] = g[total:]
        else:
            p.data.copy_(g)
    logit, loss, _ = None, None, None
    for p in model.variables.values():
        p.grad = None
        p.rm_sparse_grads = []
        p.rm_dense_tail = None
    inputs = DataInputs("cuda").load(fd, Xr, yr)
    loss = model._loss(inputs)
    (loss / world).backward()
    model.check_ids()
    logit = model.final_logit.detach().cpu().reshape(-1)
    torch.testing.assert_close(logit.double(), ref_logit.reshape(-1)[rank * b : (rank + 1) * b].double(), rtol=1e-5, atol=2e-6)

    def close(name, got, exp):
        if exp.numel() == 0:  # a table with fewer rows than ranks leaves some shards empty
            assert got.numel() == 0
            return
        got, exp = got.double().cpu(), exp.double().cpu()
        atol = 1e-5 * max(float(exp.abs().max()), 1e-30)
        torch.testing.assert_close(got, exp, rtol=1e-5, atol=atol, msg=lambda m: f"{name}: {m}")

    for name, p in model.variables.items():
        exp = ref_grads[name]
        if name in ("feat_embed_table", "feat_bias_table", "linear_w"):
            got = dense_table_grad(p).reshape(p.shape[0], -1)  # this rank's shard (owner-summed, not all-reduced)
            e2 = exp.reshape(exp.shape[0], -1)
            for f, rows in enumerate(rows_of):
                lo = plan.local_offsets[f]
                close(f"{name}[table {f}]", got[lo : lo + rows.numel()], e2[rows])
            if name == "linear_w":
                tail = p.rm_dense_tail[1].clone()
                dist.all_reduce(tail)
                close("linear_w[dense tail]", tail, e2[total:].reshape(-1))
        else:
            g = p.grad.clone()
            dist.all_reduce(g)
            close(name, g, exp)
    # drop this script's autograd graph: it keeps the parameters' grad accumulators alive, and those were created on
    # the default stream - a captured backward that reused them would make the legacy stream wait on the capture
    del loss
    model.final_logit = None
    # one full step through the public call (all-reduce of the dense bucket inside optimizer_step)
    model.fit_on_batch(Xr, yr)
    torch.cuda.synchronize()
    dist.barrier()
    if model.shard.peer is not None and os.environ.get("DIST_GRAPH", "0") == "1":
        # the peer-memory step has no host sync: capture it and replay twice
        inputs = DataInputs("cuda").load(fd, Xr, yr)
        model.compile_step(inputs, warmup=1)
        l1 = model.fit_on_batch(inputs, None).clone()
        l2 = model.fit_on_batch(inputs, None).clone()
        torch.cuda.synchronize()
        model.check_ids()
        assert torch.isfinite(l1).all() and torch.isfinite(l2).all() and float(l2) < float(l1) + 1e-3, (l1, l2)
        dist.barrier()
    # the fused tower path (k = 64, peer memory): logits and every parameter after one step, sharded == single GPU
    tower_msg = ""
    if model.shard.peer is not None:
        import bench

        tower_msg = " | tower: " + bench.sharded_parity_check(world, rank, torch.device("cuda", local_rank))
    if rank == 0:
        print(f"dist_check ok: world={world} k={k} mode={'p2p' if model.shard.peer is not None else 'a2a'} "
              f"sharded DeepFM == single-GPU on the global batch{tower_msg}", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
