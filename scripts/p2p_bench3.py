"""Why is rm_gather_fm_fwd_p2p slow over a peer pointer when rm_gather_fwd is not?  (torchrun, 2 ranks)

Same allocation, same ids, four access patterns:
  vec_m26_peer   rm_gather_fwd, 26 tables, ids [B,26], table pointer = peer      (kernel of p2p_bench, data of p2p_bench2)
  fm_m26_peer    rm_gather_fm_fwd_p2p W=1 peer, lockstep field walk              (the slow case)
  fm_m1_peer     rm_gather_fm_fwd_p2p W=1 peer, m=1 over the whole allocation
  push_m26       rm_gather_fwd, LOCAL table, output rows stored into the PEER's row buffer (owner-push)
"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from recman_b200 import _C, ops
    from recman_b200.th.dist import PeerMemory

    pm = PeerMemory(world, rank)
    k, m, B = 64, 26, 65536
    peer = (rank + 1) % world
    st = torch.cuda.current_stream().cuda_stream
    xpeer = pm.alloc((B, m * k), zero=True)
    xptrs = pm.ptrs_of(xpeer)
    for Vl in [int(v) for v in os.environ.get("ROWS", "50000,1000000,5000000").split(",")]:
        t = pm.alloc((m * Vl, k), zero=False)
        ptrs = pm.ptrs_of(t)
        fs = torch.full((m,), Vl, dtype=torch.int64, device="cuda")
        lo = (torch.arange(m, dtype=torch.int64, device="cuda") * Vl).contiguous()
        offs = (torch.arange(m + 1, dtype=torch.int64, device="cuda") * Vl).contiguous()
        ids = torch.randint(0, Vl, (B, m), device="cuda", dtype=torch.int64)
        ids1 = torch.randint(0, m * Vl, (B * m, 1), device="cuda", dtype=torch.int64)
        fs1 = torch.full((1,), m * Vl, dtype=torch.int64, device="cuda")
        lo1 = torch.zeros(1, dtype=torch.int64, device="cuda")
        dense = torch.randn(B, 13, device="cuda")
        out = torch.empty(B, m * k, device="cuda")
        res = {}

        def timeit(fn, iters=5):
            fn(); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return round(e0.elapsed_time(e1) / iters, 4)

        def vec(tab, dst):
            return lambda: _C.call("rm_gather_fwd", tab, offs.data_ptr(), ids.data_ptr(), B, m, k, dst, m * k, None, st)

        def fm(tab, idt, fsz, loff, smode):
            def f():
                ops.gather_fm_fwd_p2p([tab], None, None, k, fsz, loff, idt, None, None)
            return f

        L, P = ptrs[rank], ptrs[peer]
        for name, fn in [("vec_m26_local", vec(L, out.data_ptr())), ("vec_m26_peer", vec(P, out.data_ptr())),
                         ("fm_m26_local", fm(L, ids, fs, lo, 0)), ("fm_m26_peer", fm(P, ids, fs, lo, 0)),
                         ("fm_m1_peer", fm(P, ids1, fs1, lo1, 0)),
                         ("push_m26", vec(L, xptrs[peer]))]:
            res[name] = timeit(fn)
            dist.barrier()
        if rank == 0:
            print(Vl, json.dumps(res), flush=True)
        del t
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


main()
