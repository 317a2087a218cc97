"""Microbenchmark of rm_gather_fm_fwd_p2p itself with local / peer / mixed table pointers (2 ranks)."""
import ctypes, json, os, sys
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
    for rows in [int(v) for v in os.environ.get("ROWS", "100000,2000000,8000000").split(",")]:
        Vl = rows // 2  # local rows per table at W=2
        t = pm.alloc((m * Vl, k), zero=False)
        ptrs = pm.ptrs_of(t)
        fs = torch.full((m,), rows, dtype=torch.int64, device="cuda")
        lo = (torch.arange(m, dtype=torch.int64, device="cuda") * Vl).contiguous()
        ids = torch.randint(0, rows, (B, m), device="cuda", dtype=torch.int64)
        ids_even = (ids // 2 * 2).contiguous()      # every id owned by "rank 0" slot
        ids_odd = (ids // 2 * 2 + 1).clamp(max=rows - 1).contiguous()
        dense = torch.randn(B, 13, device="cuda")
        res = {}

        def run(tabs, idt):
            return lambda: ops.gather_fm_fwd_p2p(tabs, None, None, k, fs, lo, idt, dense, None)

        def timeit(fn, iters=5):
            fn(); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return round(e0.elapsed_time(e1) / iters, 4)

        L, P = ptrs[rank], ptrs[peer]
        for name, tabs, idt in [("LL_mixed_ids", [L, L], ids), ("LP_mixed_ids", [L, P], ids), ("PL_mixed_ids", [P, L], ids),
                                ("PP_mixed_ids", [P, P], ids), ("LP_all_local", [L, P], ids_even),
                                ("LP_all_peer", [L, P], ids_odd), ("W1_local", [L], (ids // 2).contiguous()),
                                ("W1_peer", [P], (ids // 2).contiguous())]:
            if len(tabs) == 1:
                fs1 = torch.full((m,), Vl, dtype=torch.int64, device="cuda")
                fn = (lambda tabs=tabs, idt=idt, fs1=fs1: ops.gather_fm_fwd_p2p(tabs, None, None, k, fs1, lo, idt, dense, None))
            else:
                fn = run(tabs, idt)
            res[name] = timeit(fn)
            dist.barrier()
        if rank == 0:
            print(rows, json.dumps(res), flush=True)
        del t
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


main()
