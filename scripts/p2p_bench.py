"""NVLink peer-memory microbenchmark (torchrun, 2+ ranks): random 256-byte-row reads from / writes to a PEER buffer
of growing size, through the library's own kernels (rm_gather_fwd with a peer table pointer, rm_unpack_rows with a
peer destination).  Answers: does random access into a large peer mapping fall off a (TLB) cliff?"""
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
    k, n = 64, 26 * 65536
    peer = (rank + 1) % world
    res = {}
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty(n, k, device="cuda")
    src = torch.randn(n, k + 4, device="cuda")
    for gb in [float(v) for v in os.environ.get("SIZES", "0.5,2,8,32").split(",")]:
        V = int(gb * (1 << 30) // (k * 4))
        t = pm.alloc((V, k), zero=False)
        ptrs = pm.ptrs_of(t)
        ids = torch.randint(0, V, (n, 1), device="cuda", dtype=torch.int64)
        ids_sorted = torch.sort(ids.reshape(-1)).values.reshape(n, 1).contiguous()
        pos = torch.randperm(V, device="cuda")[:n].to(torch.int32).contiguous() if V >= n else torch.randint(0, V, (n,), device="cuda", dtype=torch.int32)
        offs = torch.tensor([0, V], dtype=torch.int64, device="cuda")
        dist.barrier(); torch.cuda.synchronize()

        def timeit(fn, iters=5):
            fn(); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        def rd(ptr, idt):
            return lambda: _C.call("rm_gather_fwd", ptr, offs.data_ptr(), idt.data_ptr(), n, 1, k, out.data_ptr(), k, None, st)

        def wr(ptr):
            # rows j -> x[pos[j]] (m=1, ld=k): scattered 256-byte writes
            return lambda: _C.call("rm_unpack_rows", src.data_ptr(), n, k + 4, pos.data_ptr(), 1, k, ptr, k, None, None, st)

        r = {}
        for name, fn in [("read_local", rd(ptrs[rank], ids)), ("read_peer", rd(ptrs[peer], ids)),
                         ("read_peer_sorted", rd(ptrs[peer], ids_sorted)), ("write_local", wr(ptrs[rank])),
                         ("write_peer", wr(ptrs[peer]))]:
            ms = timeit(fn)
            r[name] = {"ms": round(ms, 4), "GBs_rows": round(n * k * 4 / ms / 1e6, 1)}
            dist.barrier()
        res[f"{gb}GB"] = r
        if rank == 0:
            print(gb, json.dumps(r), flush=True)
        del t
    if rank == 0:
        json.dump(res, open("gpurun_out/p2p_bench.json", "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


main()
