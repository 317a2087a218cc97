"""Kernel micro-benchmarks at the C5 shape (tuning harness): CUDA-event times and algorithmic GB/s per kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops

dev = "cuda"
B, m, k, nd = 65536, 26, int(os.environ.get("KB_K", "64")), 13
rows = int(os.environ.get("KB_ROWS", "1000000"))
torch.manual_seed(0)
table = torch.empty(m * rows, k, device=dev).normal_(0.0, 0.01)
bias_t = torch.zeros(m * rows, device=dev); lin_t = torch.zeros(m * rows, device=dev)
offs = (torch.arange(m + 1, device=dev) * rows).long()
ids_pool = [torch.randint(0, rows, (B, m), device=dev) for _ in range(4)]
dense = torch.randn(B, nd, device=dev); lin_dense = torch.randn(nd, device=dev)
st = ops.new_status(dev)

def timeit(fn, n=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
fwd_bytes = B * (m * (8 + 8 * k + 8) + 8 * nd + 4 * k + 8)
t = timeit(lambda i: ops.gather_fm_fwd(table, bias_t, lin_t, offs, ids_pool[i % 4], dense, lin_dense, status=st))
res["gather_fm_fwd"] = (round(t, 4), round(fwd_bytes / t / 1e6, 1))
t = timeit(lambda i: ops.gather(table, offs, ids_pool[i % 4], status=st))
res["gather_fwd"] = (round(t, 4), round(B * m * (8 + 8 * k) / t / 1e6, 1))
t = timeit(lambda i: ops.segment_plan(ids_pool[i % 4], offs, m * rows))
res["segment_plan"] = (round(t, 4), None)
plans = [ops.segment_plan(ids_pool[i], offs, m * rows) for i in range(4)]
nu = plans[0].num_unique()
x, fm, lin, S = ops.gather_fm_fwd(table, bias_t, lin_t, offs, ids_pool[0], dense, lin_dense, status=st)
ld = x.shape[1]
dx = torch.randn(B, ld, device=dev); g_fm = torch.randn(B, device=dev); g_lin = torch.randn(B, device=dev)
bwd_bytes = B * m * (4 + 8 * k) + B * (4 * k + 8) + nu * (4 * k + 8)
for var in ("",):  # (the kernel variants once swept here through environment knobs are gone: one configuration each)
    t = timeit(lambda i: ops.emb_fm_bwd(dx, x, ld, S, g_fm, g_lin, plans[i % 4], k, True, True, True))
    res["emb_fm_bwd"] = (round(t, 4), round(bwd_bytes / t / 1e6, 1))
    grad = dx[:, : m * k].contiguous()
    t = timeit(lambda i: ops.segment_reduce(grad, plans[i % 4], k, ld=m * k))
    res["segment_reduce"] = (round(t, 4), round((B * m * (4 + 4 * k) + nu * (4 * k + 8)) / t / 1e6, 1))
    upd_bytes = bwd_bytes - nu * (4 * k + 8) + nu * (8 + 8 * k + 16)  # uniq id + row r/w + bias/lin r/w per unique row
    t = timeit(lambda i: ops.emb_fm_bwd_update(dx, x, ld, S, g_fm, g_lin, plans[i % 4], k, table, bias_t, lin_t, 0, 1e-3))
    res["emb_fm_bwd_update"] = (round(t, 4), round(upd_bytes / t / 1e6, 1))
rows_out, ob, ol = ops.emb_fm_bwd(dx, x, ld, S, g_fm, g_lin, plans[0], k, True, True, True)
sg = ops.SparseGrad(plans[0].uniq_rows, rows_out, plans[0].n_unique)
t = timeit(lambda i: ops.sparse_opt_step(table, sg, 0, 1e-3, 0.0))
res["sparse_opt k"] = (round(t, 4), round(nu * (8 + 12 * k) / t / 1e6, 1))
a = torch.empty(256 * 1024 * 1024, device=dev); b = torch.empty_like(a)
t = timeit(lambda i: b.copy_(a))
res["torch copy 1GiB (peak ref)"] = (round(t, 4), round(2 * a.numel() * 4 / t / 1e6, 1))
print(json.dumps({"B": B, "m": m, "k": k, "rows": rows, "n_unique": nu, "results(ms, GB/s)": res}, indent=1))
