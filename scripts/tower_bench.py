"""Fused-tower kernel micro-benchmark at the C5 shape: CUDA-event times + algorithmic GB/s of rm_tower_fwd,
rm_tower_plan and rm_tower_bwd_update.  env: KB_ROWS (rows per table, default 10M), KB_ITERS, KB_ONLY (fwd|bwd|plan)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops

dev = "cuda"
B, m, k, nd, N1 = int(os.environ.get("KB_B", "65536")), 26, 64, 13, 32
rows = int(os.environ.get("KB_ROWS", "10000000"))
iters = int(os.environ.get("KB_ITERS", "20"))
only = os.environ.get("KB_ONLY", "")
torch.manual_seed(0)
table = torch.empty(m * rows, k, device=dev).normal_(0.0, 0.01)
scal = torch.zeros(m * rows + nd, 2, device=dev)
offs = (torch.arange(m + 1, device=dev) * rows).long()
ids_pool = [torch.randint(0, rows, (B, m), device=dev) for _ in range(4)]
dense = torch.randn(B, nd, device=dev)
W1 = torch.randn(m * k + nd, N1, device=dev) * 0.05
b1 = torch.zeros(N1, device=dev)
st = ops.new_status(dev)

def timeit(fn, n=iters):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
lin_dense = scal[m * rows:, 1]
fwd_bytes = B * (m * (8 + 4 * k + 8) + 4 * nd + 4 * k + 4 * N1 + 8)
if only in ("", "fwd"):
    t = timeit(lambda i: ops.tower_fwd(table, scal[: m * rows], offs, ids_pool[i % 4], dense, lin_dense, W1, b1, status=st))
    res["tower_fwd"] = (round(t, 4), round(fwd_bytes / t / 1e6, 1))
    t = timeit(lambda i: ops.tower_fwd(table, scal[: m * rows], offs, ids_pool[i % 4], dense, lin_dense, W1, b1, want_x=True, status=st))
    res["tower_fwd +x"] = (round(t, 4), round((fwd_bytes + B * m * k * 4) / t / 1e6, 1))
if only in ("", "plan"):
    t = timeit(lambda i: ops.tower_plan(ids_pool[i % 4], offs, m * rows, status=st))
    res["tower_plan"] = (round(t, 4), None)
if only in ("", "bwd"):
    plans = [ops.tower_plan(ids_pool[i], offs, m * rows, status=st) for i in range(4)]
    y1, fm, lin, S, _ = ops.tower_fwd(table, scal[: m * rows], offs, ids_pool[0], dense, lin_dense, W1, b1, status=st)
    g1 = torch.randn(B, N1, device=dev) * 1e-3; g_fm = torch.randn(B, device=dev) * 1e-3; g_lin = torch.randn(B, device=dev) * 1e-3
    keys = plans[0].sorted_keys
    nu = int((keys[1:] != keys[:-1]).sum().item()) + 1
    bwd_bytes = B * m * (8 + 4 * k) + B * (4 * N1 + 4 * k + 8) + nu * (4 * k + 16)
    t = timeit(lambda i: ops.tower_bwd_update(table, scal[: m * rows], plans[i % 4], g1, S, g_fm, g_lin, W1, 0, 1e-3, status=st))
    res["tower_bwd_update adam"] = (round(t, 4), round(bwd_bytes / t / 1e6, 1))
    t = timeit(lambda i: ops.tower_bwd_update(table, scal[: m * rows], plans[i % 4], g1, S, g_fm, g_lin, W1, 2, 1e-3, status=st))
    res["tower_bwd_update gd"] = (round(t, 4), round(bwd_bytes / t / 1e6, 1))
    t = timeit(lambda i: ops.tower_bwd_update(table, scal[: m * rows], plans[i % 4], g1, S, g_fm, g_lin, W1, 0, 1e-3, update=False, status=st))
    res["tower_bwd (no update)"] = (round(t, 4), None)
torch.cuda.synchronize()
print(json.dumps({"B": B, "m": m, "k": k, "rows": rows, "status": int(st.item()), "results(ms, GB/s)": res}, indent=1))
