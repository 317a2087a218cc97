"""Which C-ABI call breaks CUDA-graph capture?  Captures each op alone and reports."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops, _C

dev = "cuda"
B, m, k = 4096, 8, 16
rows = 1000
table = torch.randn(m * rows, k, device=dev)
offs = (torch.arange(m + 1, device=dev) * rows).long()
ids = torch.randint(0, rows, (B, m), device=dev)
grad = torch.randn(B, m * k, device=dev)
st = ops.new_status(dev)

def try_capture(name, fn):
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            out = fn()
        g.replay()
        torch.cuda.synchronize()
        print(f"{name}: capture OK", flush=True)
    except Exception as e:
        print(f"{name}: capture FAILED: {str(e).splitlines()[0]}", flush=True)
        torch.cuda.synchronize()

try_capture("gather", lambda: ops.gather(table, offs, ids, status=st))
try_capture("gather_fm_fwd", lambda: ops.gather_fm_fwd(table, None, None, offs, ids, None, None, status=st))
try_capture("workspace_query", lambda: _C.lib.rm_segment_plan_workspace_bytes(B * m))
try_capture("segment_plan", lambda: ops.segment_plan(ids, offs, m * rows))
plan = ops.segment_plan(ids, offs, m * rows)
try_capture("segment_reduce", lambda: ops.segment_reduce(grad, plan, k, ld=m * k))
sg = ops.SparseGrad(plan.uniq_rows, ops.segment_reduce(grad, plan, k, ld=m * k), plan.n_unique)
try_capture("sparse_opt", lambda: ops.sparse_opt_step(table, sg, 0, 0.01, 0.0))
p = torch.randn(1000, device=dev); gg = torch.randn(1000, device=dev)
try_capture("dense_opt", lambda: ops.dense_opt_step(p, gg, 0, 0.01, 0.0))
x = torch.randn(B, 64, device=dev); w = torch.randn(3, 64, device=dev); bb = torch.randn(3, 64, device=dev); wo = torch.randn(64, device=dev); w0 = torch.zeros(1, device=dev)
try_capture("cross_fwd", lambda: ops.cross_fwd(x, w, bb, wo, w0))
logit, dots = ops.cross_fwd(x, w, bb, wo, w0)
go = torch.randn(B, device=dev)
try_capture("cross_bwd", lambda: ops.cross_bwd(x, w, bb, wo, dots, go))
def fwd_bwd():
    t = torch.randn(64, 64, device=dev, requires_grad=True)
    y = (t @ t).sum()
    y.backward()
    return t.grad
try_capture("torch_fwd_bwd", fwd_bwd)
