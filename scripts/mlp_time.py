"""Timing of the narrow-layer backward kernels at the C5 shape (tuning helper)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops
B, m, k, nd, N = 65536, 26, 64, 13, 32
d = m * k + nd; ld = (d + 3) // 4 * 4
x = torch.randn(B, ld, device="cuda"); x[:, d:] = 0
W = torch.randn(d, N, device="cuda") * 0.1
g = torch.randn(B, N, device="cuda"); S = torch.randn(B, k, device="cuda"); gf = torch.randn(B, device="cuda")
out = torch.empty(B, ld, device="cuda"); G = torch.empty(B, m * k, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)
print(json.dumps({"dx": timeit(lambda: ops.linear_bwd_input(g, W, d_ld=ld, out=out)),
                  "dx_fm": timeit(lambda: ops.linear_bwd_input_fm(g, W, m, k, x, S, gf, out=G)),
                  "dW": timeit(lambda: ops.linear_bwd_weight(x, ld, d, g))}))
