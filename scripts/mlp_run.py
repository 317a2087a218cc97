"""Tiny driver for profiling the narrow-layer backward kernels at the C5 shape (used under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops

B, d, ld, N = 65536, 1677, 1680, 32
x = torch.randn(B, ld, device="cuda"); x[:, d:] = 0
W = torch.randn(d, N, device="cuda") * 0.1
g = torch.randn(B, N, device="cuda")
out = torch.empty(B, ld, device="cuda")
for _ in range(3):
    ops.linear_bwd_input(g, W, d_ld=ld, out=out)
    ops.linear_bwd_weight(x, ld, d, g)
torch.cuda.synchronize()
