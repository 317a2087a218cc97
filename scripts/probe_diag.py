"""Diagnostics for the MN-major UMMA operand layout (rm_umma_probe): prints where a single non-zero lands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recman_b200 import ops

torch.manual_seed(0)
for variant in (0,):  # variant 1 (LBO / SBO swapped) reads outside the tile: illegal address on the device
    K = 64
    At = torch.randint(-4, 5, (K, 128)).float()
    Bt = torch.randint(-4, 5, (K, 32)).float()
    D, st = ops.umma_probe(At.cuda(), Bt.cuda(), variant)
    torch.cuda.synchronize()
    exp = At.t() @ Bt
    err = (D.cpu() - exp).abs()
    print(f"variant {variant}: status {int(st.item())} max err {float(err.max())} bad rows {int((err.max(1).values > 0).sum())} "
          f"bad cols {int((err.max(0).values > 0).sum())}")
    # single non-zero in A at (k0, m0), B = ones in row k0 with value n+1
    for (k0, m0) in [(0, 0), (0, 5), (0, 37), (0, 100), (3, 0), (9, 2), (17, 64), (63, 127)]:
        At = torch.zeros(K, 128); At[k0, m0] = 1.0
        Bt = torch.zeros(K, 32); Bt[k0] = torch.arange(1, 33).float()
        D, st = ops.umma_probe(At.cuda(), Bt.cuda(), variant)
        D = D.cpu()
        nz = torch.nonzero(D)
        rows = sorted(set(nz[:, 0].tolist()))
        print(f"  A[{k0},{m0}]=1: nonzero rows {rows[:8]} (n={len(rows)}), row values {D[rows[0]].tolist()[:8] if rows else None}")
    # K = 8 only (one k-step)
    At = torch.randint(-4, 5, (8, 128)).float(); Bt = torch.randint(-4, 5, (8, 32)).float()
    D, st = ops.umma_probe(At.cuda(), Bt.cuda(), variant)
    print(f"  K=8: max err {float((D.cpu() - At.t() @ Bt).abs().max())}")
