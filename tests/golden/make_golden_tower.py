"""Golden fixture for the fused DeepFM tower (round 2): one small DeepFM step at k = 64, hidden (32, 32), computed by
the CPU oracle in fp64 - forward pieces (first-layer pre-activation, FM / first-order logits, logit, loss), every
gradient the kernels produce (dL/dy1, dL/dlogit, summed embedding-gradient rows, k=1 gradients, dW1, dW2, ...) and the
tables after one fresh-optimizer GD step.  Same role as make_golden.py: the reference cannot run here (no TensorFlow),
so the fixture freezes the oracle and gives the GPU test seeded inputs with expected outputs that do not depend on the
oracle code at test time.

    python tests/golden/make_golden_tower.py        # rewrites tests/golden/tower_v1.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

SIZES = [50, 7, 300, 3, 64]
K, B, ND, N1, N2, LR = 64, 200, 3, 32, 32, 0.5


def inputs(seed=2019):
    g = torch.Generator().manual_seed(seed)
    m = len(SIZES)
    total = sum(SIZES)
    d = dict(
        table=torch.randn(total, K, generator=g) * 0.1,
        scal=torch.randn(total, 2, generator=g) * 0.1,  # (bias, first-order weight) per row
        ids=torch.stack([torch.randint(0, v, (B,), generator=g) for v in SIZES], 1),
        dense=torch.randn(B, ND, generator=g),
        lin_dense=torch.randn(ND, generator=g) * 0.1,
        W1=torch.randn(m * K + ND, N1, generator=g) * 0.05, b1=torch.randn(N1, generator=g) * 0.05,
        W2=torch.randn(N1, N2, generator=g) * 0.2, b2=torch.randn(N2, generator=g) * 0.1,
        w3=torch.randn(N2, 1, generator=g) * 0.2, b3=torch.randn(1, generator=g) * 0.1,
        w0=torch.randn(1, generator=g) * 0.1,
        y=(torch.rand(B, generator=g) < 0.3).float(),
    )
    d["ids"][:5] = 0  # duplicates: five samples share row 0 of every table
    d["ids"][-1] = torch.tensor([v - 1 for v in SIZES])
    return d


def expected(d):
    """fp64 autograd over the oracle's layers (DeepFM composition, tf/core/DeepFM.py:107-163)."""
    m = len(SIZES)
    offs = np.concatenate([[0], np.cumsum(SIZES)]).astype(np.int64)
    leaf = lambda t: t.double().clone().requires_grad_()
    T, SC, W1, b1, W2, b2, w3, b3, w0, LD = map(leaf, (d["table"], d["scal"], d["W1"], d["b1"], d["W2"], d["b2"], d["w3"],
                                                       d["b3"], d["w0"], d["lin_dense"]))
    ids = d["ids"]
    tabs = [T[int(offs[f]):int(offs[f + 1])] for f in range(m)]
    biases = [SC[int(offs[f]):int(offs[f + 1]), 0:1] for f in range(m)]
    embeds, bias = oracle.feat_embedding_layer(tabs, [ids[:, f] for f in range(m)], biases)
    fm = oracle.fm_layer(embeds, bias).reshape(-1)
    rows = torch.from_numpy(oracle.global_rows(ids.numpy(), offs))
    dn = d["dense"].double()
    lin = SC[:, 1][rows].sum(1) + dn @ LD
    x = oracle.dnn_combiner([embeds, dn])
    y1 = x @ W1 + b1
    y1.retain_grad()
    act = oracle.leaky_relu_tf
    dnn = (act(act(y1) @ W2 + b2) @ w3 + b3).reshape(-1)
    logit = (lin + w0) + fm + dnn
    logit.retain_grad()
    pred = oracle.prediction(logit.reshape(-1, 1), "classification")
    loss = oracle.create_loss(d["y"].double(), pred, "classification")
    loss.backward()
    out = dict(y1=y1, fm=fm, lin=lin, S=embeds.sum(1), logit=logit, loss=loss.reshape(1), g1=y1.grad, g=logit.grad,
               d_table=T.grad, d_scal=SC.grad, dW1=W1.grad, db1=b1.grad, dW2=W2.grad, db2=b2.grad, dw3=w3.grad.reshape(-1),
               db3=b3.grad, dw0=w0.grad, dlin_dense=LD.grad,
               table_gd=oracle.fresh_optimizer_step(T.detach(), T.grad, "gd", LR),
               scal_gd=oracle.fresh_optimizer_step(SC.detach(), SC.grad, "gd", LR))
    return {k_: v.detach().numpy() for k_, v in out.items()}


def main():
    d = inputs()
    out = {f"in_{k_}": v.numpy() for k_, v in d.items()}
    out.update({f"out_{k_}": v for k_, v in expected(d).items()})
    out["meta"] = np.array([K, B, ND, N1, N2], dtype=np.int64)
    out["sizes"] = np.array(SIZES, dtype=np.int64)
    out["lr"] = np.array([LR])
    path = os.path.join(HERE, "tower_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
