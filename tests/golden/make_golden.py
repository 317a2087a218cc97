"""Generate the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference (dev-wei/recman) cannot be executed here (TensorFlow absent, recman.th is an empty stub, CrossNet is
missing) and stores no outputs, so - apart from the notebook known-answer test (cin_notebook_kat.json) - fixtures
can only be produced by the oracle itself.  They freeze the oracle's behaviour (any later edit to oracle/ that
changes results fails tests/test_golden.py) and give the GPU tests seeded inputs with fp64 expected outputs that do
not depend on the oracle code at test time.

    python tests/golden/make_golden.py        # rewrites tests/golden/hotpath_v1.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def main():
    g = torch.Generator().manual_seed(2019)
    out = {}
    # ---- A1-A3: multi-table gather -------------------------------------------------------------
    sizes = [7, 1, 50, 13]
    k = 8
    B = 21
    tabs = [torch.randn(v, k, generator=g) for v in sizes]
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in sizes], 1)
    e, _ = oracle.feat_embedding_layer(tabs, [ids[:, f] for f in range(len(sizes))])
    out["gather_table"] = torch.cat(tabs).numpy()
    out["gather_offsets"] = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    out["gather_ids"] = ids.numpy()
    out["gather_out"] = e.numpy()
    # ---- A2: sqrtn pooled ---------------------------------------------------------------------
    counts = torch.tensor([2, 0, 5, 1, 3])
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)])
    values = torch.randint(0, 13, (int(counts.sum()),), generator=g)
    out["pooled_values"], out["pooled_offsets"] = values.numpy(), offsets.numpy()
    out["pooled_out"] = oracle.embedding_lookup_sqrtn(tabs[3], values, offsets)[:, 0].numpy()
    # ---- A4: segment sum ------------------------------------------------------------------------
    keys = oracle.global_rows(ids.numpy(), out["gather_offsets"]).reshape(-1)
    grad = torch.randn(B * len(sizes), k, generator=g).numpy()
    uniq, sums, order, seg = oracle.segment_sum_sorted(keys, grad)
    out["seg_grad"], out["seg_uniq"], out["seg_sums"], out["seg_order"], out["seg_start"] = grad, uniq, sums, order, seg
    # ---- A5: FM ---------------------------------------------------------------------------------
    e64 = (torch.randn(9, 5, 8, generator=g) * 0.3).double().requires_grad_()
    b64 = torch.randn(9, 5, 1, generator=g).double().requires_grad_()
    y = oracle.fm_layer(e64, b64)
    gy = torch.randn(9, 1, generator=g).double()
    y.backward(gy)
    out.update(fm_e=e64.detach().numpy(), fm_bias=b64.detach().numpy(), fm_out=y.detach().numpy(), fm_gout=gy.numpy(),
               fm_de=e64.grad.numpy(), fm_dbias=b64.grad.numpy())
    # ---- A6: cross ------------------------------------------------------------------------------
    d, L, Bc = 37, 3, 6
    x = (torch.randn(Bc, d, generator=g) * 0.5).double().requires_grad_()
    w = (torch.randn(L, d, generator=g) / np.sqrt(d)).double().requires_grad_()
    bb = (torch.randn(L, d, generator=g) * 0.1).double().requires_grad_()
    wo = (torch.randn(d, 1, generator=g) / np.sqrt(d)).double().requires_grad_()
    w0 = torch.randn(1, generator=g).double().requires_grad_()
    yc = oracle.cross_net(x, w, bb, wo, w0)
    gc = torch.randn(Bc, 1, generator=g).double()
    yc.backward(gc)
    out.update(cross_x=x.detach().numpy(), cross_w=w.detach().numpy(), cross_b=bb.detach().numpy(),
               cross_wo=wo.detach().numpy(), cross_w0=w0.detach().numpy(), cross_out=yc.detach().numpy(),
               cross_gout=gc.numpy(), cross_dx=x.grad.numpy(), cross_dw=w.grad.numpy(), cross_db=bb.grad.numpy(),
               cross_dwo=wo.grad.numpy(), cross_dw0=w0.grad.numpy())
    # ---- A7: CIN (full layer stack, leaky_relu 0.2, split-half) -----------------------------------
    m, D, units, Bn = 4, 8, (6, 4, 6), 5
    xc = (torch.randn(Bn, m, D, generator=g) * 0.5).double().requires_grad_()
    shapes, final = oracle.cin_layer_shapes(m, units)
    filt = [(torch.randn(*s, generator=g) / np.sqrt(s[1])).double().requires_grad_() for s in shapes]
    fb = [(torch.randn(s[-1], generator=g) * 0.1).double().requires_grad_() for s in shapes]
    cw = torch.randn(final, 1, generator=g).double().requires_grad_()
    cw0 = torch.randn(1, generator=g).double().requires_grad_()
    yn = oracle.cin(xc, filt, fb, cw, cw0)
    gn = torch.randn(Bn, 1, generator=g).double()
    yn.backward(gn)
    out.update(cin_x=xc.detach().numpy(), cin_w=cw.detach().numpy(), cin_w0=cw0.detach().numpy(), cin_out=yn.detach().numpy(),
               cin_gout=gn.numpy(), cin_dx=xc.grad.numpy(), cin_dw=cw.grad.numpy())
    for i in range(len(units)):
        out[f"cin_filter_{i}"] = filt[i].detach().numpy()
        out[f"cin_bias_{i}"] = fb[i].detach().numpy()
        out[f"cin_dfilter_{i}"] = filt[i].grad.numpy()
        out[f"cin_dbias_{i}"] = fb[i].grad.numpy()
    # ---- loss + fresh optimizers ------------------------------------------------------------------
    yt = (torch.rand(12, generator=g) < 0.4).double()
    pp = torch.rand(12, generator=g).double()
    pp[0], pp[1] = 0.0, 1.0
    out.update(bce_y=yt.numpy(), bce_p=pp.numpy(), bce_out=oracle.binary_crossentropy(yt, pp).numpy())
    p0 = torch.randn(10, generator=g).double()
    g0 = torch.randn(10, generator=g).double()
    g0[3] = 0.0
    out.update(opt_p=p0.numpy(), opt_g=g0.numpy())
    for name in ("adam", "adagrad", "gd"):
        out[f"opt_{name}"] = oracle.fresh_optimizer_step(p0, g0, name, 0.01).numpy()
    np.savez_compressed(os.path.join(HERE, "hotpath_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_v1.npz"), sum(v.nbytes for v in out.values()), "bytes")


if __name__ == "__main__":
    main()
