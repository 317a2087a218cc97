"""K5 parity: CIN layer forward/backward (CUDA-core fp32 path and tcgen05 paths) and the xDeepFM model vs the oracle."""
import numpy as np
import pytest
import torch

import oracle
from tests import parity_util as pu

pytestmark = pytest.mark.gpu

ACTS = {0: (lambda t: t), 1: oracle.relu, 2: oracle.leaky_relu_tf}


def _layer_oracle(x0, xk, W, bias, act):
    """One CIN layer in the oracle's op order: returns act(Z.W + b) as [B,N,D] (layers.py:711-739)."""
    B, m, D = x0.shape
    H = xk.shape[1]
    z = torch.einsum("bpd,bqd->bdpq", x0, xk).reshape(B, D, m * H)
    f = z @ W + bias
    return ACTS[act](f).permute(0, 2, 1), f.permute(0, 2, 1)


CASES = [  # B, m, H, D, N
    (3, 2, 2, 4, 16),
    (37, 6, 6, 8, 20),
    (64, 26, 26, 16, 200),
    (40, 26, 100, 16, 200),
    (9, 5, 7, 12, 10),
    (5, 3, 4, 64, 6),
    (130, 4, 3, 1, 8),
]


def _make(B, m, H, D, N, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = (torch.randn(B, m, D, generator=g) * 0.5).double().requires_grad_()
    xk_full = (torch.randn(B, 2 * H, D, generator=g) * 0.5).double()  # xk is the first half of a wider tensor
    xk = xk_full[:, :H].clone().requires_grad_()
    W = (torch.randn(m * H, N, generator=g) / np.sqrt(m * H)).double().requires_grad_()
    bias = (torch.randn(N, generator=g) * 0.1).double().requires_grad_()
    dout = torch.randn(B, N, D, generator=g).double()
    return x0, xk_full, xk, W, bias, dout


@pytest.mark.parametrize("B,m,H,D,N", CASES)
@pytest.mark.parametrize("act", [0, 2])
def test_cin_layer_simt_fwd_bwd(B, m, H, D, N, act):
    from recman_b200 import ops

    x0, xk_full, xk, W, bias, dout = _make(B, m, H, D, N, B * 7 + H)
    out64, pre64 = _layer_oracle(x0, xk, W, bias, act)
    out64.backward(dout)
    xk_dev = xk_full.float().cuda()[:, :H]  # batch stride 2*H*D
    x0d, Wd, bd = x0.detach().float().cuda(), W.detach().float().cuda(), bias.detach().float().cuda()
    out, pre = ops.cin_layer_fwd(x0d, xk_dev, Wd, bd, act, 0)
    scale = max(1.0, float(pre64.abs().max()))
    torch.testing.assert_close(pre.cpu().double(), pre64.detach(), rtol=1e-5, atol=2e-6 * scale)
    torch.testing.assert_close(out.cpu().double(), out64.detach(), rtol=1e-5, atol=2e-6 * scale)
    dx0 = torch.ones(B, m, D, device="cuda")  # accumulated into
    dxk_full = torch.zeros(B, 2 * H, D, device="cuda")
    dW, dbias = ops.cin_layer_bwd(x0d, xk_dev, Wd, pre, dout.float().cuda(), act, 0, dx0, dxk_full[:, :H])
    for name, got, exp in [("dW", dW, W.grad), ("dbias", dbias, bias.grad), ("dx0", dx0 - 1, x0.grad),
                           ("dxk", dxk_full[:, :H], xk.grad)]:
        e = exp.double()
        torch.testing.assert_close(got.cpu().double(), e, rtol=1e-5, atol=1e-5 * float(e.abs().max()),
                                   msg=lambda m_: f"{name}: {m_}")
    assert torch.all(dxk_full[:, H:] == 0)
    # determinism
    dx0b = torch.ones(B, m, D, device="cuda")
    dxkb = torch.zeros(B, 2 * H, D, device="cuda")
    dW2, dbias2 = ops.cin_layer_bwd(x0d, xk_dev, Wd, pre, dout.float().cuda(), act, 0, dx0b, dxkb[:, :H])
    assert torch.equal(dW, dW2) and torch.equal(dbias, dbias2) and torch.equal(dx0, dx0b) and torch.equal(dxk_full, dxkb)


def test_cin_notebook_kat_on_gpu():
    """notes/xDeepFM.ipynb cell 6 through the CUDA kernels (all-ones filters, identity activation)."""
    from recman_b200 import ops

    x = torch.tensor([[[1, 2, 3, 4], [5, 6, 7, 8]]], dtype=torch.float32, device="cuda")
    z16 = torch.zeros(16, device="cuda")
    f0, _ = ops.cin_layer_fwd(x, x, torch.ones(4, 16, device="cuda"), z16, 0, 0)
    nxt, direct0 = f0[:, :8], f0[:, 8:]
    f1, _ = ops.cin_layer_fwd(x, nxt, torch.ones(16, 16, device="cuda"), z16, 0, 0)
    pooled = torch.cat([direct0.sum(-1), f1.sum(-1)], dim=1).cpu().reshape(-1).tolist()
    assert pooled == [344.0] * 8 + [27648.0] * 16


@pytest.mark.parametrize("precision", ["fp32"])
def test_xdeepfm_parity(precision):
    from recman_b200.th import xDeepFM
    from recman_b200.th.layers import leaky_relu

    fd = pu.make_feat_dict([50, 7, 1000, 3, 200, 31, 2, 90], n_dense=13)
    X, y = pu.synth_batch(fd, 256, seed=11)
    hp = dict(embedding_size=16, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), cin_cross_layer_units=[20, 20, 12],
              cin_dropout=[1, 1, 1, 1], cin_precision=precision, deep_activation=leaky_relu, cin_activation=leaky_relu,
              learning_rate=0.01)
    model = xDeepFM(fd, hp, batch_size=256)
    rep = pu.compare(model, X, y)
    assert any(k.startswith("grad:cin_filter_") for k in rep)


def test_xdeepfm_toy_frame_with_multival():
    """The 16-row toy frame shape of examples/xDeepFM_test.py:24-45: 3 SparseFeat + 1 MultiValCsvFeat (a-d)."""
    from recman_b200.th import xDeepFM
    from recman_b200.th.layers import leaky_relu

    fd = pu.make_feat_dict([7, 7, 4], n_dense=0, multi_tags=["a", "b", "c", "d"])
    X, y = pu.synth_batch(fd, 16, seed=12)
    hp = dict(embedding_size=8, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), cin_cross_layer_units=[10, 10, 10],
              cin_dropout=[1, 1, 1, 1], cin_precision="fp32", deep_activation=leaky_relu, cin_activation=leaky_relu)
    model = xDeepFM(fd, hp, batch_size=128)
    pu.compare(model, X, y)


# ----------------------------------------------------------------------------------------------- tcgen05 path
TC_CASES = [  # B, m, H, D, N
    (16, 26, 26, 16, 200),   # one 256-row tile, layer 0 of C3
    (40, 26, 100, 16, 200),  # 640 rows: 3 tiles, last one partial
    (3, 2, 2, 4, 16),        # the notebook shape: 12 rows, tiny K (one partial stage)
    (37, 6, 6, 8, 20),       # C1-like: m=6 -> MPAD 8, N=20 -> NPAD 32
    (9, 5, 7, 12, 10),       # D does not divide the tile
    (200, 8, 50, 16, 256),   # widest N
    (64, 32, 3, 16, 100),    # m = 32 (largest supported)
    (2048, 26, 100, 16, 200),  # the C3 layer shape (K = 2600: six k'' splits), 128 tiles
    (33, 7, 90, 8, 24),      # MPAD 8, 720 k'' -> two splits, the second one ragged; partial tile
]


@pytest.mark.parametrize("B,m,H,D,N", TC_CASES)
# 3xTF32 (parity mode) meets north_star's 1e-5 at every K: the tensor core accumulates in fp32 with round-toward-zero,
# a bias that grows linearly with the number of accumulate steps (3.3e-5 of max|F| at K = 2600 in one accumulation), so
# the kernel accumulates at most 512 k'' per CTA in TMEM and adds the splits in fp32 round-to-nearest (cin_tc.cu).
@pytest.mark.parametrize("precision,rtol,atol", [(1, 1e-5, 1e-5), (2, 5e-3, 3e-3)])
@pytest.mark.parametrize("act", [2])
def test_cin_layer_tcgen05_fwd(B, m, H, D, N, precision, rtol, atol, act):
    """tensor-core forward vs the fp64 oracle: 3xTF32 within 1e-5, single-pass TF32 within ~1e-3."""
    from recman_b200 import ops

    x0, xk_full, xk, W, bias, dout = _make(B, m, H, D, N, B * 3 + N)
    out64, pre64 = _layer_oracle(x0, xk, W, bias, act)
    xk_dev = xk_full.float().cuda()[:, :H]
    x0d, Wd, bd = x0.detach().float().cuda(), W.detach().float().cuda(), bias.detach().float().cuda()
    out, pre = ops.cin_layer_fwd(x0d, xk_dev, Wd, bd, act, precision)
    torch.cuda.synchronize()
    assert ops.cin_tc_status() == 0, "a pipeline wait timed out inside the tcgen05 kernel"
    scale = max(1.0, float(pre64.abs().max()))
    torch.testing.assert_close(pre.cpu().double(), pre64.detach(), rtol=rtol, atol=atol * scale)
    torch.testing.assert_close(out.cpu().double(), out64.detach(), rtol=rtol, atol=atol * scale)
    # against the CUDA-core fp32 path of the same library
    out_s, pre_s = ops.cin_layer_fwd(x0d, xk_dev, Wd, bd, act, 0)
    torch.testing.assert_close(pre, pre_s, rtol=rtol, atol=atol * scale)
    # deterministic
    out2, _ = ops.cin_layer_fwd(x0d, xk_dev, Wd, bd, act, precision)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,m,H,D,N", [c for c in TC_CASES if c[3] % 4 == 0])
@pytest.mark.parametrize("precision,tol", [(1, 1e-5), (2, 4e-3)])
@pytest.mark.parametrize("act", [2])
def test_cin_layer_tcgen05_bwd(B, m, H, D, N, precision, tol, act):
    """tensor-core backward (dZ GEMM with in-TMEM contraction into dx0/dxk; slabbed dW GEMM) vs the fp64 oracle."""
    from recman_b200 import ops

    x0, xk_full, xk, W, bias, dout = _make(B, m, H, D, N, B * 5 + N)
    out64, pre64 = _layer_oracle(x0, xk, W, bias, act)
    xk_dev = xk_full.float().cuda()[:, :H]
    x0d, Wd, bd = x0.detach().float().cuda(), W.detach().float().cuda(), bias.detach().float().cuda()
    _, pre = ops.cin_layer_fwd(x0d, xk_dev, Wd, bd, act, 0)
    # The activation derivative is discontinuous at 0: an element whose fp32 pre-activation rounds to the other side
    # of the kink than the fp64 one (expected ~B*N*D*1e-7 of them) flips a whole dF entry.  The backward is handed
    # `pre`, so the reference uses the mask of that same `pre` on the fp64 linear part.
    slope = torch.where(pre.cpu().double() > 0, 1.0, 0.2) if act == 2 else torch.ones_like(pre64)
    pre64.backward(dout * slope)
    doutd = dout.float().cuda()

    def run():
        dx0 = torch.ones(B, m, D, device="cuda")
        dxk_full = torch.zeros(B, 2 * H, D, device="cuda")
        dW, dbias = ops.cin_layer_bwd(x0d, xk_dev, Wd, pre, doutd, act, precision, dx0, dxk_full[:, :H])
        torch.cuda.synchronize()
        assert ops.cin_tc_status() == 0, "a pipeline wait timed out inside the tcgen05 backward"
        return dW, dbias, dx0, dxk_full

    dW, dbias, dx0, dxk_full = run()
    for name, got, exp in [("dW", dW, W.grad), ("dbias", dbias, bias.grad), ("dx0", dx0 - 1, x0.grad),
                           ("dxk", dxk_full[:, :H], xk.grad)]:
        e = exp.double()
        torch.testing.assert_close(got.cpu().double(), e, rtol=tol, atol=tol * float(e.abs().max()),
                                   msg=lambda m_: f"{name}: {m_}")
    assert torch.all(dxk_full[:, H:] == 0)
    dW2, dbias2, dx0b, dxkb = run()  # deterministic
    assert torch.equal(dW, dW2) and torch.equal(dbias, dbias2) and torch.equal(dx0, dx0b) and torch.equal(dxk_full, dxkb)


def test_cin_tcgen05_notebook_kat():
    from recman_b200 import ops

    x = torch.tensor([[[1, 2, 3, 4], [5, 6, 7, 8]]], dtype=torch.float32, device="cuda")
    z16 = torch.zeros(16, device="cuda")
    f0, _ = ops.cin_layer_fwd(x, x, torch.ones(4, 16, device="cuda"), z16, 0, 1)
    nxt, direct0 = f0[:, :8], f0[:, 8:]
    f1, _ = ops.cin_layer_fwd(x, nxt, torch.ones(16, 16, device="cuda"), z16, 0, 1)
    pooled = torch.cat([direct0.sum(-1), f1.sum(-1)], dim=1).cpu().reshape(-1).tolist()
    assert pooled == [344.0] * 8 + [27648.0] * 16  # small integers are exact in tf32 hi/lo arithmetic


def test_xdeepfm_parity_3xtf32():
    from recman_b200.th import xDeepFM
    from recman_b200.th.layers import leaky_relu

    fd = pu.make_feat_dict([50, 7, 1000, 3, 200, 31, 2, 90], n_dense=13)
    X, y = pu.synth_batch(fd, 256, seed=11)
    hp = dict(embedding_size=16, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), cin_cross_layer_units=[20, 20, 12],
              cin_dropout=[1, 1, 1, 1], cin_precision="3xtf32", deep_activation=leaky_relu, cin_activation=leaky_relu,
              learning_rate=0.01)
    model = xDeepFM(fd, hp, batch_size=256)
    pu.compare(model, X, y)


@pytest.mark.parametrize("n0", [0, 10])
@pytest.mark.parametrize("precision", [0, 1])
def test_cin_layer_with_split_half_and_sum_pool(n0, precision):
    """CINLayerPoolFunction (layer + split-half + sum-pool over D, layers.py:738-751) == the layer followed by
    autograd's slicing and sum: same outputs, same gradients of x0 / xk / W / bias."""
    from recman_b200.autograd import CINLayerFunction, CINLayerPoolFunction

    B, m, H, D, N = 37, 6, 9, 8, 20
    g = torch.Generator().manual_seed(3)
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.5).cuda().requires_grad_()
    x0a, xka, Wa, ba = mk(B, m, D), mk(B, H, D), mk(m * H, N), mk(N)
    x0b, xkb, Wb, bb = [t.detach().clone().requires_grad_() for t in (x0a, xka, Wa, ba)]
    gn = torch.randn(B, n0, D, generator=g).cuda()
    gp = torch.randn(B, N - n0, generator=g).cuda()
    nxt, pooled = CINLayerPoolFunction.apply(x0a, xka, Wa, ba, 2, precision, n0)
    ((nxt * gn).sum() + (pooled * gp).sum()).backward()
    out = CINLayerFunction.apply(x0b, xkb, Wb, bb, 2, precision)
    nxt_r, pooled_r = out[:, :n0], out[:, n0:].sum(dim=-1)
    ((nxt_r * gn).sum() + (pooled_r * gp).sum()).backward()
    assert torch.equal(nxt, nxt_r)
    torch.testing.assert_close(pooled, pooled_r, rtol=1e-6, atol=1e-6)
    for a, b_, name in [(x0a, x0b, "dx0"), (xka, xkb, "dxk"), (Wa, Wb, "dW"), (ba, bb, "dbias")]:
        torch.testing.assert_close(a.grad, b_.grad, rtol=1e-6, atol=1e-6 * float(b_.grad.abs().max()),
                                   msg=lambda m_: f"{name}: {m_}")
    # only the pooled branch has a gradient (the last layer, or a detached next layer)
    x0c, xkc, Wc, bc = [t.detach().clone().requires_grad_() for t in (x0a, xka, Wa, ba)]
    _, pooled_c = CINLayerPoolFunction.apply(x0c, xkc, Wc, bc, 2, precision, n0)
    (pooled_c * gp).sum().backward()
    x0d, xkd, Wd, bd = [t.detach().clone().requires_grad_() for t in (x0a, xka, Wa, ba)]
    (CINLayerFunction.apply(x0d, xkd, Wd, bd, 2, precision)[:, n0:].sum(dim=-1) * gp).sum().backward()
    torch.testing.assert_close(Wc.grad, Wd.grad, rtol=1e-6, atol=1e-6 * float(Wd.grad.abs().max()))
    torch.testing.assert_close(xkc.grad, xkd.grad, rtol=1e-6, atol=1e-6 * float(xkd.grad.abs().max()))
