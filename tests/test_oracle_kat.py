"""Pin the oracle: known-answer test from the reference notebook + brute-force identities.

The reference stores no outputs (notes/xDeepFM.ipynb cell 6 has inputs and all-ones
filters only) so the KAT values are re-derived here with an independent nested-loop
numpy CIN and compared with tests/golden/cin_notebook_kat.json.
"""
import itertools
import json
import math
import os

import numpy as np
import pytest
import torch

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def naive_cin_pooled(x0, filters, biases, act):
    """5-nested-loop CIN returning the pooled [B, sum H] vector (float64 numpy)."""
    B, m, D = x0.shape
    xk = x0.copy()
    outs = []
    L = len(filters)
    for i, W in enumerate(filters):
        H = xk.shape[1]
        N = W.shape[1]
        fm = np.zeros((B, N, D))
        for b in range(B):
            for d in range(D):
                for n in range(N):
                    acc = 0.0
                    for p in range(m):
                        for q in range(H):
                            acc += x0[b, p, d] * xk[b, q, d] * W[p * H + q, n]
                    fm[b, n, d] = act(acc + biases[i][n])
        if i != L - 1:
            xk, direct = fm[:, : N // 2], fm[:, N // 2 :]
        else:
            direct = fm
        outs.append(direct)
    return np.concatenate(outs, axis=1).sum(-1)


def test_cin_notebook_kat():
    # notes/xDeepFM.ipynb cell 6: m=2, D=4, units (16,16), all-ones filters, no bias/activation
    x = torch.tensor([[[1, 2, 3, 4], [5, 6, 7, 8]]], dtype=torch.float64)
    f0 = torch.ones(1, 4, 16, dtype=torch.float64)
    f1 = torch.ones(1, 16, 16, dtype=torch.float64)
    z = torch.zeros(16, dtype=torch.float64)
    pooled = oracle.cin(x, [f0, f1], [z, z], None, None, activation=lambda t: t, return_pooled=True)
    naive = naive_cin_pooled(x.numpy(), [f0[0].numpy(), f1[0].numpy()], [z.numpy(), z.numpy()], lambda t: t)
    np.testing.assert_allclose(pooled.numpy(), naive, rtol=0, atol=0)
    with open(os.path.join(GOLDEN, "cin_notebook_kat.json")) as f:
        kat = json.load(f)
    assert pooled.numpy().reshape(-1).tolist() == kat["pooled"]
    # layer maps: (x_age + x_occ)^2 per d ; layer 1: 8 * layer0 * (x_age + x_occ)
    s = np.array([6.0, 8.0, 10.0, 12.0])
    assert kat["layer0_row"] == (s**2).tolist()
    assert kat["layer1_row"] == (8 * s**3).tolist()
    assert kat["pooled"] == [float((s**2).sum())] * 8 + [float((8 * s**3).sum())] * 16


@pytest.mark.parametrize("B,m,D,units", [(3, 4, 5, (6, 4)), (2, 3, 2, (4, 6, 2)), (1, 2, 3, (2,))])
def test_cin_vs_naive(B, m, D, units):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, m, D, generator=g, dtype=torch.float64)
    shapes, final = oracle.cin_layer_shapes(m, units)
    filters = [torch.randn(*s, generator=g, dtype=torch.float64) for s in shapes]
    biases = [torch.randn(s[-1], generator=g, dtype=torch.float64) for s in shapes]
    cin_w = torch.randn(final, 1, generator=g, dtype=torch.float64)
    cin_w0 = torch.randn(1, generator=g, dtype=torch.float64)
    pooled = oracle.cin(x, filters, biases, cin_w, cin_w0, return_pooled=True)
    lrelu = lambda t: max(0.2 * t, t)
    naive = naive_cin_pooled(x.numpy(), [f[0].numpy() for f in filters], [b.numpy() for b in biases], lrelu)
    np.testing.assert_allclose(pooled.numpy(), naive, rtol=1e-12, atol=1e-12)
    out = oracle.cin(x, filters, biases, cin_w, cin_w0)
    np.testing.assert_allclose(out.numpy(), naive @ cin_w.numpy() + cin_w0.numpy(), rtol=1e-12, atol=1e-12)
    assert out.shape == (B, 1)


def test_cin_shapes_default_units():
    shapes, final = oracle.cin_layer_shapes(26, (200, 200, 200))
    assert shapes == [(1, 676, 200), (1, 2600, 200), (1, 2600, 200)]
    assert final == 400


def test_fm_vs_pairwise():
    g = torch.Generator().manual_seed(1)
    e = torch.randn(5, 7, 3, generator=g, dtype=torch.float64)
    b = torch.randn(5, 7, 1, generator=g, dtype=torch.float64)
    y = oracle.fm_layer(e, b)
    ref = torch.zeros(5, 1, dtype=torch.float64)
    for n in range(5):
        acc = b[n].sum()
        for i, j in itertools.combinations(range(7), 2):
            acc = acc + (e[n, i] * e[n, j]).sum()
        ref[n, 0] = acc
    torch.testing.assert_close(y, ref, rtol=1e-12, atol=1e-12)


def test_cross_vs_matrix_form():
    g = torch.Generator().manual_seed(2)
    B, d, L = 4, 6, 3
    x = torch.randn(B, d, generator=g, dtype=torch.float64)
    w = torch.randn(L, d, generator=g, dtype=torch.float64)
    bb = torch.randn(L, d, generator=g, dtype=torch.float64)
    wo = torch.randn(d, 1, generator=g, dtype=torch.float64)
    w0 = torch.randn(1, generator=g, dtype=torch.float64)
    y = oracle.cross_net(x, w, bb, wo, w0)
    for n in range(B):
        x0 = x[n].reshape(d, 1)
        xl = x0.clone()
        for l in range(L):
            xl = (x0 @ xl.T) @ w[l].reshape(d, 1) + bb[l].reshape(d, 1) + xl  # x0 x_l^T w
        torch.testing.assert_close(y[n], (xl.T @ wo).reshape(1) + w0, rtol=1e-12, atol=1e-12)


def test_sqrtn_lookup_and_autograd_variant():
    g = torch.Generator().manual_seed(3)
    table = torch.randn(9, 4, generator=g, dtype=torch.float64)
    values = torch.tensor([1, 2, 4, 1, 0, 3, 3], dtype=torch.int64)
    offsets = torch.tensor([0, 3, 3, 4, 7], dtype=torch.int64)  # n = 3, 0, 1, 3
    out = oracle.embedding_lookup_sqrtn(table, values, offsets)
    assert out.shape == (4, 1, 4)
    torch.testing.assert_close(out[0, 0], (table[1] + table[2] + table[4]) / math.sqrt(3))
    assert torch.equal(out[1, 0], torch.zeros(4, dtype=torch.float64))
    torch.testing.assert_close(out[2, 0], table[1])
    torch.testing.assert_close(out[3, 0], (table[0] + 2 * table[3]) / math.sqrt(3))
    from oracle.layers import _lookup_sqrtn_autograd

    torch.testing.assert_close(_lookup_sqrtn_autograd(table, values, offsets), out)


def test_embedding_layer_concat_order():
    t0 = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    t1 = 100 + torch.arange(6, dtype=torch.float32).reshape(2, 3)
    ids0 = torch.tensor([[3], [0]])
    ids1 = torch.tensor([1, 1])
    e, b = oracle.feat_embedding_layer([t0, t1], [ids0, ids1], [t0[:, :1], t1[:, :1]])
    assert e.shape == (2, 2, 3) and b.shape == (2, 2, 1)
    assert torch.equal(e[0, 0], t0[3]) and torch.equal(e[0, 1], t1[1]) and torch.equal(e[1, 0], t0[0])
    assert torch.equal(b[:, :, 0], torch.tensor([[9.0, 103.0], [0.0, 103.0]]))


def test_sparse_linear_matches_gather_form():
    g = torch.Generator().manual_seed(4)
    sizes = [5, 4, 1, 3]
    kinds = ["sparse", "multi", "dense", "sparse"]
    W = torch.randn(sum(sizes), 1, generator=g, dtype=torch.float64)
    w0 = torch.randn(1, generator=g, dtype=torch.float64)
    ids_a = torch.tensor([4, 0, 2])
    mv = (torch.tensor([1, 2, 0, 3, 3]), torch.tensor([0, 2, 3, 5]))
    dense = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64)
    ids_b = torch.tensor([0, 1, 2])
    y = oracle.sparse_linear(W, w0, sizes, [ids_a, mv, dense, ids_b], kinds)
    Wf = W.reshape(-1)
    exp = torch.stack(
        [
            Wf[4] + Wf[5 + 1] + Wf[5 + 2] + 0.5 * Wf[9] + Wf[10 + 0],
            Wf[0] + 0.0 + -1.0 * Wf[9] + Wf[10 + 1],  # tag id 0 (unknown) is zeroed
            Wf[2] + 2 * Wf[5 + 3] + 2.0 * Wf[9] + Wf[10 + 2],
        ]
    ) + w0
    torch.testing.assert_close(y.reshape(-1), exp)


def test_bce_keras_form():
    y = torch.tensor([1.0, 0.0, 1.0, 0.0])
    p = torch.tensor([0.9, 0.2, 0.0, 1.0])
    eps = 1e-7
    pc = np.clip(p.numpy().astype(np.float64), eps, 1 - eps)
    exp = -(y.numpy() * np.log(pc + eps) + (1 - y.numpy()) * np.log(1 - pc + eps)).mean()
    got = oracle.binary_crossentropy(y.double(), p.double())
    assert abs(got.item() - exp) < 1e-12
    assert oracle.binary_crossentropy(y, p).dtype == torch.float32


def test_leaky_relu_slope_and_l2():
    x = torch.tensor([-1.0, 2.0])
    assert torch.equal(oracle.leaky_relu_tf(x), torch.tensor([-0.2, 2.0]))
    assert oracle.l2_loss(torch.tensor([3.0, 4.0])).item() == 12.5


def test_dcn_sums_dnn_logit_twice():
    g = torch.Generator().manual_seed(5)
    B, m, k, nd = 3, 2, 2, 1
    d = m * k + nd
    e = torch.randn(B, m, k, generator=g, dtype=torch.float64)
    dense = torch.randn(B, nd, generator=g, dtype=torch.float64)
    dnn_p = ([torch.randn(d, 4, generator=g, dtype=torch.float64)], [torch.zeros(4, dtype=torch.float64)],
             torch.randn(4, 1, generator=g, dtype=torch.float64), torch.zeros(1, dtype=torch.float64))
    cr_p = (torch.randn(2, d, generator=g, dtype=torch.float64), torch.randn(2, d, generator=g, dtype=torch.float64),
            torch.randn(d, 1, generator=g, dtype=torch.float64), torch.zeros(1, dtype=torch.float64))
    x = oracle.dnn_combiner([e, dense])
    got = oracle.dcn_logit(e, None, dense, dnn_p, cr_p)
    exp = 2 * oracle.dnn(x, *dnn_p, activation=oracle.relu) + oracle.cross_net(x, *cr_p)
    torch.testing.assert_close(got, exp)


@pytest.mark.parametrize("name", ["fm", "cross", "cin", "dnn"])
def test_gradcheck_fp64(name):
    g = torch.Generator().manual_seed(11)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64, requires_grad=True)
    if name == "fm":
        assert torch.autograd.gradcheck(oracle.fm_layer, (rn(2, 3, 4), rn(2, 3, 1)))
    elif name == "cross":
        assert torch.autograd.gradcheck(oracle.cross_net, (rn(2, 5), rn(3, 5), rn(3, 5), rn(5, 1), rn(1)))
    elif name == "cin":
        shapes, final = oracle.cin_layer_shapes(3, (4, 2))
        f = [rn(*s) for s in shapes]
        b = [rn(s[-1]) for s in shapes]
        fn = lambda x, f0, f1, b0, b1, w, w0: oracle.cin(x, [f0, f1], [b0, b1], w, w0, activation=torch.tanh)
        assert torch.autograd.gradcheck(fn, (rn(2, 3, 2), f[0], f[1], b[0], b[1], rn(final, 1), rn(1)))
    else:
        fn = lambda x, w0, w1, b0, b1, w, c: oracle.dnn(x, [w0, w1], [b0, b1], w, c, activation=torch.tanh)
        assert torch.autograd.gradcheck(fn, (rn(2, 5), rn(5, 4), rn(4, 3), rn(4), rn(3), rn(3, 1), rn(1)))


def test_segment_sum_sorted():
    keys = np.array([5, 2, 5, 9, 2, 5], dtype=np.int64)
    grads = np.arange(12, dtype=np.float32).reshape(6, 2)
    uniq, sums, order, seg = oracle.segment_sum_sorted(keys, grads)
    assert uniq.tolist() == [2, 5, 9]
    assert order.tolist() == [1, 4, 0, 2, 5, 3]
    assert seg.tolist() == [0, 2, 5, 6]
    np.testing.assert_array_equal(sums, np.array([[10, 12], [14, 17], [6, 7]], dtype=np.float32))
    dense = oracle.dense_table_grad(keys, grads, 10)
    np.testing.assert_allclose(dense[uniq], sums)
    u, s, o, sg = oracle.segment_sum_sorted(np.zeros(0, np.int64), np.zeros((0, 2), np.float32))
    assert u.shape == (0,) and s.shape == (0, 2) and sg.tolist() == [0]


def test_fresh_adam_first_step_is_sign_like():
    p = torch.zeros(3, dtype=torch.float64)
    g = torch.tensor([0.5, -2.0, 0.0], dtype=torch.float64)
    new = oracle.fresh_optimizer_step(p, g, "adam", 0.01)
    assert abs(new[0].item() + 0.01) < 1e-6 and abs(new[1].item() - 0.01) < 1e-6 and new[2].item() == 0.0
    assert torch.equal(oracle.fresh_optimizer_step(p, g, "gd", 0.1), -0.1 * g)


def test_segment_sum_sorted_long_segments_are_chunked():
    """Hot rows: a segment longer than SEG_LONG positions is summed chunk-wise (SEG_CHUNK positions sequentially per
    partial, partials in chunk order) - the association the CUDA kernels commit to.  Short segments stay sequential."""
    from oracle.segment import SEG_CHUNK, SEG_LONG

    rng = np.random.RandomState(0)
    n_long = 3 * SEG_CHUNK + 5
    keys = np.concatenate([np.full(n_long, 7), np.full(SEG_LONG, 3), [9]]).astype(np.int64)
    perm = rng.permutation(len(keys))
    keys = keys[perm]
    g = (rng.randn(len(keys), 2) * 1e3).astype(np.float32)
    uniq, sums, order, seg = oracle.segment_sum_sorted(keys, g)
    assert uniq.tolist() == [3, 7, 9] and np.diff(seg).tolist() == [SEG_LONG, n_long, 1]
    gs = g[order]
    seq = np.zeros(2, np.float32)
    for row in gs[seg[0]:seg[1]]:
        seq = (seq + row).astype(np.float32)
    assert np.array_equal(sums[0], seq)  # exactly SEG_LONG positions: still the plain sequential sum
    total = np.zeros(2, np.float32)
    for c0 in range(int(seg[1]), int(seg[2]), SEG_CHUNK):
        part = np.zeros(2, np.float32)
        for row in gs[c0:min(c0 + SEG_CHUNK, int(seg[2]))]:
            part = (part + row).astype(np.float32)
        total = (total + part).astype(np.float32)
    assert np.array_equal(sums[1], total)
    _, plain, _, _ = oracle.segment_sum_sorted(keys, g, chunked=False)
    np.testing.assert_allclose(sums, plain, rtol=1e-5, atol=1e-2)
    dense = oracle.dense_table_grad(keys, g, 10)
    np.testing.assert_allclose(sums, dense[uniq], rtol=1e-5, atol=1e-2)
