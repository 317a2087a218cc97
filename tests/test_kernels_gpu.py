"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle.

Bit-exact for gathered rows, gradient row indices and the unfused segment sums;
fp32 within rtol 1e-5 (+ atol scaled to the data) for logits and gradients.
"""
import math

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _ops():
    from recman_b200 import ops

    return ops


def _tables(sizes, k, seed=0, scale=0.1):
    g = torch.Generator().manual_seed(seed)
    tabs = [torch.randn(v, k, generator=g) * scale for v in sizes]
    offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int64)
    return tabs, torch.cat(tabs, 0), offs


def _ids(sizes, B, seed=1, dup=False):
    g = torch.Generator().manual_seed(seed)
    cols = []
    for v in sizes:
        hi = min(v, 3) if dup else v
        cols.append(torch.randint(0, hi, (B,), generator=g))
    ids = torch.stack(cols, 1).contiguous()
    if B > 0:
        ids[0] = 0  # id 0 = the "unknown" row (tf/inputs.py:116-126)
        ids[-1] = torch.tensor([v - 1 for v in sizes])  # last row of every table
    return ids


def assert_close(got, exp, rtol=RTOL, atol_scale=1e-6):
    got = got.detach().cpu().double()
    exp = exp.detach().cpu().double()
    atol = atol_scale * max(1.0, float(exp.abs().max())) if exp.numel() else 0.0
    torch.testing.assert_close(got, exp, rtol=rtol, atol=atol)


@pytest.mark.parametrize("k", [1, 4, 8, 12, 16, 20, 64, 128, 192, 7])
@pytest.mark.parametrize("B", [1, 37, 1000])
def test_gather_bit_exact(k, B):
    ops = _ops()
    sizes = [1, 5, 944, 1683, 3, 22, 796]
    tabs, table, offs = _tables(sizes, k)
    ids = _ids(sizes, B)
    exp, _ = oracle.feat_embedding_layer(tabs, [ids[:, f] for f in range(len(sizes))])
    st = ops.new_status("cuda")
    got = ops.gather(table.cuda(), offs.cuda(), ids.cuda(), status=st)
    assert torch.equal(got.cpu(), exp)
    assert int(st.item()) == 0


def test_gather_into_strided_rows_and_oob_flag():
    ops = _ops()
    sizes = [10, 20]
    k = 16
    tabs, table, offs = _tables(sizes, k)
    ids = _ids(sizes, 50)
    out = torch.full((50, 2 * k + 4), -1.0, device="cuda")
    ops.gather(table.cuda(), offs.cuda(), ids.cuda(), out=out)
    exp, _ = oracle.feat_embedding_layer(tabs, [ids[:, 0], ids[:, 1]])
    assert torch.equal(out[:, : 2 * k].cpu(), exp.reshape(50, -1))
    assert torch.all(out[:, 2 * k :] == -1.0)
    bad = ids.clone()
    bad[3, 1] = 20  # one past the end of table 1
    bad[4, 0] = -1
    st = ops.new_status("cuda")
    got = ops.gather(table.cuda(), offs.cuda(), bad.cuda(), status=st).cpu()
    assert int(st.item()) != 0
    assert torch.all(got[3, 1] == 0) and torch.all(got[4, 0] == 0)
    assert torch.equal(got[5], exp[5])


def test_gather_empty_batch():
    ops = _ops()
    tabs, table, offs = _tables([4, 4], 8)
    got = ops.gather(table.cuda(), offs.cuda(), torch.zeros(0, 2, dtype=torch.int64, device="cuda"))
    assert got.shape == (0, 2, 8)


@pytest.mark.parametrize("k", [1, 8, 20])
def test_gather_pooled_sqrtn(k):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    table = torch.randn(30, k, generator=g)
    row_offset, rows = 9, 20  # the multi-val field's table sits in the middle of the big one
    counts = torch.tensor([3, 0, 1, 6, 2, 0, 4])
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)])
    values = torch.randint(0, rows, (int(counts.sum()),), generator=g)
    exp = oracle.embedding_lookup_sqrtn(table[row_offset : row_offset + rows], values, offsets)[:, 0]
    got = ops.gather_pooled(table.cuda(), row_offset, rows, values.cuda(), offsets.cuda())
    assert torch.equal(got.cpu(), exp)  # same add order, IEEE div/sqrt -> bit exact


@pytest.mark.parametrize("m,k", [(6, 8), (26, 16), (26, 64), (3, 12), (5, 7), (2, 128), (4, 1)])
@pytest.mark.parametrize("B", [1, 33, 513])
def test_fm_fwd_bwd(m, k, B):
    ops = _ops()
    g = torch.Generator().manual_seed(B + m)
    e = (torch.randn(B, m, k, generator=g) * 0.3).requires_grad_()
    bias = torch.randn(B, m, 1, generator=g).requires_grad_()
    y = oracle.fm_layer(e, bias)
    gout = torch.randn(B, 1, generator=g)
    y.backward(gout)
    y64 = oracle.fm_layer(e.detach().double(), bias.detach().double())
    got, S = ops.fm_fwd(e.detach().cuda(), bias.detach().cuda())
    # the fp32 oracle itself is only ~1e-6 from fp64: compare with both
    assert_close(got, y64.reshape(-1), rtol=RTOL, atol_scale=2e-6)
    assert_close(S, e.detach().sum(1))
    de, dbias = ops.fm_bwd(e.detach().cuda(), S, gout.cuda())
    assert_close(de, e.grad)
    assert_close(dbias, bias.grad.reshape(B, m))
    de2, _ = ops.fm_bwd(e.detach().cuda(), None, gout.cuda(), want_bias=False)  # S recomputed
    assert_close(de2, e.grad)
    acc = torch.ones(B, m, k, device="cuda")
    ops.fm_bwd(e.detach().cuda(), S, gout.cuda(), d_embeds=acc, accumulate=True)
    assert_close(acc, e.grad + 1)


@pytest.mark.parametrize("k,n_dense", [(8, 2), (16, 13), (64, 13), (12, 0), (128, 1)])
def test_fused_gather_fm_front_end(k, n_dense):
    ops = _ops()
    sizes = [7, 50, 3, 1000, 12, 2]
    m = len(sizes)
    B = 257
    tabs, table, offs = _tables(sizes, k)
    g = torch.Generator().manual_seed(9)
    bias_t = torch.randn(int(offs[-1]), generator=g)
    lin_t = torch.randn(int(offs[-1]), generator=g)
    ids = _ids(sizes, B)
    dense = torch.randn(B, n_dense, generator=g) if n_dense else None
    lin_dense = torch.randn(n_dense, generator=g) if n_dense else None
    e, _ = oracle.feat_embedding_layer(tabs, [ids[:, f] for f in range(m)])
    rows = ids + offs[:m]
    bias = bias_t[rows].unsqueeze(-1)
    fm_exp = oracle.fm_layer(e.double(), bias.double()).reshape(-1)
    lin_exp = lin_t[rows].double().sum(1)
    if n_dense:
        lin_exp = lin_exp + dense.double() @ lin_dense.double()
    st = ops.new_status("cuda")
    x, fm, lin, S = ops.gather_fm_fwd(
        table.cuda(), bias_t.cuda(), lin_t.cuda(), offs.cuda(), ids.cuda(),
        dense.cuda() if n_dense else None, lin_dense.cuda() if n_dense else None, status=st,
    )
    assert int(st.item()) == 0
    d = m * k + n_dense
    assert x.shape[1] % 4 == 0 and x.shape[1] >= d
    assert torch.equal(x[:, : m * k].cpu(), e.reshape(B, -1))  # gathered rows bit-exact
    if n_dense:
        assert torch.equal(x[:, m * k : d].cpu(), dense)
    assert_close(fm, fm_exp, atol_scale=2e-6)
    assert_close(lin, lin_exp, atol_scale=2e-6)
    assert_close(S, e.sum(1))
    # without the optional tables
    x2, fm2, lin2, _ = ops.gather_fm_fwd(table.cuda(), None, None, offs.cuda(), ids.cuda(), None, None)
    assert torch.equal(x2[:, : m * k].cpu(), e.reshape(B, -1))
    assert_close(fm2, oracle.fm_layer(e.double(), torch.zeros(B, m, 1, dtype=torch.float64)).reshape(-1), atol_scale=2e-6)
    assert torch.all(lin2 == 0)


@pytest.mark.parametrize("k", [1, 8, 16, 64, 12, 6, 192])
@pytest.mark.parametrize("dup", [False, True])
def test_segment_plan_and_reduce_bit_exact(k, dup):
    ops = _ops()
    sizes = [5, 1, 300, 40]
    m = len(sizes)
    B = 301
    offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int64)
    ids = _ids(sizes, B, seed=3, dup=dup)
    g = torch.Generator().manual_seed(4)
    grad = torch.randn(B, m, k, generator=g)
    keys = oracle.global_rows(ids.numpy(), offs.numpy()).reshape(-1)
    uniq, sums, order, seg = oracle.segment_sum_sorted(keys, grad.reshape(-1, k).numpy())
    plan = ops.segment_plan(ids.cuda(), offs.cuda(), int(offs[-1]))
    n = plan.num_unique()
    assert n == len(uniq)
    assert np.array_equal(plan.uniq_rows[:n].cpu().numpy(), uniq)  # gradient row indices: bit exact
    assert np.array_equal(plan.sorted_pos.cpu().numpy().astype(np.int64), order)  # stable sort
    assert np.array_equal(plan.seg_start[: n + 1].cpu().numpy().astype(np.int64), seg)
    rows = ops.segment_reduce(grad.cuda(), plan, k)
    assert np.array_equal(rows[:n].cpu().numpy(), sums)  # same add order -> bit exact
    # determinism: run twice, identical bits
    plan2 = ops.segment_plan(ids.cuda(), offs.cuda(), int(offs[-1]))
    rows2 = ops.segment_reduce(grad.cuda(), plan2, k)
    assert torch.equal(rows[:n], rows2[:n]) and torch.equal(plan.sorted_pos, plan2.sorted_pos)
    # against the float64 dense gradient
    dense = oracle.dense_table_grad(keys, grad.reshape(-1, k).numpy(), int(offs[-1]))
    sg = ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique)
    assert_close(sg.to_dense(int(offs[-1])), torch.from_numpy(dense), atol_scale=2e-6)


def test_segment_plan_no_offsets_and_empty():
    ops = _ops()
    keys = torch.tensor([7, 7, 0, 3, 7, 3], dtype=torch.int64)
    plan = ops.segment_plan(keys.cuda(), None, 8)
    n = plan.num_unique()
    assert plan.uniq_rows[:n].tolist() == [0, 3, 7]
    assert plan.seg_start[: n + 1].tolist() == [0, 1, 3, 6]
    assert plan.sorted_pos.tolist() == [2, 3, 5, 0, 1, 4]
    empty = ops.segment_plan(torch.zeros(0, 3, dtype=torch.int64, device="cuda"), None, 8)
    assert empty.num_unique() == 0


@pytest.mark.parametrize("k", [8, 16, 64])
def test_fused_embedding_backward(k):
    ops = _ops()
    sizes = [9, 2, 500, 31, 4]
    m = len(sizes)
    B = 400
    tabs, table, offs = _tables(sizes, k, scale=0.3)
    ids = _ids(sizes, B, seed=8, dup=False)
    ids[:, 1] = 1  # one very hot row: a 400-long segment
    g = torch.Generator().manual_seed(2)
    n_dense = 3
    dense = torch.randn(B, n_dense, generator=g)
    x, fm, lin, S = ops.gather_fm_fwd(table.cuda(), None, None, offs.cuda(), ids.cuda(), dense.cuda(), None)
    ld = x.shape[1]
    dx = torch.randn(B, ld, generator=g)
    g_fm = torch.randn(B, generator=g)
    g_lin = torch.randn(B, generator=g)
    # oracle: dE = dx + g_fm*(S - e), scatter-added
    e = x[:, : m * k].cpu().reshape(B, m, k).double()
    dE = dx[:, : m * k].reshape(B, m, k).double() + g_fm.double().reshape(B, 1, 1) * (e.sum(1, keepdim=True) - e)
    keys = oracle.global_rows(ids.numpy(), offs.numpy()).reshape(-1)
    total = int(offs[-1])
    exp_rows = oracle.dense_table_grad(keys, dE.reshape(-1, k).numpy(), total)
    exp_bias = oracle.dense_table_grad(keys, g_fm.double().repeat_interleave(m).reshape(-1, 1).numpy(), total)
    exp_lin = oracle.dense_table_grad(keys, g_lin.double().repeat_interleave(m).reshape(-1, 1).numpy(), total)
    plan = ops.segment_plan(ids.cuda(), offs.cuda(), total)
    rows, ob, ol = ops.emb_fm_bwd(dx.cuda(), x, ld, S, g_fm.cuda(), g_lin.cuda(), plan, k, True, True, True)
    assert_close(ops.SparseGrad(plan.uniq_rows, rows, plan.n_unique).to_dense(total), torch.from_numpy(exp_rows), atol_scale=5e-6)
    assert_close(ops.SparseGrad(plan.uniq_rows, ob, plan.n_unique).to_dense(total), torch.from_numpy(exp_bias), atol_scale=5e-6)
    assert_close(ops.SparseGrad(plan.uniq_rows, ol, plan.n_unique).to_dense(total), torch.from_numpy(exp_lin), atol_scale=5e-6)
    rows2, _, _ = ops.emb_fm_bwd(dx.cuda(), x, ld, S, g_fm.cuda(), g_lin.cuda(), plan, k, True, False, False)
    n = plan.num_unique()
    assert torch.equal(rows[:n], rows2[:n])  # deterministic
    # no FM, no DNN variants
    r3, _, _ = ops.emb_fm_bwd(dx.cuda(), None, ld, None, None, None, plan, k)
    exp3 = oracle.dense_table_grad(keys, dx[:, : m * k].reshape(-1, k).double().numpy(), total)
    assert_close(ops.SparseGrad(plan.uniq_rows, r3, plan.n_unique).to_dense(total), torch.from_numpy(exp3), atol_scale=5e-6)


@pytest.mark.parametrize("d,L", [(429, 6), (50, 3), (130, 1), (1000, 2), (1677, 2), (7, 0)])
@pytest.mark.parametrize("B", [1, 4096 // 8, 77])
def test_cross_fwd_bwd(d, L, B):
    ops = _ops()
    g = torch.Generator().manual_seed(d + L + B)
    x = (torch.randn(B, d, generator=g) * 0.5).requires_grad_()
    w = (torch.randn(L, d, generator=g) / math.sqrt(d)).requires_grad_()
    b = (torch.randn(L, d, generator=g) * 0.1).requires_grad_()
    wo = (torch.randn(d, 1, generator=g) / math.sqrt(d)).requires_grad_()
    w0 = torch.randn(1, generator=g).requires_grad_()
    params = [x, w, b, wo, w0]
    y = oracle.cross_net(*params)
    gout = torch.randn(B, 1, generator=g)
    y.backward(gout)
    p64 = [p.detach().double().requires_grad_() for p in params]
    y64 = oracle.cross_net(*p64)
    y64.backward(gout.double())
    ld = (d + 3) // 4 * 4 + 4
    xbuf = torch.zeros(B, ld, device="cuda")
    xbuf[:, :d] = x.detach().cuda()
    logit, dots = ops.cross_fwd(xbuf[:, :d], w.detach().cuda(), b.detach().cuda(), wo.detach().reshape(-1).cuda(), w0.detach().cuda())
    assert_close(logit, y64.reshape(-1), atol_scale=2e-6)
    dx, dw, db, dwo, dw0 = ops.cross_bwd(
        xbuf[:, :d], w.detach().cuda(), b.detach().cuda(), wo.detach().reshape(-1).cuda(), dots, gout.cuda()
    )
    for got, e64 in zip([dx, dw, db, dwo.reshape(d, 1), dw0], [p.grad for p in p64]):
        if e64 is not None:  # L == 0: w, b are unused
            assert_close(got, e64, atol_scale=5e-6)
    # accumulate into an existing dx (DCN: x feeds both towers)
    acc = torch.ones(B, d, device="cuda")
    ops.cross_bwd(xbuf[:, :d], w.detach().cuda(), b.detach().cuda(), wo.detach().reshape(-1).cuda(), dots, gout.cuda(), dx=acc, accumulate=True)
    assert_close(acc, p64[0].grad + 1, atol_scale=5e-6)
    # determinism of the batch-reduced parameter gradients
    again = ops.cross_bwd(xbuf[:, :d], w.detach().cuda(), b.detach().cuda(), wo.detach().reshape(-1).cuda(), dots, gout.cuda())
    assert torch.equal(again[1], dw) and torch.equal(again[2], db) and torch.equal(again[3], dwo)


@pytest.mark.parametrize("opt", ["adam", "adagrad", "gd"])
@pytest.mark.parametrize("k", [1, 16, 6])
def test_optimizer_steps(opt, k):
    ops = _ops()
    from recman_b200 import _C

    g = torch.Generator().manual_seed(0)
    table = torch.randn(50, k, generator=g)
    uniq = torch.tensor([3, 10, 11, 49], dtype=torch.int64)
    rows = torch.randn(8, k, generator=g)  # only the first 4 are valid
    n = torch.tensor([4], dtype=torch.int32)
    lr, l2 = 0.01, 0.001
    exp = table.clone().double()
    grad = rows[:4].double() + l2 * exp[uniq]
    exp[uniq] = oracle.fresh_optimizer_step(exp[uniq], grad, opt, lr)
    t = table.clone().cuda()
    pad = torch.cat([uniq, torch.zeros(4, dtype=torch.int64)])
    ops.sparse_opt_step(t, ops.SparseGrad(pad.cuda(), rows.cuda(), n.cuda()), _C.OPT_KINDS[opt], lr, l2)
    assert_close(t, exp, rtol=1e-5, atol_scale=1e-6)
    p = torch.randn(1001, generator=g)
    gr = torch.randn(1001, generator=g)
    gr[5] = 0.0
    expd = oracle.fresh_optimizer_step(p.double(), gr.double(), opt, lr)
    pc = p.clone().cuda()
    ops.dense_opt_step(pc, gr.cuda(), _C.OPT_KINDS[opt], lr, 0.0)
    assert_close(pc, expd, rtol=1e-5, atol_scale=1e-6)


@pytest.mark.parametrize("opt", ["adam", "adagrad", "gd"])
def test_dense_opt_step_multi_matches_single(opt):
    """One multi-tensor launch == one rm_dense_opt_step call per tensor, bit for bit (more tensors than one batch)."""
    ops = _ops()
    from recman_b200 import _C

    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 400 * 429, 33, 1025] + [5 + i for i in range(100)]
    ps = [torch.randn(n, generator=g).cuda() for n in sizes]
    gs = [torch.randn(n, generator=g).cuda() for n in sizes]
    single = [p.clone() for p in ps]
    for p, gr in zip(single, gs):
        ops.dense_opt_step(p, gr, _C.OPT_KINDS[opt], 0.01, 0.0)
    multi = [p.clone() for p in ps]
    ops.dense_opt_step_multi(list(zip(multi, gs)), _C.OPT_KINDS[opt], 0.01, 0.0)
    for a, b in zip(single, multi):
        assert torch.equal(a, b)
    exp = oracle.fresh_optimizer_step(ps[2].double().cpu(), gs[2].double().cpu(), opt, 0.01)
    assert_close(multi[2], exp, rtol=1e-5, atol_scale=1e-6)


@pytest.mark.parametrize("B,d,N", [(1, 5, 4), (300, 429, 32), (1000, 1677, 32), (257, 130, 64), (64, 32, 32), (513, 100, 12)])
def test_narrow_linear_backward_kernels(B, d, N):
    """dx = g W^T (padded row buffer, zero tail) and dW = x^T g (slabbed, deterministic) vs fp64; exact fp32 FMA."""
    ops = _ops()
    g = torch.Generator().manual_seed(B + d + N)
    ld = (d + 3) // 4 * 4
    x = torch.zeros(B, ld)
    x[:, :d] = torch.randn(B, d, generator=g)
    W = torch.randn(d, N, generator=g) * 0.1
    gy = torch.randn(B, N, generator=g)
    dx = ops.linear_bwd_input(gy.cuda(), W.cuda(), d_ld=ld)
    assert dx.shape == (B, ld)
    assert_close(dx[:, :d], gy.double() @ W.double().t(), rtol=1e-5, atol_scale=2e-6)
    assert torch.all(dx[:, d:] == 0)
    dW = ops.linear_bwd_weight(x.cuda(), ld, d, gy.cuda())
    assert_close(dW, x[:, :d].double().t() @ gy.double(), rtol=1e-5, atol_scale=2e-6)
    assert torch.equal(dW, ops.linear_bwd_weight(x.cuda(), ld, d, gy.cuda()))  # run-to-run bit-identical


@pytest.mark.parametrize("k", [16, 64])
def test_hot_rows_take_the_chunked_long_segment_path(k):
    """Skewed ids: a few rows collect hundreds to thousands of positions.  Long segments are summed chunk-wise (64
    positions per row group, partials in chunk order) - same association as the oracle, so still bit-exact - and the
    fused-update variant equals reduce + optimizer step."""
    ops = _ops()
    from recman_b200 import _C

    rng = np.random.RandomState(5)
    sizes = [1000, 50, 7]
    m, B = len(sizes), 3000
    offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int64)
    ids = np.stack([(rng.zipf(1.3, size=B) - 1) % v for v in sizes], 1).astype(np.int64)
    ids[:, 2] = 3  # one segment of 3000 positions = 94 chunks
    g = torch.Generator().manual_seed(6)
    grad = torch.randn(B, m, k, generator=g)
    keys = oracle.global_rows(ids, offs.numpy()).reshape(-1)
    uniq, sums, order, seg = oracle.segment_sum_sorted(keys, grad.reshape(-1, k).numpy())
    assert int(np.diff(seg).max()) == B and (np.diff(seg) > oracle.segment.SEG_LONG).sum() >= 3
    plan = ops.segment_plan(torch.from_numpy(ids).cuda(), offs.cuda(), int(offs[-1]))
    n = plan.num_unique()
    rows = ops.segment_reduce(grad.cuda(), plan, k)
    assert np.array_equal(rows[:n].cpu().numpy(), sums)
    assert torch.equal(rows[:n], ops.segment_reduce(grad.cuda(), plan, k)[:n])  # run-to-run identical
    dense = oracle.dense_table_grad(keys, grad.reshape(-1, k).numpy(), int(offs[-1]))
    assert_close(rows[:n], torch.from_numpy(dense[uniq]), rtol=1e-5, atol_scale=2e-6)
    # fused FM backward + update on the same plan == unfused reduce followed by the optimizer kernels
    ld = m * k
    x = torch.randn(B, ld, generator=g).cuda()
    dx = torch.randn(B, ld, generator=g).cuda()
    S = x.view(B, m, k).sum(1).contiguous()
    g_fm, g_lin = torch.randn(B, generator=g).cuda(), torch.randn(B, generator=g).cuda()
    total = int(offs[-1])
    tab = torch.randn(total, k, generator=g).cuda()
    bt, lt = torch.randn(total, generator=g).cuda(), torch.randn(total, generator=g).cuda()
    tab2, bt2, lt2 = tab.clone(), bt.clone(), lt.clone()
    r, ob, ol = ops.emb_fm_bwd(dx, x, ld, S, g_fm, g_lin, plan, k, True, True, True)
    exp_rows = (dx.view(B, m, k) + g_fm[:, None, None] * (S[:, None, :] - x.view(B, m, k))).reshape(-1, k)
    dense2 = oracle.dense_table_grad(keys, exp_rows.cpu().numpy(), total)
    assert_close(r[:n], torch.from_numpy(dense2[uniq]), rtol=1e-5, atol_scale=5e-6)
    ops.sparse_opt_step(tab, ops.SparseGrad(plan.uniq_rows, r, plan.n_unique), _C.OPT_KINDS["adam"], 0.01)
    ops.sparse_opt_step(bt, ops.SparseGrad(plan.uniq_rows, ob, plan.n_unique), _C.OPT_KINDS["adam"], 0.01)
    ops.sparse_opt_step(lt, ops.SparseGrad(plan.uniq_rows, ol, plan.n_unique), _C.OPT_KINDS["adam"], 0.01)
    ops.emb_fm_bwd_update(dx, x, ld, S, g_fm, g_lin, plan, k, tab2, bt2, lt2, _C.OPT_KINDS["adam"], 0.01)
    assert torch.equal(tab, tab2) and torch.equal(bt, bt2) and torch.equal(lt, lt2)
