"""Parity of the fused DeepFM tower kernels (rm_tower_*, through the C ABI) against the CPU oracle.

Forward: gathered rows bit-exact, FM / first-order / first-layer outputs within 1e-5 of the fp64 oracle.
Backward: per-row summed gradients, k=1 gradients, dW1 and the updated tables within 1e-5; run-to-run bit-identical.
"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _ops():
    from recman_b200 import ops

    return ops


def assert_close(got, exp, rtol=RTOL, atol_scale=1e-6, msg=""):
    got = got.detach().cpu().double()
    exp = exp.detach().cpu().double()
    atol = atol_scale * max(1e-30, float(exp.abs().max())) if exp.numel() else 0.0
    torch.testing.assert_close(got, exp, rtol=rtol, atol=atol, msg=lambda m: f"{msg}: {m}")


def _setup(sizes, k, B, nd, N1, seed=0, dup=False):
    g = torch.Generator().manual_seed(seed)
    m = len(sizes)
    tabs = [torch.randn(v, k, generator=g) * 0.1 for v in sizes]
    offs = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int64)
    total = int(offs[-1])
    scal = torch.randn(total, 2, generator=g) * 0.1
    cols = []
    for v in sizes:
        hi = min(v, 3) if dup else v
        cols.append(torch.randint(0, hi, (B,), generator=g))
    ids = torch.stack(cols, 1).contiguous()
    if B > 1:
        ids[0] = 0
        ids[-1] = torch.tensor([v - 1 for v in sizes])
    dense = torch.randn(B, nd, generator=g) if nd else None
    d = m * k + nd
    W1 = torch.randn(d, N1, generator=g) * 0.05
    b1 = torch.randn(N1, generator=g) * 0.05
    lin_dense = torch.randn(nd, generator=g) * 0.1 if nd else None
    return dict(tabs=tabs, table=torch.cat(tabs, 0), offs=offs, scal=scal, ids=ids, dense=dense, W1=W1, b1=b1,
                lin_dense=lin_dense, m=m, k=k, B=B, nd=nd, N1=N1, total=total)


def _forward_oracle(s, dtype=torch.float64):
    """DeepFM front end + first DNN matmul through the oracle's layers (layers.py:238-261, 457-478, 589-592)."""
    m = s["m"]
    ids = s["ids"]
    tabs = [t.to(dtype) for t in s["tabs"]]
    offs = s["offs"]
    biases = [s["scal"][int(offs[f]):int(offs[f + 1]), 0:1].to(dtype) for f in range(m)]
    embeds, bias = oracle.feat_embedding_layer(tabs, [ids[:, f] for f in range(m)], biases)  # [B,m,k], [B,m,1]
    fm = oracle.fm_layer(embeds, bias).reshape(-1)
    rows = oracle.global_rows(ids.numpy(), offs.numpy())
    lin = s["scal"][:, 1].to(dtype)[torch.from_numpy(rows)].sum(1)
    x = embeds.reshape(embeds.shape[0], -1)
    if s["nd"]:
        dn = s["dense"].to(dtype)
        lin = lin + dn @ s["lin_dense"].to(dtype)
        x = torch.cat([x, dn], 1)
    y1 = x @ s["W1"].to(dtype) + s["b1"].to(dtype)
    return embeds, fm, lin, y1, embeds.sum(1)


def test_umma_mn_major_layout():
    """The MN-major SWIZZLE_128B operand layout the weight-gradient GEMM relies on: exact small-integer product."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    K = 64
    At = torch.randint(-4, 5, (K, 128), generator=g).float()
    Bt = torch.randint(-4, 5, (K, 32), generator=g).float()
    D, st = ops.umma_probe(At.cuda(), Bt.cuda(), 0)
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    assert torch.equal(D.cpu(), At.t() @ Bt)


@pytest.mark.parametrize("shape", [
    # sizes, k, B, nd, N1
    ([50, 7, 1000, 3, 200], 64, 300, 3, 32),
    ([944, 1683, 3, 22, 796, 11, 64], 64, 401, 13, 32),
    ([50, 7, 1000], 32, 200, 0, 16),
    ([5, 9], 32, 128, 13, 64),
    ([30] * 26, 64, 515, 13, 32),
    ([17, 5, 300, 41], 32, 77, 5, 24),
])
def test_tower_forward(shape):
    ops = _ops()
    sizes, k, B, nd, N1 = shape
    s = _setup(sizes, k, B, nd, N1)
    assert ops.tower_supported(len(sizes), k, nd, N1)
    st = ops.new_status("cuda")
    dev = lambda t: None if t is None else t.cuda()
    y1, fm, lin, S, x = ops.tower_fwd(dev(s["table"]), dev(s["scal"]), dev(s["offs"]), dev(s["ids"]), dev(s["dense"]),
                                      dev(s["lin_dense"]), dev(s["W1"]), dev(s["b1"]), want_x=True, status=st)
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    embeds, o_fm, o_lin, o_y1, o_S = _forward_oracle(s)
    m = len(sizes)
    assert torch.equal(x[:, : m * k].cpu(), embeds.float().reshape(B, -1))  # gathered rows: bit-exact
    if nd:
        assert torch.equal(x[:, m * k : m * k + nd].cpu(), s["dense"])
    assert_close(S, o_S, msg="S")
    assert_close(fm, o_fm, msg="fm")
    assert_close(lin, o_lin, msg="lin")
    assert_close(y1, o_y1, msg="y1")
    # without the row buffer: same results
    y1b, fmb, linb, Sb, xb = ops.tower_fwd(dev(s["table"]), dev(s["scal"]), dev(s["offs"]), dev(s["ids"]),
                                           dev(s["dense"]), dev(s["lin_dense"]), dev(s["W1"]), dev(s["b1"]))
    assert xb is None and torch.equal(y1b, y1) and torch.equal(fmb, fm) and torch.equal(linb, lin) and torch.equal(Sb, S)


def test_tower_forward_bad_id_zero_fills_and_flags():
    ops = _ops()
    sizes, k, B, nd, N1 = [50, 7, 1000], 64, 130, 2, 32
    s = _setup(sizes, k, B, nd, N1)
    s["ids"][5, 1] = 7  # == feat_size: outside
    s["ids"][77, 2] = -3
    st = ops.new_status("cuda")
    dev = lambda t: None if t is None else t.cuda()
    y1, fm, lin, S, x = ops.tower_fwd(dev(s["table"]), dev(s["scal"]), dev(s["offs"]), dev(s["ids"]), dev(s["dense"]),
                                      dev(s["lin_dense"]), dev(s["W1"]), dev(s["b1"]), want_x=True, status=st)
    torch.cuda.synchronize()
    assert int(st.item()) & 1
    assert torch.count_nonzero(x[5, k : 2 * k]) == 0 and torch.count_nonzero(x[77, 2 * k : 3 * k]) == 0
    assert torch.isfinite(y1).all()


def _backward_reference(s, g1, g_fm, g_lin):
    """fp64: per unique row the summed gradient of (MLP layer-1 input gradient + FM backward), k=1 gradients, dW1."""
    m, k, B = s["m"], s["k"], s["B"]
    rows = oracle.global_rows(s["ids"].numpy(), s["offs"].numpy())  # [B, m]
    T = s["table"].double()
    x = T[torch.from_numpy(rows)]  # [B, m, k]
    S = x.sum(1)
    W1 = s["W1"].double()
    dx = (g1.double() @ W1[: m * k].t()).reshape(B, m, k)
    G = dx + g_fm.double()[:, None, None] * (S[:, None, :] - x)  # dL/d e[b,f,:]  (layers.py:457-478 backward)
    dense_rows = torch.zeros(s["total"], k, dtype=torch.float64)
    dense_rows.index_add_(0, torch.from_numpy(rows.reshape(-1)), G.reshape(-1, k))
    dsc = torch.zeros(s["total"], 2, dtype=torch.float64)
    gsc = torch.stack([g_fm.double(), g_lin.double()], 1)[:, None, :].expand(B, m, 2).reshape(-1, 2)
    dsc.index_add_(0, torch.from_numpy(rows.reshape(-1)), gsc)
    dW1 = x.reshape(B, m * k).t() @ g1.double()
    return rows, dense_rows, dsc, dW1, S.float()


@pytest.mark.parametrize("case", [
    # sizes, B, dup, unit
    ([50, 7, 1000, 3, 200], 300, False, 2048),
    ([944, 1683, 3, 22, 796, 11, 64], 1111, False, 256),
    ([40, 40, 40], 700, True, 128),       # 3 distinct ids per field: segments of ~230 positions across tiles and units
    ([100000] * 4, 4096, False, 2048),    # almost all singletons
    ([30] * 26, 515, False, 512),
])
@pytest.mark.parametrize("opt", ["gd", "adam"])
def test_tower_backward(case, opt):
    from recman_b200 import _C

    ops = _ops()
    sizes, B, dup, unit = case
    k, nd, N1 = 64, 3, 32
    s = _setup(sizes, k, B, nd, N1, seed=5, dup=dup)
    m = s["m"]
    g = torch.Generator().manual_seed(11)
    g1 = torch.randn(B, N1, generator=g) * 1e-2
    g_fm = torch.randn(B, generator=g) * 1e-2
    g_lin = torch.randn(B, generator=g) * 1e-2
    rows, o_rows, o_sc, o_dW1, S = _backward_reference(s, g1, g_fm, g_lin)
    table, scal = s["table"].cuda(), s["scal"].cuda()
    st = ops.new_status("cuda")
    plan = ops.tower_plan(s["ids"].cuda(), s["offs"].cuda(), s["total"], unit=unit, status=st)
    torch.cuda.synchronize()
    keys = plan.sorted_keys.cpu().numpy().view(np.uint32).astype(np.int64)
    pos = plan.sorted_pos.cpu().numpy()
    # sorted by (row, position): bit-exact against a stable argsort
    order = np.argsort(rows.reshape(-1), kind="stable")
    assert np.array_equal(pos, order.astype(np.int32)) and np.array_equal(keys, rows.reshape(-1)[order])
    fb = plan.field_bounds.cpu().numpy()
    assert np.array_equal(fb, np.arange(m + 1) * B)
    ub_all = plan.unit_bounds.cpu().numpy()
    ub, hot = ub_all[:-1], int(ub_all[-1])
    for p in ub:  # every cut sits on a segment head
        assert p in fb or keys[p] != keys[p - 1]
    # the hot-row flag behind the cuts: some row holds more than 32 positions (sorted keys compared 32 apart)
    probe = np.arange(0, len(keys) - 32, 32)
    assert hot == int(np.any(keys[probe] == keys[probe + 32]))

    lr = 0.5 if opt == "gd" else 1e-3
    kind = _C.OPT_KINDS[opt]
    args = (plan, g1.cuda(), S.cuda(), g_fm.cuda(), g_lin.cuda(), s["W1"].cuda(), kind, lr)
    # 1) gradients only
    t0, s0 = table.clone(), scal.clone()
    dW1, out_rows, out_scal = ops.tower_bwd_update(t0, s0, *args, update=False, debug=True, status=st)
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    assert torch.equal(t0, table) and torch.equal(s0, scal)
    closing = np.flatnonzero(np.append(keys[1:] != keys[:-1], True))  # sorted position closing each segment
    uniq = keys[closing]
    assert_close(out_rows.cpu()[closing], o_rows[uniq], atol_scale=1e-5, msg="summed gradient rows")
    assert_close(out_scal.cpu()[closing], o_sc[uniq], atol_scale=1e-5, msg="k=1 gradients")
    mask = np.ones(B * m, dtype=bool)
    mask[closing] = False
    assert torch.count_nonzero(out_rows.cpu()[mask]) == 0
    assert_close(dW1, o_dW1, atol_scale=1e-5, msg="dW1")
    # 2) fused update, twice: deterministic and equal to the oracle's fresh-optimizer step on the touched rows
    outs = []
    for _ in range(2):
        t1, s1 = table.clone(), scal.clone()
        dW1b = ops.tower_bwd_update(t1, s1, *args, status=st)
        torch.cuda.synchronize()
        outs.append((t1, s1, dW1b))
    assert int(st.item()) == 0
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
    assert torch.equal(outs[0][2], dW1)
    t1, s1, _ = outs[0]
    exp_t = s["table"].double().clone()
    exp_s = s["scal"].double().clone()
    ut = torch.from_numpy(uniq)
    exp_t[ut] = oracle.fresh_optimizer_step(exp_t[ut], o_rows[ut], opt, lr)
    exp_s[ut] = oracle.fresh_optimizer_step(exp_s[ut], o_sc[ut], opt, lr)
    assert_close(t1, exp_t, atol_scale=2e-6, msg="updated table")
    assert_close(s1, exp_s, atol_scale=2e-6, msg="updated k=1 tables")
    untouched = np.ones(s["total"], dtype=bool)
    untouched[uniq] = False
    assert torch.equal(t1.cpu()[untouched], s["table"][untouched])


def test_tower_backward_drops_out_of_range_ids():
    """ids outside their table sort behind every row and never reach the update (no foreign-row writes)."""
    ops = _ops()
    sizes, B = [50, 7, 1000], 260
    k, nd, N1 = 64, 0, 32
    s = _setup(sizes, k, B, nd, N1, seed=2)
    s["ids"][3, 1] = 9
    s["ids"][100, 0] = -1
    st = ops.new_status("cuda")
    plan = ops.tower_plan(s["ids"].cuda(), s["offs"].cuda(), s["total"], unit=256, status=st)
    torch.cuda.synchronize()
    assert int(st.item()) & 1
    fb = plan.field_bounds.cpu().numpy()
    assert fb[-1] == B * 3 - 2 and fb[1] == B - 1 and fb[2] == 2 * B - 2
    keys = plan.sorted_keys.cpu().numpy().view(np.uint32)
    assert (keys[-2:] == s["total"]).all()
    g = torch.Generator().manual_seed(1)
    g1 = torch.randn(B, N1, generator=g) * 1e-2
    g_fm = torch.randn(B, generator=g) * 1e-2
    S = torch.randn(B, k, generator=g)
    t1, s1 = s["table"].cuda(), s["scal"].cuda()
    st2 = ops.new_status("cuda")
    ops.tower_bwd_update(t1, s1, plan, g1.cuda(), S.cuda(), g_fm.cuda(), None, s["W1"].cuda(), 2, 0.1, status=st2)
    torch.cuda.synchronize()
    assert int(st2.item()) == 0
    valid = s["ids"].clone()
    valid[3, 1] = 0
    valid[100, 0] = 0
    rows = oracle.global_rows(valid.numpy(), s["offs"].numpy())
    touched = np.zeros(s["total"], dtype=bool)
    touched[rows.reshape(-1)] = True
    # rows 0 of fields 0 / 1 may or may not be touched by other samples; every untouched row must be unchanged
    rows_real = rows.copy().reshape(-1)
    keep = np.ones(B * 3, dtype=bool)
    keep[3 * 3 + 1] = False
    keep[100 * 3 + 0] = False
    touched2 = np.zeros(s["total"], dtype=bool)
    touched2[rows_real[keep]] = True
    assert torch.equal(t1.cpu()[~touched2], s["table"][~touched2])
    assert torch.isfinite(t1).all()


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "identity"])
@pytest.mark.parametrize("task", ["classification", "regression"])
@pytest.mark.parametrize("B", [1, 300, 4097])
def test_deepfm_head_forward_backward(act, task, B):
    """rm_deepfm_head: second DNN layer .. loss and every gradient against fp64 autograd over the oracle's layers."""
    from recman_b200 import _C

    ops = _ops()
    g = torch.Generator().manual_seed(B)
    N = 32
    y1 = torch.randn(B, N, generator=g)
    fm = torch.randn(B, generator=g) * 0.3
    lin = torch.randn(B, generator=g) * 0.3
    w0 = torch.randn(1, generator=g) * 0.1
    W2 = torch.randn(N, N, generator=g) * 0.2
    b2 = torch.randn(N, generator=g) * 0.1
    w3 = torch.randn(N, 1, generator=g) * 0.2
    b3 = torch.randn(1, generator=g) * 0.1
    y = (torch.rand(B, generator=g) < 0.3).float() if task == "classification" else torch.randn(B, generator=g)
    leaf = lambda t: t.double().clone().requires_grad_()
    Y1, FM, LIN, W0, W2d, B2, W3, B3 = map(leaf, (y1, fm, lin, w0, W2, b2, w3, b3))
    fn = {"relu": oracle.relu, "leaky_relu": oracle.leaky_relu_tf, "identity": (lambda t: t)}[act]
    h2 = fn(fn(Y1) @ W2d + B2)
    logit = (LIN + W0 + FM).reshape(-1, 1) + (h2 @ W3 + B3)
    pred = oracle.prediction(logit, task)
    loss = oracle.create_loss(y.double(), pred, task)
    loss.backward()
    out = ops.deepfm_head(y1.cuda(), fm.cuda(), lin.cuda(), w0.cuda(), W2.cuda(), b2.cuda(), w3.reshape(-1).cuda(),
                          b3.cuda(), y.cuda(), _C.ACT_KINDS[act], 0 if task == "classification" else 1)
    torch.cuda.synchronize()
    assert_close(out["logit"], logit.detach().reshape(-1), atol_scale=2e-6, msg="logit")
    assert_close(out["pred"], pred.detach().reshape(-1), atol_scale=2e-6, msg="pred")
    assert_close(out["loss"], loss.detach().reshape(1), atol_scale=2e-6, msg="loss")
    for name, got, exp in [("g1", out["g1"], Y1.grad), ("g", out["g"], FM.grad), ("g(lin)", out["g"], LIN.grad),
                           ("dW2", out["dW2"], W2d.grad), ("db2", out["db2"], B2.grad),
                           ("dw3", out["dw3"], W3.grad.reshape(-1)), ("db3", out["dscal"], B3.grad),
                           ("dw0", out["dscal"], W0.grad), ("db1", out["db1"], Y1.grad.sum(0))]:
        assert_close(got, exp, atol_scale=1e-5, msg=name)
    # forward only, and run-to-run identical reductions
    lg, pr = ops.deepfm_head(y1.cuda(), fm.cuda(), lin.cuda(), w0.cuda(), W2.cuda(), b2.cuda(), w3.reshape(-1).cuda(),
                             b3.cuda(), None, _C.ACT_KINDS[act], 0 if task == "classification" else 1)
    assert torch.equal(lg, out["logit"]) and torch.equal(pr, out["pred"])
    out2 = ops.deepfm_head(y1.cuda(), fm.cuda(), lin.cuda(), w0.cuda(), W2.cuda(), b2.cuda(), w3.reshape(-1).cuda(),
                           b3.cuda(), y.cuda(), _C.ACT_KINDS[act], 0 if task == "classification" else 1)
    for k in out:
        assert torch.equal(out[k], out2[k]), k
    # with the samples' dense features: the first layer's / first-order term's gradients that involve them
    nd = 13
    dense = torch.randn(B, nd, generator=g)
    out3 = ops.deepfm_head(y1.cuda(), fm.cuda(), lin.cuda(), w0.cuda(), W2.cuda(), b2.cuda(), w3.reshape(-1).cuda(),
                           b3.cuda(), y.cuda(), _C.ACT_KINDS[act], 0 if task == "classification" else 1, dense=dense.cuda())
    for k in out:
        assert torch.equal(out[k], out3[k]), k
    assert_close(out3["dW1_dense"], dense.double().t() @ Y1.grad, atol_scale=1e-5, msg="dW1_dense")
    assert_close(out3["dlin_dense"], dense.double().t() @ FM.grad, atol_scale=1e-5, msg="dlin_dense")


def test_c5_shape_rows_indices_and_gradients():
    """BASELINE config 5 shape: B = 65 536, 26 fields x 1 M rows, k = 64 (6.7 GB of tables; the kernels see the same
    tile counts, unit cuts and id ranges as the timed workload).  Gathered rows and the sorted (row, position) indices
    bit-exact; summed gradient rows, k=1 gradients and dW1 within 1e-5 of the fp64 reference on the touched rows."""
    ops = _ops()
    B, m, k, nd, N1, rows_per = 65536, 26, 64, 13, 32, 1_000_000
    g = torch.Generator().manual_seed(2019)
    dev = "cuda"
    total = m * rows_per
    table = torch.empty(total, k, device=dev).normal_(0.0, 0.05, generator=torch.Generator(device=dev).manual_seed(1))
    scal = torch.empty(total, 2, device=dev).normal_(0.0, 0.05, generator=torch.Generator(device=dev).manual_seed(2))
    offs = (torch.arange(m + 1) * rows_per).long()
    ids = torch.randint(0, rows_per, (B, m), generator=g)
    ids[:64, :] = 7          # a hot row per field: 64-position segments
    ids[0] = 0
    ids[-1] = rows_per - 1
    dense = torch.randn(B, nd, generator=g)
    W1 = torch.randn(m * k + nd, N1, generator=g) * 0.05
    b1 = torch.randn(N1, generator=g) * 0.05
    lin_dense = torch.randn(nd, generator=g) * 0.1
    st = ops.new_status(dev)
    y1, fm, lin, S, x = ops.tower_fwd(table, scal, offs.cuda(), ids.cuda(), dense.cuda(), lin_dense.cuda(), W1.cuda(),
                                      b1.cuda(), want_x=True, status=st)
    rows = oracle.global_rows(ids.numpy(), offs.numpy())  # [B, m]
    rows_t = torch.from_numpy(rows).cuda()
    gathered = table[rows_t.reshape(-1)].reshape(B, m * k)
    assert torch.equal(x[:, : m * k], gathered)  # gathered rows: bit-exact
    # forward values on a sample of the batch against fp64
    sel = torch.arange(0, B, 257)
    xs = gathered[sel.cuda()].double().cpu()
    xin = torch.cat([xs, dense[sel].double()], 1)
    assert_close(y1[sel.cuda()], xin @ W1.double() + b1.double(), msg="y1")
    e = xs.reshape(len(sel), m, k)
    sc = scal[rows_t[sel.cuda()].reshape(-1)].reshape(len(sel), m, 2).double().cpu()
    assert_close(fm[sel.cuda()], sc[:, :, 0].sum(1) + 0.5 * ((e.sum(1) ** 2).sum(1) - (e ** 2).sum((1, 2))), msg="fm")
    assert_close(lin[sel.cuda()], sc[:, :, 1].sum(1) + dense[sel].double() @ lin_dense.double(), msg="lin")
    # sorted (row, position) pairs: bit-exact against a stable argsort
    plan = ops.tower_plan(ids.cuda(), offs.cuda(), total, status=st)
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    keys = plan.sorted_keys.cpu().numpy().view(np.uint32).astype(np.int64)
    order = np.argsort(rows.reshape(-1), kind="stable")
    assert np.array_equal(plan.sorted_pos.cpu().numpy(), order.astype(np.int32))
    assert np.array_equal(keys, rows.reshape(-1)[order])
    # backward: gradients on the touched rows
    g1 = torch.randn(B, N1, generator=g) * 1e-2
    g_fm = torch.randn(B, generator=g) * 1e-2
    g_lin = torch.randn(B, generator=g) * 1e-2
    dW1, out_rows, out_scal = ops.tower_bwd_update(table, scal, plan, g1.cuda(), S, g_fm.cuda(), g_lin.cuda(), W1.cuda(),
                                                   0, 1e-3, update=False, debug=True, status=st)
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    # fp64 reference of the per-position gradient rows, on the GPU in chunks (109 M values)
    Sd = gathered.reshape(B, m, k).double().sum(1)
    dx = (g1.cuda().double() @ W1.cuda().double()[: m * k].t()).reshape(B, m, k)
    G = dx + g_fm.cuda().double()[:, None, None] * (Sd[:, None, :] - gathered.reshape(B, m, k).double())
    closing = np.flatnonzero(np.append(keys[1:] != keys[:-1], True))
    starts = np.concatenate([[0], closing[:-1] + 1])
    Gs = G.reshape(-1, k)[torch.from_numpy(order).cuda()]  # per-position rows in sorted order
    csum = torch.cat([torch.zeros(1, k, dtype=torch.float64, device=dev), Gs.cumsum(0)])
    exp_rows = csum[torch.from_numpy(closing + 1).cuda()] - csum[torch.from_numpy(starts).cuda()]
    got_rows = out_rows[torch.from_numpy(closing).cuda()]
    assert_close(got_rows, exp_rows, atol_scale=1e-5, msg="summed gradient rows")
    gsc = torch.stack([g_fm, g_lin], 1).double().cuda()[torch.from_numpy(order // m).cuda()]
    cs2 = torch.cat([torch.zeros(1, 2, dtype=torch.float64, device=dev), gsc.cumsum(0)])
    exp_sc = cs2[torch.from_numpy(closing + 1).cuda()] - cs2[torch.from_numpy(starts).cuda()]
    assert_close(out_scal[torch.from_numpy(closing).cuda()], exp_sc, atol_scale=1e-5, msg="k=1 gradients")
    assert_close(dW1, gathered.double().t() @ g1.cuda().double(), atol_scale=1e-5, msg="dW1")


def _shard_tables(s, W):
    """Cyclic row sharding of the test tables: rank r keeps rows r, r+W, ... of every table (local row = row // W)."""
    m = s["m"]
    sizes = [int(s["offs"][f + 1] - s["offs"][f]) for f in range(m)]
    local_sizes = [(v + W - 1) // W for v in sizes]
    loffs = np.concatenate([[0], np.cumsum(local_sizes)])
    total_local = int(loffs[-1])
    tabs, scals = [], []
    for r in range(W):
        T = torch.zeros(total_local, s["k"])
        Sc = torch.zeros(total_local, 2)
        for f in range(m):
            rows = torch.arange(r, max(sizes[f], r), W)
            T[loffs[f] : loffs[f] + rows.numel()] = s["table"][int(s["offs"][f]) + rows]
            Sc[loffs[f] : loffs[f] + rows.numel()] = s["scal"][int(s["offs"][f]) + rows]
        tabs.append(T.cuda())
        scals.append(Sc.cuda())
    return sizes, loffs, total_local, tabs, scals


@pytest.mark.parametrize("W", [2, 3, 4, 5, 8])
def test_tower_forward_sharded_rows_from_their_owners(W):
    """rm_tower_fwd_p2p with the W shards living on one device: same outputs as the unsharded kernel, bit for bit
    (the rows are the same rows, wherever they live), including tables with fewer rows than ranks."""
    ops = _ops()
    sizes, k, B, nd, N1 = [50, 7, 1000, 3, 200, 31, 2, 90, 1], 64, 300, 5, 32
    s = _setup(sizes, k, B, nd, N1, seed=4)
    szs, loffs, total_local, tabs, scals = _shard_tables(s, W)
    dev = lambda t: None if t is None else t.cuda()
    st = ops.new_status("cuda")
    y1, fm, lin, S, _ = ops.tower_fwd(dev(s["table"]), dev(s["scal"]), dev(s["offs"]), dev(s["ids"]), dev(s["dense"]),
                                      dev(s["lin_dense"]), dev(s["W1"]), dev(s["b1"]), status=st)
    y1p, fmp, linp, Sp = ops.tower_fwd_p2p([t.data_ptr() for t in tabs], [t.data_ptr() for t in scals], k,
                                           torch.tensor(szs, dtype=torch.int64).cuda(),
                                           torch.tensor(loffs[:-1], dtype=torch.int64).cuda(), dev(s["ids"]),
                                           dev(s["dense"]), dev(s["lin_dense"]), dev(s["W1"]), dev(s["b1"]), status=st)
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    assert torch.equal(y1p, y1) and torch.equal(fmp, fm) and torch.equal(linp, lin) and torch.equal(Sp, S)


@pytest.mark.parametrize("W", [2, 3, 4])
def test_tower_backward_sharded_owner_view(W):
    """Every owner runs rm_tower_shard_plan + rm_tower_bwd_update on its shard over the ids / per-sample operands of all
    W ranks (here: one device); together the shards receive the update the unsharded kernel applies to the full table,
    and the owners' dW1 partials add up to the unsharded dW1."""
    from recman_b200 import _C

    ops = _ops()
    sizes, k, b, nd, N1 = [50, 7, 1000, 3, 200, 31, 2, 90, 1], 64, 200, 0, 32
    B = W * b  # global batch: rank r owns samples [r*b, (r+1)*b)
    s = _setup(sizes, k, B, nd, N1, seed=8)
    m = s["m"]
    szs, loffs, total_local, tabs, scals = _shard_tables(s, W)
    g = torch.Generator().manual_seed(5)
    g1 = (torch.randn(B, N1, generator=g) * 1e-2).cuda()
    g_fm = (torch.randn(B, generator=g) * 1e-2).cuda()
    g_lin = (torch.randn(B, generator=g) * 1e-2).cuda()
    st = ops.new_status("cuda")
    # unsharded reference run of the same kernels
    table, scal = s["table"].cuda(), s["scal"].cuda()
    _, _, _, S, _ = ops.tower_fwd(table, scal, s["offs"].cuda(), s["ids"].cuda(), None, None, s["W1"].cuda(), s["b1"].cuda(),
                                  status=st)
    plan = ops.tower_plan(s["ids"].cuda(), s["offs"].cuda(), s["total"], unit=256, status=st)
    dW1 = ops.tower_bwd_update(table, scal, plan, g1, S, g_fm, g_lin, s["W1"].cuda(), _C.OPT_KINDS["gd"], 0.5, status=st)
    # owners
    gids = s["ids"].to(torch.int32).cuda()
    dW1_sum = torch.zeros_like(dW1, dtype=torch.float64)
    fs = torch.tensor(szs, dtype=torch.int64).cuda()
    lo = torch.tensor(loffs[:-1], dtype=torch.int64).cuda()
    for r in range(W):
        n_cap = B * m
        tp = ops.tower_shard_plan(gids, W, r, fs, lo, total_local, n_cap, B, unit=256, status=st)
        dW1_r = ops.tower_bwd_update(tabs[r], scals[r], tp, g1, S, g_fm, g_lin, s["W1"].cuda(), _C.OPT_KINDS["gd"], 0.5,
                                     status=st)
        dW1_sum += dW1_r.double()
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    assert_close(dW1_sum, dW1.double(), atol_scale=1e-5, msg="dW1 partials")
    for r in range(W):
        for f in range(m):
            rows = torch.arange(r, max(szs[f], r), W)
            got = tabs[r][loffs[f] : loffs[f] + rows.numel()].cpu()
            exp = table[int(s["offs"][f]) + rows.cuda()].cpu()
            assert torch.equal(got, exp), (r, f)  # same positions, same order, same arithmetic: bit-identical rows
            assert torch.equal(scals[r][loffs[f] : loffs[f] + rows.numel()].cpu(), scal[int(s["offs"][f]) + rows.cuda()].cpu())
