"""The oracle against the committed golden fixtures (CPU) and the CUDA kernels against the same fixtures (GPU)."""
import os

import numpy as np
import pytest
import torch

import oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_v1.npz"))
T = lambda name, dt=torch.float64: torch.from_numpy(G[name]).to(dt)


def _tabs():
    offs = G["gather_offsets"]
    table = torch.from_numpy(G["gather_table"])
    return [table[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)], table, offs


def test_oracle_reproduces_golden_gather_pool_segment():
    tabs, table, offs = _tabs()
    ids = torch.from_numpy(G["gather_ids"])
    e, _ = oracle.feat_embedding_layer(tabs, [ids[:, f] for f in range(ids.shape[1])])
    assert np.array_equal(e.numpy(), G["gather_out"])
    p = oracle.embedding_lookup_sqrtn(tabs[3], torch.from_numpy(G["pooled_values"]), torch.from_numpy(G["pooled_offsets"]))
    assert np.array_equal(p[:, 0].numpy(), G["pooled_out"])
    keys = oracle.global_rows(G["gather_ids"], offs).reshape(-1)
    uniq, sums, order, seg = oracle.segment_sum_sorted(keys, G["seg_grad"])
    assert np.array_equal(uniq, G["seg_uniq"]) and np.array_equal(sums, G["seg_sums"])
    assert np.array_equal(order, G["seg_order"]) and np.array_equal(seg, G["seg_start"])


def test_oracle_reproduces_golden_fm_cross_cin_loss_opt():
    y = oracle.fm_layer(T("fm_e"), T("fm_bias"))
    np.testing.assert_allclose(y.numpy(), G["fm_out"], rtol=1e-13, atol=1e-13)
    yc = oracle.cross_net(T("cross_x"), T("cross_w"), T("cross_b"), T("cross_wo"), T("cross_w0"))
    np.testing.assert_allclose(yc.numpy(), G["cross_out"], rtol=1e-13, atol=1e-13)
    filt = [T(f"cin_filter_{i}") for i in range(3)]
    fb = [T(f"cin_bias_{i}") for i in range(3)]
    yn = oracle.cin(T("cin_x"), filt, fb, T("cin_w"), T("cin_w0"))
    np.testing.assert_allclose(yn.numpy(), G["cin_out"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(oracle.binary_crossentropy(T("bce_y"), T("bce_p")).numpy(), G["bce_out"], rtol=1e-13)
    for name in ("adam", "adagrad", "gd"):
        np.testing.assert_allclose(oracle.fresh_optimizer_step(T("opt_p"), T("opt_g"), name, 0.01).numpy(),
                                   G[f"opt_{name}"], rtol=1e-13, atol=1e-15)


def _close(got, exp, scale=1e-5):
    got = got.detach().cpu().double().numpy()
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=scale * max(np.abs(exp).max(), 1e-30))


@pytest.mark.gpu
def test_kernels_against_golden():
    from recman_b200 import _C, ops

    tabs, table, offs = _tabs()
    ids = torch.from_numpy(G["gather_ids"]).cuda()
    offs_d = torch.from_numpy(offs).cuda()
    got = ops.gather(table.cuda(), offs_d, ids)
    assert np.array_equal(got.cpu().numpy(), G["gather_out"])  # gathered rows: bit exact
    pooled = ops.gather_pooled(table.cuda(), int(offs[3]), int(offs[4] - offs[3]), torch.from_numpy(G["pooled_values"]).cuda(),
                               torch.from_numpy(G["pooled_offsets"]).cuda())
    assert np.array_equal(pooled.cpu().numpy(), G["pooled_out"])
    plan = ops.segment_plan(ids, offs_d, int(offs[-1]))
    n = plan.num_unique()
    assert np.array_equal(plan.uniq_rows[:n].cpu().numpy(), G["seg_uniq"])  # gradient row indices: bit exact
    assert np.array_equal(plan.sorted_pos.cpu().numpy().astype(np.int64), G["seg_order"])
    rows = ops.segment_reduce(torch.from_numpy(G["seg_grad"]).cuda(), plan, G["seg_grad"].shape[1])
    assert np.array_equal(rows[:n].cpu().numpy(), G["seg_sums"])  # same add order: bit exact
    # FM
    e, bias, gy = T("fm_e", torch.float32).cuda(), T("fm_bias", torch.float32).cuda(), T("fm_gout", torch.float32).cuda()
    y, S = ops.fm_fwd(e, bias)
    _close(y, G["fm_out"].reshape(-1))
    de, db = ops.fm_bwd(e, S, gy)
    _close(de, G["fm_de"])
    _close(db, G["fm_dbias"].reshape(db.shape))
    # cross
    f32 = lambda n_: T(n_, torch.float32).cuda()
    logit, dots = ops.cross_fwd(f32("cross_x"), f32("cross_w"), f32("cross_b"), f32("cross_wo").reshape(-1), f32("cross_w0"))
    _close(logit, G["cross_out"].reshape(-1))
    dx, dw, dbb, dwo, dw0 = ops.cross_bwd(f32("cross_x"), f32("cross_w"), f32("cross_b"), f32("cross_wo").reshape(-1), dots,
                                          f32("cross_gout"))
    for got_, name in [(dx, "cross_dx"), (dw, "cross_dw"), (dbb, "cross_db"), (dwo.reshape(-1, 1), "cross_dwo"), (dw0, "cross_dw0")]:
        _close(got_, G[name])
    # optimizers
    for name in ("adam", "adagrad", "gd"):
        p = f32("opt_p").clone()
        ops.dense_opt_step(p, f32("opt_g"), _C.OPT_KINDS[name], 0.01, 0.0)
        _close(p, G[f"opt_{name}"], scale=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "3xtf32"])
def test_cin_stack_against_golden(precision):
    """The whole CIN layer (3 layers, split-half, pooling, head) through recman.th.layers.CIN, fwd + bwd."""
    from recman_b200.th.layers import CIN, leaky_relu

    variables = {}
    for i in range(3):
        variables[f"cin_filter_{i}"] = torch.nn.Parameter(T(f"cin_filter_{i}", torch.float32).cuda())
        variables[f"cin_bias_{i}"] = torch.nn.Parameter(T(f"cin_bias_{i}", torch.float32).cuda())
    variables["cin_w"] = torch.nn.Parameter(T("cin_w", torch.float32).cuda())
    variables["cin_w0"] = torch.nn.Parameter(T("cin_w0", torch.float32).cuda())
    x = T("cin_x", torch.float32).cuda().requires_grad_()
    layer = CIN(variables, (6, 4, 6), leaky_relu, [1, 1, 1, 1], precision=precision)
    y = layer(x)
    _close(y, G["cin_out"], scale=2e-6)
    y.backward(T("cin_gout", torch.float32).cuda())
    _close(x.grad, G["cin_dx"])
    _close(variables["cin_w"].grad, G["cin_dw"])
    for i in range(3):
        _close(variables[f"cin_filter_{i}"].grad, G[f"cin_dfilter_{i}"])
        _close(variables[f"cin_bias_{i}"].grad, G[f"cin_dbias_{i}"])


# ----------------------------------------------------------------------------------------------- fused DeepFM tower
GT = np.load(os.path.join(os.path.dirname(__file__), "golden", "tower_v1.npz"))


def test_oracle_reproduces_golden_tower():
    """tests/golden/tower_v1.npz is what the oracle computes today (any edit to oracle/ that moves a DeepFM logit,
    gradient or update fails here)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "make_golden_tower", os.path.join(os.path.dirname(__file__), "golden", "make_golden_tower.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    d = mod.inputs()
    for k_, v in d.items():
        assert np.array_equal(GT[f"in_{k_}"], v.numpy()), k_
    for k_, v in mod.expected(d).items():
        np.testing.assert_allclose(v, GT[f"out_{k_}"], rtol=1e-12, atol=1e-14, err_msg=k_)


@pytest.mark.gpu
def test_tower_kernels_against_golden():
    """rm_tower_fwd -> rm_deepfm_head -> rm_tower_plan -> rm_tower_bwd_update (GD step) through the C ABI against the
    committed fp64 vectors: logits, loss and every gradient within 1e-5, updated tables within 1e-5."""
    from recman_b200 import _C, ops

    I = lambda n, dt=None: (torch.from_numpy(GT[f"in_{n}"]) if dt is None else torch.from_numpy(GT[f"in_{n}"]).to(dt)).cuda()
    O = lambda n: torch.from_numpy(GT[f"out_{n}"])
    K, B, ND, N1, N2 = (int(v) for v in GT["meta"])
    sizes = GT["sizes"]
    m = len(sizes)
    total = int(sizes.sum())
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)).cuda()
    lr = float(GT["lr"][0])

    def close(name, got, exp, scale=1e-5):
        exp = exp.double()
        torch.testing.assert_close(got.detach().cpu().double().reshape(exp.shape), exp, rtol=1e-5,
                                   atol=scale * max(float(exp.abs().max()), 1e-30), msg=lambda m_: f"{name}: {m_}")

    table, scal = I("table").clone(), I("scal").clone()
    st = ops.new_status("cuda")
    y1, fm, lin, S, _ = ops.tower_fwd(table, scal, offs, I("ids"), I("dense"), I("lin_dense"), I("W1"), I("b1"), status=st)
    for n, got in (("y1", y1), ("fm", fm), ("lin", lin), ("S", S)):
        close(n, got, O(n))
    out = ops.deepfm_head(y1, fm, lin, I("w0"), I("W2"), I("b2"), I("w3").reshape(-1).contiguous(), I("b3"), I("y"),
                          _C.ACT_KINDS["leaky_relu"], 0, dense=I("dense"))
    close("logit", out["logit"], O("logit"))
    close("loss", out["loss"], O("loss"))
    for n, key in (("g1", "g1"), ("g", "g"), ("dW2", "dW2"), ("db2", "db2"), ("dw3", "dw3"), ("db3", "dscal"),
                   ("dw0", "dscal"), ("db1", "db1"), ("dlin_dense", "dlin_dense")):
        close(n, out[key], O(n))
    close("dW1[dense rows]", out["dW1_dense"], O("dW1")[m * K :])
    # backward over the sorted plan: gradients only, then the in-kernel GD update
    plan = ops.tower_plan(I("ids"), offs, total, status=st)
    keys = plan.sorted_keys.cpu().numpy().view(np.uint32).astype(np.int64)
    closing = np.flatnonzero(np.append(keys[1:] != keys[:-1], True))
    t0, s0 = table.clone(), scal.clone()
    dW1, rows, sc = ops.tower_bwd_update(t0, s0, plan, out["g1"], S, out["g"], out["g"], I("W1"), _C.OPT_KINDS["gd"], lr,
                                         update=False, debug=True, status=st)
    close("dW1[embedding rows]", dW1, O("dW1")[: m * K])
    close("summed gradient rows", rows.cpu()[closing], O("d_table")[keys[closing]])
    close("k=1 gradients", sc.cpu()[closing], O("d_scal")[keys[closing]])
    ops.tower_bwd_update(table, scal, plan, out["g1"], S, out["g"], out["g"], I("W1"), _C.OPT_KINDS["gd"], lr, status=st)
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    close("table after the GD step", table, O("table_gd"))
    close("(bias, weight) after the GD step", scal, O("scal_gd"))
    untouched = np.setdiff1d(np.arange(total), keys)
    assert torch.equal(table.cpu()[untouched], torch.from_numpy(GT["in_table"])[untouched])
