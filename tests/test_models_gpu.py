"""Model-level parity on the GPU: recman.th DeepFM / DCN (/ xDeepFM in test_cin_gpu.py) vs the fp64 CPU oracle -
logits, loss and every gradient, with injected weights; fused and unfused front ends; optimizer step."""
import numpy as np
import pytest
import torch

import oracle
from tests import parity_util as pu

pytestmark = pytest.mark.gpu

CRITEO_SMALL = [50, 7, 1000, 3, 200, 31, 2, 90]


@pytest.mark.parametrize("k", [8, 16, 64])
@pytest.mark.parametrize("B", [256, 1000])
def test_deepfm_fused_parity(k, B):
    from recman_b200.th import DeepFM

    fd = pu.make_feat_dict(CRITEO_SMALL, n_dense=13)
    X, y = pu.synth_batch(fd, B, seed=k + B)
    model = DeepFM(fd, embedding_size=k, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=B)
    rep = pu.compare(model, X, y)
    assert "grad:feat_embed_table" in rep and "grad:feat_bias_table" in rep and "grad:linear_w" in rep


def test_deepfm_unfused_with_multival_ml100k_shape():
    """C1 shape: 5 sparse + 2 dense + 1 multi-valued field, k=8, B=256 -> unfused kernels (K1 pooled, K3, K2)."""
    from recman_b200.th import DeepFM

    fd = pu.make_feat_dict([944, 1683, 3, 22, 796], n_dense=2, multi_tags=[f"g{i}" for i in range(19)])
    X, y = pu.synth_batch(fd, 256, seed=3)
    model = DeepFM(fd, embedding_size=8, deep_dropout=(1, 1, 1), batch_size=256)
    pu.compare(model, X, y)


def test_deepfm_k_not_multiple_of_4_uses_scalar_kernels():
    from recman_b200.th import DeepFM

    fd = pu.make_feat_dict([20, 30, 5], n_dense=1)
    X, y = pu.synth_batch(fd, 100, seed=4)
    model = DeepFM(fd, embedding_size=6, deep_dropout=(1, 1, 1), batch_size=100)
    pu.compare(model, X, y)


@pytest.mark.parametrize("use_fm,use_deep", [(True, False), (False, True)])
def test_deepfm_towers(use_fm, use_deep):
    from recman_b200.th import DeepFM

    fd = pu.make_feat_dict(CRITEO_SMALL, n_dense=3)
    X, y = pu.synth_batch(fd, 300, seed=5)
    model = DeepFM(fd, embedding_size=16, deep_dropout=(1, 1, 1), use_fm=use_fm, use_deep=use_deep, batch_size=300)
    pu.compare(model, X, y)


@pytest.mark.parametrize("k,L", [(16, 6), (8, 3)])
def test_dcn_parity(k, L):
    from recman_b200.th import DCN

    fd = pu.make_feat_dict(CRITEO_SMALL * 2, n_dense=13)
    X, y = pu.synth_batch(fd, 512, seed=6)
    model = DCN(fd, embedding_size=k, deep_hidden_units=(64, 64, 64), deep_dropout=(1, 1, 1, 1), cross_layer_num=L,
                cross_layer_l2_reg=1e-5, deep_l2_reg=1e-5, batch_size=512)
    pu.compare(model, X, y)


@pytest.mark.parametrize("opt", ["adam", "adagrad", "gd"])
@pytest.mark.parametrize("l2_mode,emb_l2", [("dense", 1e-5), ("touched", 1e-5), ("dense", 0.0)])
def test_fit_on_batch_matches_oracle_update(opt, l2_mode, emb_l2):
    """One full step (forward, backward, fresh optimizer) against the oracle's update rule on dense gradients."""
    from recman_b200.th import DeepFM
    from recman_b200.th.input import DataInputs

    fd = pu.make_feat_dict([40, 9, 300], n_dense=2)
    X, y = pu.synth_batch(fd, 128, seed=7)
    lr = 0.01
    lin_l2 = 1e-5 if l2_mode == "dense" else 0.0
    model = DeepFM(fd, embedding_size=8, deep_dropout=(1, 1, 1), batch_size=128, optimizer=opt, learning_rate=lr,
                   embedding_l2_reg=emb_l2, embedding_l2_mode=l2_mode, linear_l2_reg=lin_l2)
    with torch.no_grad():
        model._out(DataInputs("cuda").load(fd, X, y))
    pu.randomize_variables(model)
    st, _, o_loss = pu.oracle_loss(model, X, y, torch.float64)
    if l2_mode == "touched":  # oracle_loss sees l2_reg == 0 on the layer; the touched-rows term is added below
        assert model.embeddings.l2_reg == 0.0
    o_loss.backward()
    before = {k: v.detach().clone() for k, v in st.items()}
    model.fit_on_batch(X, y)
    for name, p in model.variables.items():
        g = st[name].grad if st[name].grad is not None else torch.zeros_like(st[name])
        if l2_mode == "touched" and name == "feat_embed_table" and emb_l2:
            touched = g.abs().sum(1) != 0
            g = g + emb_l2 * before[name] * touched[:, None]
        exp = oracle.fresh_optimizer_step(before[name], g, opt, lr)
        got = p.detach().cpu().double()
        if opt == "adam":
            # fresh Adam is sign-like (lr*g/(|g|+3e-6)): where |g| is at rounding-noise level the step is
            # ill-conditioned, so those entries are only required to stay within one step of the old value
            solid = g.abs() > 1e-4 * g.abs().max()
            assert torch.all((got - before[name]).abs() <= lr * 1.0001 + 1e-12), name
            torch.testing.assert_close(got[solid], exp[solid], rtol=1e-5, atol=lr * 2e-3, msg=lambda m: f"{name}: {m}")
        else:
            torch.testing.assert_close(got, exp, rtol=1e-5, atol=1e-7, msg=lambda m: f"{name}: {m}")


def test_predict_and_evaluate_batches():
    from recman_b200.th import DeepFM
    from recman_b200.th.metric import LogLoss, RocAucScore

    fd = pu.make_feat_dict([30, 12], n_dense=1)
    X, y = pu.synth_batch(fd, 130, seed=8)
    model = DeepFM(fd, embedding_size=8, batch_size=64, epoch=1, eval_metric=(LogLoss(), RocAucScore()),
                   learning_rate=0.01)
    p = model.predict(X)
    assert p.shape == (130,) and np.all((p > 0) & (p < 1))
    p2 = model.predict({k: v[:128] for k, v in X.items()})  # 128 % 64 == 0: the reference's empty trailing batch
    assert p2.shape == (128,)
    np.testing.assert_array_equal(p[:128], p2)
    # the encode-once + prefetch path (N2) returns what per-batch encoding returns
    from recman_b200.th.DeepModel import DeepModel

    orig = DeepModel._encode_once
    DeepModel._encode_once = lambda self, X_, y_: None
    try:
        np.testing.assert_array_equal(model.predict(X), p)
    finally:
        DeepModel._encode_once = orig
    model.fit(X, y, random_seed_for_mini_batch=False)
    res = model.evaluate(X, y)
    assert len(res) == 2 and all(np.isfinite(r) for r in res)
    assert len(model.history) == 2


@pytest.mark.parametrize("model_name", ["DeepFM", "DCN"])
def test_cuda_graph_step_equals_eager(model_name):
    """compile_step: the captured fwd+bwd+K2+optimizer step replays to the same parameters as eager steps."""
    from recman_b200 import th
    from recman_b200.th.input import DataInputs

    fd = pu.make_feat_dict(CRITEO_SMALL, n_dense=13)
    batches = [pu.synth_batch(fd, 256, seed=40 + i) for i in range(4)]
    kw = dict(embedding_size=16, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=256, learning_rate=0.01,
              embedding_l2_reg=0.0, linear_l2_reg=0.0, optimizer="adagrad")
    first = DataInputs("cuda").load(fd, *batches[0])
    # compile_step runs `warmup` real steps on `first` (capture itself only records): mirror that eagerly
    eager2 = getattr(th, model_name)(fd, **kw)
    graph2 = getattr(th, model_name)(fd, **kw)
    for mdl in (eager2, graph2):
        with torch.no_grad():
            mdl._out(DataInputs("cuda").load(fd, *batches[0]))
        pu.randomize_variables(mdl, seed=5)
    eager2.fit_on_batch(first, None)      # the warm-up step
    graph2.compile_step(first, warmup=1)
    for X, y in batches[1:]:
        le = eager2.fit_on_batch(X, y)
        lg = graph2.fit_on_batch(X, y)
        torch.testing.assert_close(lg, le, rtol=1e-6, atol=1e-7)
    for name in eager2.variables:
        torch.testing.assert_close(graph2.variables[name].data, eager2.variables[name].data, rtol=1e-6, atol=1e-7,
                                   msg=lambda m_: f"{name}: {m_}")
    graph2.check_ids()


@pytest.mark.parametrize("model_name", ["DeepFM", "DCN"])
@pytest.mark.parametrize("opt", ["adam", "adagrad"])
def test_fused_sparse_update_is_bit_identical(model_name, opt):
    """N1: the optimizer update applied inside the backward kernel (rm_emb_fm_bwd_update) == K2 then rm_sparse_opt_step."""
    from recman_b200 import th

    fd = pu.make_feat_dict(CRITEO_SMALL, n_dense=13)
    X, y = pu.synth_batch(fd, 700, seed=21)  # duplicates guaranteed (tables of 2..1000 rows)
    models = []
    for fuse in (True, False):
        kw = dict(embedding_size=16, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=700,
                  embedding_l2_reg=0.0, linear_l2_reg=0.0, optimizer=opt, learning_rate=0.01)
        if model_name == "DCN":
            kw["cross_layer_num"] = 3
        model = getattr(th, model_name)(fd, **kw)
        model.hparams["fuse_sparse_update"] = fuse
        from recman_b200.th.input import DataInputs

        with torch.no_grad():
            model._out(DataInputs("cuda").load(fd, X, y))
        pu.randomize_variables(model, seed=3)
        for _ in range(2):
            model.fit_on_batch(X, y)
        torch.cuda.synchronize()
        models.append(model)
    a, b = models
    assert set(a.variables) == set(b.variables)
    for name in a.variables:
        assert torch.equal(a.variables[name].data, b.variables[name].data), name


def test_deepfm_fm_backward_fused_into_first_layer_matches_separate_kernels():
    """DeepFM with a narrow first layer, opt-in hparams["fuse_fm_backward"]: the FM backward rides in the MLP
    input-gradient epilogue (rm_linear_bwd_input_fm -> row gradients -> reduce).  Same gradients as the separate
    dx / fused-reduce kernels, and the path is actually taken."""
    import torch

    from recman_b200 import ops
    from recman_b200.th import DeepFM

    fd = pu.make_feat_dict([50, 7, 1000, 3, 200, 31], n_dense=5)
    X, y = pu.synth_batch(fd, 700, seed=3)
    grads = {}
    for fuse in (True, False):
        model = DeepFM(fd, embedding_size=16, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=700)
        model.hparams["fuse_fm_backward"] = fuse
        calls = []
        orig = ops.linear_bwd_input_fm
        ops.linear_bwd_input_fm = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        try:
            report = pu.compare(model, X, y)
        finally:
            ops.linear_bwd_input_fm = orig
        assert bool(calls) == fuse
        assert max(report.values()) < 1e-4
        _, _, g = pu.run_model_step(model, X, y)
        grads[fuse] = g
    for name in grads[True]:
        torch.testing.assert_close(grads[True][name], grads[False][name], rtol=1e-5, atol=1e-7)


def test_host_prefetcher_delivers_every_batch_in_order():
    """Double-buffered H2D (N2): get(i) returns batch i on the device while batch i+1 is already being copied."""
    from recman_b200.th.input import HostPrefetcher

    fd = pu.make_feat_dict([50, 7, 1000], n_dense=2)
    g = torch.Generator().manual_seed(0)
    host = []
    for i in range(5):
        ids = torch.randint(0, 7, (64, 3), generator=g).pin_memory()
        dense = torch.randn(64, 2, generator=g).pin_memory()
        y = torch.rand(64, generator=g).pin_memory()
        host.append((ids, dense, y))
    pf = HostPrefetcher(fd, lambda i: host[i % 5], "cuda")
    for i in list(range(7)) + [3, 4]:  # sequential use, then a jump (re-stages)
        inp = pf.get(i)
        a, b, c = host[i % 5]
        torch.cuda.synchronize()
        assert torch.equal(inp.sparse_ids.cpu(), a) and torch.equal(inp.dense.cpu(), b) and torch.equal(inp["y"].cpu(), c)
        assert pf._pending[0] == i + 1
    assert pf.h2d_bytes > 0
    # no dense features
    fd2 = pu.make_feat_dict([50, 7, 1000], n_dense=0)
    pf2 = HostPrefetcher(fd2, lambda i: (host[i % 5][0], None, host[i % 5][2]), "cuda")
    inp = pf2.get(0)
    assert inp.dense is None and torch.equal(inp.sparse_ids.cpu(), host[0][0])


def test_fit_encode_once_prefetch_path_equals_per_batch_path():
    """fit(): encoding the frame once + pinned slices + HostPrefetcher (N2) trains exactly like re-encoding every batch."""
    from recman_b200.th import DeepFM
    from recman_b200.th.DeepModel import DeepModel

    fd = pu.make_feat_dict([50, 7, 1000, 3], n_dense=3)
    X, y = pu.synth_batch(fd, 1000, seed=11)  # 1000 = 3 full batches of 300 + a ragged one
    kw = dict(embedding_size=8, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=300, epoch=2,
              learning_rate=0.01, embedding_l2_reg=0.0, linear_l2_reg=0.0)
    fast = DeepFM(fd, **kw)
    fast.fit(X, y, random_seed_for_mini_batch=False)
    slow = DeepFM(fd, **kw)
    orig = DeepModel._encode_once
    DeepModel._encode_once = lambda self, X_, y_: None
    try:
        slow.fit(X, y, random_seed_for_mini_batch=False)
    finally:
        DeepModel._encode_once = orig
    assert fast.samples_seen == slow.samples_seen == 2000
    for name, p in fast.variables.items():
        assert torch.equal(p.data, slow.variables[name].data), name


# ----------------------------------------------------------------------------------------------------------------------
# fused DeepFM tower (rm_tower_fwd / rm_tower_bwd_update): k = 64, first hidden layer 32
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("opt", ["adam", "adagrad", "gd"])
@pytest.mark.parametrize("B", [128, 777])
def test_tower_fit_on_batch_matches_oracle_update(opt, B):
    """One full step through the fused tower (forward kernel, sorted backward + in-kernel update) against the oracle's
    gradients + fresh-optimizer rule, every variable."""
    from recman_b200.th import DeepFM
    from recman_b200.th.input import DataInputs

    fd = pu.make_feat_dict([40, 9, 300, 5000, 3], n_dense=13)
    X, y = pu.synth_batch(fd, B, seed=17)
    lr = 0.05 if opt == "gd" else 0.01
    model = DeepFM(fd, embedding_size=64, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=B, optimizer=opt,
                   learning_rate=lr, embedding_l2_reg=0.0, linear_l2_reg=0.0)
    with torch.no_grad():
        model._out(DataInputs("cuda").load(fd, X, y))
    assert "scal_storage" in model._cache, "the tower path did not engage"
    pu.randomize_variables(model)
    st, _, o_loss = pu.oracle_loss(model, X, y, torch.float64)
    o_loss.backward()
    before = {k: v.detach().clone() for k, v in st.items()}
    from recman_b200 import ops

    n0 = ops.launch_count()
    model.fit_on_batch(X, y)
    torch.cuda.synchronize()
    model.check_ids()
    assert ops.launch_count() > n0
    for name, p in model.variables.items():
        g = st[name].grad if st[name].grad is not None else torch.zeros_like(st[name])
        exp = oracle.fresh_optimizer_step(before[name], g, opt, lr)
        got = p.detach().cpu().double()
        if opt == "adam":
            solid = g.abs() > 1e-4 * g.abs().max()
            assert torch.all((got - before[name]).abs() <= lr * 1.0001 + 1e-12), name
            torch.testing.assert_close(got[solid], exp[solid], rtol=1e-5, atol=lr * 2e-3, msg=lambda m: f"{name}: {m}")
        else:
            torch.testing.assert_close(got, exp, rtol=1e-5, atol=1e-7, msg=lambda m: f"{name}: {m}")


def test_tower_matches_separate_kernels_and_graph_replay():
    """Three DeepFM models with the same weights: fused tower (eager), fused tower (CUDA graph) and the separate
    kernels (hparams tower=False).  Same parameters after several steps (1e-5: the tensor-core GEMMs are 3xTF32)."""
    from recman_b200.th import DeepFM
    from recman_b200.th.input import DataInputs

    fd = pu.make_feat_dict(CRITEO_SMALL, n_dense=13)
    batches = [pu.synth_batch(fd, 512, seed=60 + i) for i in range(4)]
    kw = dict(embedding_size=64, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=512, learning_rate=0.01,
              embedding_l2_reg=0.0, linear_l2_reg=0.0, optimizer="adagrad")
    first = DataInputs("cuda").load(fd, *batches[0])
    models = []
    for tower in (True, True, False):
        mdl = DeepFM(fd, **kw)
        mdl.hparams["tower"] = tower
        with torch.no_grad():
            mdl._out(first)
        models.append(mdl)
    eager, graph, plain = models
    assert "scal_storage" in eager._cache and "scal_storage" not in plain._cache
    pu.randomize_variables(eager, seed=5)
    for other in (graph, plain):
        for name, p in eager.variables.items():
            other.variables[name].data.copy_(p.data)
    eager.fit_on_batch(first, None)
    plain.fit_on_batch(first, None)
    graph.compile_step(first, warmup=1)
    for X, y in batches[1:]:
        le = eager.fit_on_batch(X, y)
        lg = graph.fit_on_batch(X, y)
        lp = plain.fit_on_batch(X, y)
        torch.testing.assert_close(lg, le, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(lp, le, rtol=1e-5, atol=1e-6)
    for name in eager.variables:
        a = eager.variables[name].data
        torch.testing.assert_close(graph.variables[name].data, a, rtol=1e-6, atol=1e-7, msg=lambda m_: f"{name}: {m_}")
        scale = float(a.abs().max())
        torch.testing.assert_close(plain.variables[name].data, a, rtol=1e-5, atol=2e-6 * max(scale, 1e-3),
                                   msg=lambda m_: f"{name}: {m_}")
    eager.check_ids()
    graph.check_ids()


def test_tower_inference_weight_override():
    """Inference-time per-id weights (layers.py:426-437, examples/xDeepFM_test.py:124): added to linear_w when
    training=False - through the fused tower forward and through the separate kernels, against the oracle."""
    from recman_b200.th import DeepFM
    from recman_b200.th.input import DataInputs

    fd = pu.make_feat_dict([40, 9, 300], n_dense=2)
    X, y = pu.synth_batch(fd, 200, seed=3)
    outs = {}
    first = None
    for tower in (True, False):
        model = DeepFM(fd, embedding_size=64, deep_hidden_units=(32, 32), deep_dropout=(1, 1, 1), batch_size=200)
        model.hparams["tower"] = tower
        inputs = DataInputs("cuda").load(fd, X, y)
        with torch.no_grad():
            model._out(inputs)
        if first is None:
            pu.randomize_variables(model, seed=9)
            first = model
        else:  # same weights by name (the two layouts create their variables in a different order)
            for name, p in first.variables.items():
                model.variables[name].data.copy_(p.data)
        with torch.no_grad():
            base = model._out(inputs, training=False).cpu()
        fd["C1"].set_weights({3: -5.0})  # the reference's set_weights({"Outdoor": -5}) with ids for keys
        try:
            with torch.no_grad():
                pred = model._out(inputs, training=False).cpu()
                pred_train = model._out(inputs, training=True).cpu()  # the override is inference-only
        finally:
            fd["C1"].set_weights(None)
        hit = torch.from_numpy(np.asarray(X["C1"]) == 3)
        assert hit.any()
        logit = lambda p: torch.log(p.double() / (1 - p.double()))
        torch.testing.assert_close(logit(pred)[hit], logit(base)[hit] - 5.0, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(pred[~hit], base[~hit], rtol=0, atol=0)
        torch.testing.assert_close(pred_train, base, rtol=1e-6, atol=1e-7)
        outs[tower] = pred
    torch.testing.assert_close(outs[True], outs[False], rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------------------------------------------------
# the reference's default keep-probabilities (dropout 0.8) with injected masks
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [16, 64])
def test_deepfm_reference_default_dropout_with_injected_masks(k):
    """deep_dropout = (0.8, 0.8, 0.8) (tf/core/DeepFM.py:38, hparams/xDeepFM.py:28): the same Bernoulli masks go to the
    GPU path (layers.set_dropout_mask_source) and to the oracle's DNN (oracle.dnn(masks=...)); logit, loss and every
    gradient at 1e-5."""
    from recman_b200.th import DeepFM, layers
    from recman_b200.th.input import DataInputs

    fd = pu.make_feat_dict([40, 9, 300, 7], n_dense=3)
    B = 333
    X, y = pu.synth_batch(fd, B, seed=23)
    keep = (0.8, 0.8, 0.8)
    model = DeepFM(fd, embedding_size=k, deep_hidden_units=(32, 32), deep_dropout=keep, batch_size=B)
    with torch.no_grad():
        model._out(DataInputs("cuda").load(fd, X, y), training=False)
    pu.randomize_variables(model)
    d = 4 * k + 3
    g = torch.Generator().manual_seed(99)
    masks = [(torch.rand(B, n, generator=g) < p).float() for n, p in zip((d, 32, 32), keep)]
    calls = []

    def source(shape, keep_prob, device):
        m = masks[len(calls)]
        calls.append(shape)
        assert tuple(shape) == tuple(m.shape) and keep_prob == 0.8
        return m.to(device)

    layers.set_dropout_mask_source(source)
    try:
        logit, loss, grads = pu.run_model_step(model, X, y)
    finally:
        layers.set_dropout_mask_source(None)
    assert len(calls) == 3
    # oracle with the same masks
    st = pu.cpu_state(model, torch.float64)
    layer = model.embeddings
    tabs, biases = pu._split_tables(model, st, layer)
    embeds, bias = oracle.feat_embedding_layer(tabs, pu._oracle_inputs(fd, layer, X), biases)
    dense = torch.stack([torch.from_numpy(np.asarray(X[f.name], dtype=np.float32)).double() for f in fd.dense_feats], 1)
    lin = pu._oracle_linear(fd, model.linear, st, X, torch.float64)
    xin = oracle.dnn_combiner([embeds, dense])
    dnn_l = oracle.dnn(xin, *pu._dnn_params(st, 2), activation=oracle.relu, dropout=keep, masks=[m.double() for m in masks])
    o_logit = lin + oracle.fm_layer(embeds, bias) + dnn_l
    hp = model.hparams
    l2 = (hp["embedding_l2_reg"] * sum(oracle.l2_loss(t) for t in tabs) + hp["linear_l2_reg"] * oracle.l2_loss(st["linear_w"])
          + hp["deep_l2_reg"] * (oracle.l2_loss(st["dnn_layer_0_weights"]) + oracle.l2_loss(st["dnn_layer_1_weights"])
                                 + oracle.l2_loss(st["dnn_w"])))
    o_loss = oracle.create_loss(torch.from_numpy(np.array(y, dtype=np.float32)).double(), oracle.prediction(o_logit), "classification") + l2
    o_loss.backward()
    torch.testing.assert_close(logit.reshape(-1).double(), o_logit.detach().reshape(-1), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(loss.double().reshape(()), o_loss.detach().reshape(()), rtol=1e-5, atol=2e-6)
    for name, gg in grads.items():
        exp = st[name].grad if st[name].grad is not None else torch.zeros_like(st[name])
        atol = 1e-5 * max(float(exp.abs().max()), 1e-30)
        torch.testing.assert_close(gg.double().reshape(-1), exp.reshape(-1), rtol=1e-5, atol=atol, msg=lambda m: f"{name}: {m}")
