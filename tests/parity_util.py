"""Model-level parity helpers: run a recman.th model on the GPU and the same weights/inputs through the
CPU oracle, return both sides' logits, loss and gradients.  Test infrastructure (imports ``oracle``)."""

from __future__ import annotations

import numpy as np
import torch

import oracle


def make_feat_dict(sizes, n_dense, multi_tags=None):
    from recman_b200.th.input import DenseFeat, FeatureDictionary, MultiValCsvFeat, SparseFeat

    fd = FeatureDictionary()
    for i, v in enumerate(sizes):
        fd[f"C{i}"] = SparseFeat(f"C{i}", v - 1, encoder=False)  # feat_size = v (row 0 = unknown)
    for j in range(n_dense):
        fd[f"I{j}"] = DenseFeat(f"I{j}", scaler=False)
    if multi_tags:
        fd["tags"] = MultiValCsvFeat("tags", tags=multi_tags)
    return fd


def synth_batch(fd, B, seed=0, multi=None):
    """Random ids / dense values / labels as a dict of numpy columns (what DataInputs.load consumes)."""
    rng = np.random.RandomState(seed)
    X = {}
    for f in fd.sparse_feats:
        X[f.name] = rng.randint(0, f.feat_size, size=B)
    for f in fd.dense_feats:
        X[f.name] = rng.randn(B).astype(np.float32)
    for f in fd.multi_val_csv_feats:
        tags = list(f.tags) + ["zz"]  # "zz" is unknown -> id 0
        col = []
        for _ in range(B):
            n = rng.randint(0, 4)
            col.append("|".join(rng.choice(tags, size=n)) if n else "")
        X[f.name] = np.array(col, dtype=object)
    y = (rng.rand(B) < 0.3).astype(np.float32)
    return X, y


def randomize_variables(model, seed=1, scale=0.05):
    """Replace the (deterministic, partly zero) initial values by seeded random ones - weights are injected,
    never re-derived from TF's RNG stream."""
    g = torch.Generator().manual_seed(seed)
    for name, p in model.variables.items():
        p.data.copy_((torch.randn(p.shape, generator=g) * scale).to(p.device))


def cpu_state(model, dtype=torch.float64):
    return {k: v.detach().cpu().to(dtype).requires_grad_() for k, v in model.variables.items()}


def _split_tables(model, st, layer):
    tabs, biases = [], []
    T = st[layer.table_name]
    Bt = st.get(layer.bias_name)
    for f, lo in zip(layer.feats, layer.row_offsets):
        tabs.append(T[lo : lo + f.feat_size])
        if Bt is not None:
            biases.append(Bt[lo : lo + f.feat_size].reshape(-1, 1))
    return tabs, (biases if Bt is not None else None)


def _oracle_inputs(fd, layer, X):
    out = []
    for f in layer.feats:
        if f.kind == "multi":
            v, o = f(X[f.name])
            out.append((torch.from_numpy(v), torch.from_numpy(o)))
        else:
            out.append(torch.from_numpy(np.asarray(X[f.name]).astype(np.int64)))
    return out


def _oracle_linear(fd, linear, st, X, dtype):
    feats = linear.linear_feats
    kinds, inputs, sizes = [], [], []
    for f in feats:
        sizes.append(f.feat_size)
        if f.kind == "sparse":
            kinds.append("sparse")
            inputs.append(torch.from_numpy(np.asarray(X[f.name]).astype(np.int64)))
        elif f.kind == "multi":
            kinds.append("multi")
            v, o = f(X[f.name])
            inputs.append((torch.from_numpy(v), torch.from_numpy(o)))
        else:
            kinds.append("dense")
            inputs.append(torch.from_numpy(np.asarray(X[f.name], dtype=np.float32)).to(dtype))
    p = linear.prefix
    return oracle.sparse_linear(st[f"{p}linear_w"], st[f"{p}linear_w0"], sizes, inputs, kinds)


def _dnn_params(st, n_layers, prefix=""):
    return (
        [st[f"{prefix}dnn_layer_{i}_weights"] for i in range(n_layers)],
        [st[f"{prefix}dnn_layer_{i}_bias"] for i in range(n_layers)],
        st[f"{prefix}dnn_w"],
        st[f"{prefix}dnn_w0"],
    )


def _act(fn):
    from recman_b200.th import layers as L

    if fn is L.leaky_relu or fn == "leaky_relu":
        return oracle.leaky_relu_tf
    if fn is L.relu or fn == "relu":
        return oracle.relu
    raise ValueError(fn)


def oracle_loss(model, X, y, dtype=torch.float64):
    """Forward the oracle with the model's current weights.  Returns (state, logit [B,1], loss scalar)."""
    from recman_b200.th.DCN import DCN
    from recman_b200.th.DeepFM import DeepFM
    from recman_b200.th.xDeepFM import xDeepFM

    st = cpu_state(model, dtype)
    fd = model.feat_dict
    hp = model.hparams
    layer = model.embeddings
    tabs, biases = _split_tables(model, st, layer)
    embeds, bias = oracle.feat_embedding_layer(tabs, _oracle_inputs(fd, layer, X), biases)
    dense = None
    if fd.dense_feats:
        dense = torch.stack([torch.from_numpy(np.asarray(X[f.name], dtype=np.float32)).to(dtype) for f in fd.dense_feats], 1)
    yt = torch.from_numpy(np.array(y, dtype=np.float32)).to(dtype)  # np.array: a writable copy
    emb_l2 = hp.get("embedding_l2_reg", 0.0) * sum(oracle.l2_loss(t) for t in tabs) if layer.l2_reg else 0.0
    if isinstance(model, DeepFM):
        lin = _oracle_linear(fd, model.linear, st, X, dtype)
        nl = len(hp["deep_hidden_units"])
        dnn_p = _dnn_params(st, nl) if model.use_deep else None
        logit = oracle.deepfm_logit(embeds, bias, lin, dense, dnn_p, _act(hp["deep_activation"]),
                                    model.use_fm, model.use_deep)
        l2 = emb_l2 + hp["linear_l2_reg"] * oracle.l2_loss(st["linear_w"])
        if model.use_deep:
            l2 = l2 + hp["deep_l2_reg"] * (sum(oracle.l2_loss(st[f"dnn_layer_{i}_weights"]) for i in range(nl))
                                           + oracle.l2_loss(st["dnn_w"]))
    elif isinstance(model, DCN):
        lin = _oracle_linear(fd, model.linear, st, X, dtype) if model.use_linear else None
        nl = len(hp["deep_hidden_units"])
        cross = (st["cross_weights"], st["cross_bias"], st["cross_w"], st["cross_w0"])
        logit = oracle.dcn_logit(embeds, lin, dense, _dnn_params(st, nl), cross, _act(hp["deep_activation"]))
        l2 = emb_l2 + hp["deep_l2_reg"] * (sum(oracle.l2_loss(st[f"dnn_layer_{i}_weights"]) for i in range(nl))
                                           + oracle.l2_loss(st["dnn_w"]))
        if model.use_linear:
            l2 = l2 + hp["linear_l2_reg"] * oracle.l2_loss(st["linear_w"])
        if hp["cross_layer_l2_reg"]:
            l2 = l2 + hp["cross_layer_l2_reg"] * (oracle.l2_loss(st["cross_weights"]) + oracle.l2_loss(st["cross_w"]))
    elif isinstance(model, xDeepFM):
        lin = _oracle_linear(fd, model.linear, st, X, dtype)
        nl = len(hp["deep_hidden_units"])
        nc = len(hp["cin_cross_layer_units"])
        cin_p = ([st[f"cin_filter_{i}"] for i in range(nc)], [st[f"cin_bias_{i}"] for i in range(nc)], st["cin_w"], st["cin_w0"])
        logit = oracle.xdeepfm_logit(embeds, lin, dense, _dnn_params(st, nl), cin_p, _act(hp["deep_activation"]),
                                     _act(hp["cin_activation"]))
        l2 = (emb_l2 + hp["linear_l2_reg"] * oracle.l2_loss(st["linear_w"])
              + hp["deep_l2_reg"] * (sum(oracle.l2_loss(st[f"dnn_layer_{i}_weights"]) for i in range(nl)) + oracle.l2_loss(st["dnn_w"]))
              + hp["cin_l2_reg"] * (sum(oracle.l2_loss(st[f"cin_filter_{i}"]) for i in range(nc)) + oracle.l2_loss(st["cin_w"])))
    else:
        raise TypeError(type(model))
    pred = oracle.prediction(logit, model.task)
    loss = oracle.create_loss(yt, pred, model.task) + l2
    return st, logit, loss


def model_grads(model):
    """Dense view of every gradient the last backward produced (host syncs; tests only)."""
    from recman_b200.autograd import dense_table_grad

    out = {}
    for name, p in model.variables.items():
        sparse = getattr(p, "rm_sparse_grads", None) or []
        tail = getattr(p, "rm_dense_tail", None)
        if sparse or tail is not None:
            g = dense_table_grad(p)
            if tail is not None:
                first, gt = tail
                g.reshape(-1)[first:] += gt
            out[name] = g.detach().cpu()
        elif p.grad is not None:
            out[name] = p.grad.detach().cpu()
        else:
            out[name] = torch.zeros_like(p).cpu()
    return out


def run_model_step(model, X, y):
    """GPU forward+backward (no optimizer).  Returns (logit, loss, grads)."""
    from recman_b200.th.input import DataInputs

    for p in model.variables.values():
        p.grad = None
        p.rm_sparse_grads = []
        p.rm_dense_tail = None
    inputs = DataInputs("cuda").load(model.feat_dict, X, y)
    loss = model._loss(inputs)
    loss.backward()
    model.check_ids()
    return model.final_logit.detach().cpu(), loss.detach().cpu(), model_grads(model)


def compare(model, X, y, rtol=1e-5, atol_scale=2e-6, grad_atol_scale=None):
    """Initialise lazily-created variables, inject random weights, compare GPU vs fp64 oracle."""
    from recman_b200.th.input import DataInputs

    with torch.no_grad():
        model._out(DataInputs("cuda").load(model.feat_dict, X, y))  # creates the variables
    randomize_variables(model)
    logit, loss, grads = run_model_step(model, X, y)
    st, o_logit, o_loss = oracle_loss(model, X, y, torch.float64)
    o_loss.backward()
    report = {}

    def close(name, got, exp, scale, floor=1.0):
        got, exp = got.double(), exp.double()
        atol = scale * max(floor, float(exp.abs().max())) if exp.numel() else 0.0
        err = float((got - exp).abs().max()) if exp.numel() else 0.0
        report[name] = err
        torch.testing.assert_close(got, exp, rtol=rtol, atol=atol, msg=lambda m: f"{name}: {m}")

    close("logit", logit.reshape(-1), o_logit.detach().reshape(-1), atol_scale)
    close("loss", loss.reshape(()), o_loss.detach().reshape(()), atol_scale)
    # gradients: 1e-5 of the tensor's own largest entry (they are O(1/B), an absolute floor would hide errors)
    gs = grad_atol_scale if grad_atol_scale is not None else 1e-5
    for name, g in grads.items():
        exp = st[name].grad
        exp = torch.zeros_like(st[name]) if exp is None else exp
        close(f"grad:{name}", g.reshape(-1), exp.reshape(-1), gs, floor=1e-30)
    return report
