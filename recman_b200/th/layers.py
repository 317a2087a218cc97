"""recman.th.layers - the layer vocabulary of recman/tf/core/layers.py on hand-written sm_100a kernels.

Every class keeps the reference constructor (shared ``variables`` dict first, then the same arguments in
the same order), is a callable returning tensors, and has ``.l2()``.  ``variables`` maps the reference's
variable names to ``torch.nn.Parameter`` s on the GPU.  Two deliberate layout differences, both invisible
through ``DeepModel.state_dict()`` / ``load_state_dict()``:

* all per-field embedding tables live in ONE parameter ``{prefix}feat_embed_table`` ``[sum V_f, k]`` (and
  ``{prefix}feat_bias_table`` ``[sum V_f]``) so that one kernel launch serves every field; the reference's
  ``{prefix}{feat}_feat_embed`` names are row-range views of it;
* embedding gradients are ``ops.SparseGrad`` (unique rows + summed rows), never dense.

``CrossNet`` does not exist in the reference (call site only, DCN.py:135-137); its parameter names follow
the ``dnn_*`` pattern: ``cross_layer_{i}_weights``/``_bias`` are rows of ``cross_weights``/``cross_bias``
``[L, d]``, head ``cross_w`` ``[d,1]``, ``cross_w0`` ``[1]``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _C, ops
from ..autograd import (
    CINLayerFunction, CINLayerPoolFunction,
    CrossFunction,
    EmbeddingLayerFunction,
    EmbeddingLayout,
    FirstLinearFunction,
    FMFunction,
    NarrowLinearFunction,
    MultiField,
    SparseRun,
)
from .input import DenseFeat, MultiValCsvFeat, SparseFeat

__all__ = [
    "FeatEmbedding",
    "FeatEmbeddingLayer",
    "LinearCombiner",
    "LinearLayer",
    "SparseLinearCombiner",
    "SparseLinearLayer",
    "FMLayer",
    "DNNCombiner",
    "DNN",
    "CIN",
    "CrossNet",
    "PredictionLayer",
    "BatchNormalization",
    "PaddedRows",
    "set_dropout_mask_source",
    "glorot_normal",
    "glorot_uniform",
    "leaky_relu",
    "relu",
    "activation_kind",
]

DEVICE = "cuda"


# --------------------------------------------------------------------------- #
# initialisers (tf/core/utils.py:156-189)
# --------------------------------------------------------------------------- #
def calc_fan(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    if len(shape) in (3, 4):
        ks = int(np.prod(shape[:-2]))
        return shape[-2] * ks, shape[-1] * ks
    raise ValueError()


def glorot_normal(shape, gain=1.0, seed=2019, device=DEVICE, out=None):
    """Truncated normal (+-2 std), std = gain*sqrt(2/(fan_in+fan_out)).  As in the reference every call
    re-seeds with the same seed, so equal-shaped tensors start identical (layers.py:99-101,536,551-553)."""
    fan_in, fan_out = calc_fan(shape)
    std = gain * math.sqrt(2.0 / (fan_in + fan_out))
    t = torch.empty(*shape, dtype=torch.float32, device=device) if out is None else out
    g = torch.Generator(device=t.device).manual_seed(int(seed))
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=g)
    return t


def glorot_uniform(shape, gain=1.0, seed=None, device=DEVICE):
    fan_in, fan_out = calc_fan(shape)
    b = gain * math.sqrt(6.0 / (fan_in + fan_out))
    g = None
    if seed is not None:
        g = torch.Generator(device=device).manual_seed(int(seed))
    return (torch.rand(*shape, dtype=torch.float32, device=device, generator=g) * 2 - 1) * b


def leaky_relu(x):
    """tf.nn.leaky_relu: slope 0.2 (not torch's 0.01)."""
    return torch.nn.functional.leaky_relu(x, 0.2)


def relu(x):
    return torch.relu(x)


def activation_kind(fn) -> int:
    if fn is None:
        return _C.ACT_IDENTITY
    if isinstance(fn, str):
        return _C.ACT_KINDS[fn]
    if fn is leaky_relu:
        return _C.ACT_LEAKY_RELU
    if fn in (relu, torch.relu, torch.nn.functional.relu):
        return _C.ACT_RELU
    raise ValueError("CIN kernels implement identity / relu / leaky_relu(0.2); got %r" % (fn,))


def resolve_activation(fn):
    if isinstance(fn, str):
        return {"leaky_relu": leaky_relu, "relu": relu, "identity": (lambda x: x), "linear": (lambda x: x),
                "sigmoid": torch.sigmoid, "tanh": torch.tanh}[fn]
    return fn


def _param(t: torch.Tensor) -> torch.nn.Parameter:
    return torch.nn.Parameter(t, requires_grad=True)


def _l2(w: torch.Tensor) -> torch.Tensor:
    return (w * w).sum() / 2  # tf.nn.l2_loss


_dropout_mask_source = None  # callable(shape, keep_prob, device) -> 0/1 tensor | None (tests: injected Bernoulli masks)


def set_dropout_mask_source(fn):
    """Inject the Bernoulli keep-masks of every dropout site, in call order (``None`` restores the device RNG).  TF's
    random stream cannot be reproduced, so parity tests of the reference's default keep-probs (hparams/xDeepFM.py:28:
    0.8) feed the same masks to this path and to the oracle (oracle.tf_dropout)."""
    global _dropout_mask_source
    _dropout_mask_source = fn


def _dropout(x, keep_prob, training=True):
    """tf.nn.dropout(x, rate=1-keep_prob): the reference's tuples are KEEP probabilities."""
    if keep_prob >= 1 or not training:
        return x
    if _dropout_mask_source is not None:
        mask = _dropout_mask_source(tuple(x.shape), float(keep_prob), x.device)
        if mask is not None:
            return x * mask.to(x.dtype) / float(keep_prob)
    return torch.nn.functional.dropout(x, p=1.0 - float(keep_prob), training=True)


class PaddedRows:
    """The front end's row buffer: ``buf`` [B, ld] with 16-byte aligned rows, first ``d`` columns valid."""

    def __init__(self, buf: torch.Tensor, d: int):
        self.buf = buf
        self.d = d

    @property
    def shape(self):
        return (self.buf.shape[0], self.d)

    def view(self):
        return self.buf[:, : self.d]


# --------------------------------------------------------------------------- #
# embeddings (A1-A4)
# --------------------------------------------------------------------------- #
class FeatEmbeddingLayer:
    """All feature embeddings (layers.py:196-267): ``__call__(inputs) -> (embeds [B,m,k], bias [B,m,1] | None)``."""

    display_name = "FeatEmbeddingLayer"

    def __init__(self, variables, feat_dict, embedding_size, l2_reg=0.00001, use_bias=True, prefix="", seed=2019):
        self.variables = variables
        self.feat_dict = feat_dict
        self.embedding_size = embedding_size
        self.l2_reg = l2_reg
        self.use_bias = use_bias
        self.prefix = prefix
        self.seed = seed
        self.feats = list(feat_dict.embedding_feats)
        for f in self.feats:
            assert not isinstance(f, DenseFeat)
            if not isinstance(f, (SparseFeat, MultiValCsvFeat)) or getattr(f, "kind", "") == "sparse_value":
                raise NotImplementedError(f"{type(f).__name__} has no working lookup in the reference")
        sizes = [f.feat_size for f in self.feats]
        self.row_offsets = [0] + list(np.cumsum(sizes))
        self.total_rows = int(self.row_offsets[-1])
        self.status = None
        self._layout = None
        self.shard = None  # th.dist.ShardPlan: this rank keeps ceil(V_f / W) rows per table

    def set_shard(self, plan):
        """Row-sharded mode: the parameters hold only this rank's rows (row r -> rank r % W, local r // W)."""
        self.shard = plan
        self.row_offsets = list(plan.local_offsets)
        self.total_rows = int(plan.total_local)

    # names -------------------------------------------------------------------------------------------
    @property
    def table_name(self):
        return f"{self.prefix}feat_embed_table"

    @property
    def bias_name(self):
        return f"{self.prefix}feat_bias_table"

    def _upsert_variables(self):
        k = self.embedding_size
        if self.table_name not in self.variables:
            if self.shard is not None:
                table = self.shard.alloc((self.total_rows, k), zero=True)  # cudaIpc-shared in peer-memory mode
            else:
                table = torch.empty(self.total_rows, k, dtype=torch.float32, device=DEVICE)
            for i, (f, lo) in enumerate(zip(self.feats, self.row_offsets)):
                rows = f.feat_size if self.shard is None else self.shard.local_sizes[i]
                # fans come from the full table shape so that the init scale does not depend on the world size
                std_shape = [f.feat_size, k]
                fan_in, fan_out = calc_fan(std_shape)
                std = math.sqrt(2.0 / (fan_in + fan_out))
                g = torch.Generator(device=table.device).manual_seed(int(self.seed) + (0 if self.shard is None else self.shard.rank))
                torch.nn.init.trunc_normal_(table[lo : lo + rows], mean=0.0, std=std, a=-2 * std, b=2 * std, generator=g)
            self.variables[self.table_name] = _param(table)
        if self.use_bias and self.bias_name not in self.variables:
            bias = (self.shard.alloc((self.total_rows,)) if self.shard is not None
                    else torch.zeros(self.total_rows, dtype=torch.float32, device=DEVICE))
            self.variables[self.bias_name] = _param(bias)

    def feature_tables(self) -> Dict[str, torch.Tensor]:
        """Reference-named views: ``{prefix}{feat}_feat_embed`` [V_f, k], ``{prefix}{feat}_feat_bias`` [V_f, 1]."""
        self._upsert_variables()
        out = {}
        table = self.variables[self.table_name].data
        for f, lo in zip(self.feats, self.row_offsets):
            out[f"{self.prefix}{f.name}_feat_embed"] = table[lo : lo + f.feat_size]
            if self.use_bias:
                out[f"{self.prefix}{f.name}_feat_bias"] = self.variables[self.bias_name].data[lo : lo + f.feat_size].view(-1, 1)
        return out

    def layout(self) -> EmbeddingLayout:
        if self._layout is not None:
            return self._layout
        lay = EmbeddingLayout(m=len(self.feats), k=self.embedding_size, total_rows=self.total_rows)
        sparse_names = [f.name for f in self.feat_dict.sparse_feats]
        multi_names = [f.name for f in self.feat_dict.multi_val_csv_feats]
        col = 0
        while col < len(self.feats):
            f = self.feats[col]
            if isinstance(f, MultiValCsvFeat):
                lay.multi.append(MultiField(col, int(self.row_offsets[col]), f.feat_size, multi_names.index(f.name)))
                col += 1
                continue
            start = col
            id_col = sparse_names.index(f.name)
            while (col + 1 < len(self.feats) and type(self.feats[col + 1]) is SparseFeat
                   and sparse_names.index(self.feats[col + 1].name) == id_col + (col + 1 - start)):
                col += 1
            n = col - start + 1
            offs = torch.tensor(self.row_offsets[start : start + n + 1], dtype=torch.int64, device=DEVICE)
            lay.runs.append(SparseRun(start, n, id_col, offs))
            col += 1
        self._layout = lay
        return lay

    @property
    def all_one_hot(self) -> bool:
        lay = self.layout()
        return len(lay.multi) == 0 and len(lay.runs) == 1

    def __call__(self, inputs):
        self._upsert_variables()
        table = self.variables[self.table_name]
        bias_table = self.variables[self.bias_name] if self.use_bias else None
        lay = self.layout()
        csr: List[torch.Tensor] = []
        for f in self.feat_dict.multi_val_csv_feats:
            v, o = inputs[f.name]
            csr += [v, o]
        sparse_ids = getattr(inputs, "sparse_ids", None)
        if sparse_ids is None and lay.runs:
            sparse_ids = torch.cat([inputs[f.name].reshape(-1, 1) for f in self.feat_dict.sparse_feats], dim=1)
        if self.status is None:
            self.status = ops.new_status(table.device)
        res = EmbeddingLayerFunction.apply(table, bias_table, lay, self.status, sparse_ids, *csr)
        if self.use_bias:
            return res[0], res[1]
        return res, None

    def l2(self):
        """l2_reg * sum_f ||E_f||^2 / 2 over the ENTIRE tables (layers.py:188-193, 263-267)."""
        self._upsert_variables()
        if not self.l2_reg:
            return torch.zeros((), device=DEVICE)
        return self.l2_reg * _l2(self.variables[self.table_name])


class FeatEmbedding:
    """Single-feature lookup (layers.py:68-193), kept for API parity; wraps a one-feature layer."""

    display_name = "FeatEmbedding"

    def __init__(self, variables, feat, embedding_size, l2_reg=0.00001, use_bias=True, prefix="", seed=2019):
        from .input import FeatureDictionary

        assert not isinstance(feat, DenseFeat)
        fd = FeatureDictionary()
        fd[feat.name] = feat
        self.feat = feat
        self._layer = FeatEmbeddingLayer(variables, fd, embedding_size, l2_reg, use_bias, prefix=f"{prefix}{feat.name}_", seed=seed)

    def __call__(self, feat_input):
        inputs = {self.feat.name: feat_input}
        return self._layer(inputs)

    def l2(self):
        return self._layer.l2()


# --------------------------------------------------------------------------- #
# first-order linear term (A9)
# --------------------------------------------------------------------------- #
class _LinearInput:
    def __init__(self, inputs, linear_feats):
        self.inputs = inputs
        self.linear_feats = linear_feats


class LinearCombiner:
    """(layers.py:270-298).  The reference materialises a [B, sum V] one-hot matrix; here the combiner
    only records which inputs feed the linear term - the layer gathers ``linear_w`` rows directly."""

    display_name = "LinearCombiner"

    def __init__(self, linear_feats, prefix=""):
        self.linear_feats = list(linear_feats.values()) if isinstance(linear_feats, dict) else list(linear_feats)
        self.prefix = prefix

    def __call__(self, inputs):
        self.output = _LinearInput(inputs, self.linear_feats)
        return self.output


class SparseLinearCombiner(LinearCombiner):
    display_name = "SparseLinearCombiner"


class LinearLayer:
    """(layers.py:301-354 / 389-446): ``sum_f linear_w[off_f + id_f] + dense . w + linear_w0`` -> [B,1].

    ``linear_w`` is the reference's ``[sum(feat_size), 1]`` variable in ``linear_feats`` order.  Sparse fields
    use the k=1 gather (K1) and its deterministic scatter-add backward (K2); multi-valued fields contribute
    ``sum_tags w[tag]`` with tag 0 (unknown) dropped, duplicates counted (tf/core/utils.py:86-110).
    With ``training=False`` the per-id ``feat.weights`` are added to ``linear_w`` (layers.py:338-345).
    """

    display_name = "LinearRegression"

    def __init__(self, variables, linear_feats, l2_reg=0.00001, prefix="", training=True):
        self.variables = variables
        self.linear_feats = list(linear_feats.values()) if isinstance(linear_feats, dict) else list(linear_feats)
        self.l2_reg = l2_reg
        self.prefix = prefix
        self.training = training
        sizes = [f.feat_size for f in self.linear_feats]
        self.offsets = [0] + list(np.cumsum(sizes))
        self.total = int(self.offsets[-1])
        self.alloc = None  # th.dist.ShardPlan.alloc when linear_w's id rows are row-sharded

    def _upsert_variables(self):
        name = f"{self.prefix}linear_w0"
        if name not in self.variables:
            self.variables[name] = _param(torch.zeros(1, dtype=torch.float32, device=DEVICE))
        name = f"{self.prefix}linear_w"
        if name not in self.variables:
            w = self.alloc((self.total, 1)) if self.alloc is not None else torch.zeros(
                self.total, 1, dtype=torch.float32, device=DEVICE)
            self.variables[name] = _param(w)

    def effective_weight(self):
        W = self.variables[f"{self.prefix}linear_w"]
        if not self.training:
            extra = np.concatenate([np.asarray(f.weights, dtype=np.float32).reshape(-1) for f in self.linear_feats])
            if np.any(extra != 0):
                W = W + torch.from_numpy(extra).to(W.device).reshape(-1, 1)
        return W

    def __call__(self, lin_input: _LinearInput):
        self._upsert_variables()
        inputs = lin_input.inputs
        W = self.effective_weight()
        W0 = self.variables[f"{self.prefix}linear_w0"]
        logit = None
        # one-hot fields: one k=1 gather over all of them
        sp = [(i, f) for i, f in enumerate(self.linear_feats) if type(f) is SparseFeat]
        if sp:
            from .input import FeatureDictionary

            fd = FeatureDictionary((f.name, f) for _, f in sp)
            # k=1 "embedding layer" over a table that is a view of linear_w: reuse K1/K2 through a 1-column table
            ids = torch.cat([inputs[f.name].reshape(-1, 1) for _, f in sp], dim=1).contiguous()
            rows = [self.offsets[i] for i, _ in sp]
            contiguous_block = all(rows[j] + sp[j][1].feat_size == rows[j + 1] for j in range(len(sp) - 1))
            if contiguous_block:
                lo, hi = rows[0], rows[-1] + sp[-1][1].feat_size
                offs = torch.tensor([r - lo for r in rows] + [hi - lo], dtype=torch.int64, device=W.device)
                logit = _LinearGather.apply(W, lo, hi, offs, ids).sum(dim=1, keepdim=True)
            else:
                for (i, f), col in zip(sp, range(len(sp))):
                    lo, hi = self.offsets[i], self.offsets[i] + f.feat_size
                    offs = torch.tensor([0, hi - lo], dtype=torch.int64, device=W.device)
                    part = _LinearGather.apply(W, lo, hi, offs, ids[:, col : col + 1].contiguous())
                    logit = part if logit is None else logit + part
        for i, f in enumerate(self.linear_feats):
            lo = self.offsets[i]
            if isinstance(f, MultiValCsvFeat):
                values, offsets = inputs[f.name]
                B = offsets.numel() - 1
                w = W[lo : lo + f.feat_size, 0]
                contrib = torch.where(values > 0, w[values], torch.zeros((), device=w.device))
                sample = torch.repeat_interleave(torch.arange(B, device=w.device), offsets[1:] - offsets[:-1])
                part = torch.zeros(B, device=w.device).index_add(0, sample, contrib).reshape(-1, 1)
                logit = part if logit is None else logit + part
            elif isinstance(f, DenseFeat):
                part = inputs[f.name].reshape(-1, 1).to(torch.float32) * W[lo, 0]
                logit = part if logit is None else logit + part
        return logit + W0

    def l2(self):
        self._upsert_variables()
        if not self.l2_reg:  # no graph through linear_w: its gradient stays sparse
            return torch.zeros((), device=DEVICE)
        return self.l2_reg * _l2(self.variables[f"{self.prefix}linear_w"])


class SparseLinearLayer(LinearLayer):
    display_name = "SparseLinearLayer"


class _LinearGather(torch.autograd.Function):
    """k=1 gather of rows [lo, hi) of ``linear_w`` for a block of one-hot fields -> [B, n_fields]."""

    @staticmethod
    def forward(ctx, W, lo, hi, offs, ids):
        table = W.detach()[lo:hi].reshape(-1, 1)
        out = ops.gather(table, offs, ids).reshape(ids.shape[0], ids.shape[1])
        ctx.save_for_backward(offs, ids)
        ctx.lo, ctx.hi, ctx.total = lo, hi, W.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        offs, ids = ctx.saved_tensors
        plan = ops.segment_plan(ids, offs, ctx.hi - ctx.lo)
        rows = ops.segment_reduce(g.contiguous(), plan, 1, ld=ids.shape[1])
        sg = ops.SparseGrad(plan.uniq_rows, rows.reshape(-1), plan.n_unique)
        dW = torch.zeros(ctx.total, 1, dtype=torch.float32, device=g.device)
        dW[ctx.lo : ctx.hi] = sg.to_dense(ctx.hi - ctx.lo)  # linear_w is small here; the fused front end keeps it sparse
        return dW, None, None, None, None


# --------------------------------------------------------------------------- #
# FM (A5)
# --------------------------------------------------------------------------- #
class FMLayer:
    """Factorization-machine layer (layers.py:449-481): ``(embeds [B,m,k], bias [B,m,1]) -> [B,1]``."""

    def __init__(self, dropout=(1, 1)):
        self.dropout = dropout
        self.training = True

    def __call__(self, embeddings, embedding_bias):
        assert embeddings.dim() == 3
        embedding_bias = _dropout(embedding_bias, self.dropout[0], self.training) if embedding_bias is not None else None
        embeddings = _dropout(embeddings, self.dropout[1], self.training)
        return FMFunction.apply(embeddings, embedding_bias)

    def l2(self):
        return torch.zeros((), device=DEVICE)


# --------------------------------------------------------------------------- #
# DNN (A8) - GEMMs stay on cuBLAS through torch (not in the hand-written set)
# --------------------------------------------------------------------------- #
class DNNCombiner:
    """(layers.py:484-501): flatten + concat.  The model front ends produce the combined row directly."""

    display_name = "DNNCombiner"

    def __init__(self, prefix=""):
        self.prefix = prefix

    def __call__(self, inputs: list):
        self.result = torch.cat([t.reshape(t.shape[0], -1).to(torch.float32) for t in inputs], dim=1)
        return self.result


def compute_hidden_units_s2(num_hidden_layers, input_neurons, output_neurons=1):
    return [round((input_neurons + output_neurons) * 2 / 3) for _ in range(num_hidden_layers)]


class DNN:
    """(layers.py:504-628).  ``inputs`` is a tensor [B,d] or a ``PaddedRows``."""

    display_name = "DeepNeuralNetwork"

    def __init__(self, variables, hidden_units, dropout, activation, l2_reg=0.00001, prefix="", seed=2019):
        assert len(hidden_units) > 0
        assert len(hidden_units) + 1 == len(dropout)
        self.variables = variables
        self.hidden_units = list(hidden_units)
        self.dropout = dropout
        self.activation = resolve_activation(activation)
        self.l2_reg = l2_reg
        self.prefix = prefix
        self.seed = seed
        self.training = True

    def _create_weights(self, d_in):
        v, p = self.variables, self.prefix
        dims = [d_in] + self.hidden_units
        for i in range(len(self.hidden_units)):
            if f"{p}dnn_layer_{i}_weights" not in v:
                v[f"{p}dnn_layer_{i}_weights"] = _param(glorot_normal([dims[i], dims[i + 1]], seed=self.seed))
            if f"{p}dnn_layer_{i}_bias" not in v:
                v[f"{p}dnn_layer_{i}_bias"] = _param(torch.zeros(dims[i + 1], dtype=torch.float32, device=DEVICE))
        if f"{p}dnn_w" not in v:
            v[f"{p}dnn_w"] = _param(glorot_normal([self.hidden_units[-1], 1], seed=self.seed))
        if f"{p}dnn_w0" not in v:
            v[f"{p}dnn_w0"] = _param(torch.zeros(1, dtype=torch.float32, device=DEVICE))

    def first_layer(self, d_in):
        """(W, b) of the first layer, created on demand: the fused DeepFM tower computes ``x @ W + b`` itself."""
        if any(u is None for u in self.hidden_units):
            self.hidden_units = compute_hidden_units_s2(len(self.hidden_units), d_in)
        self._create_weights(d_in)
        return self.variables[f"{self.prefix}dnn_layer_0_weights"], self.variables[f"{self.prefix}dnn_layer_0_bias"]

    def from_first_layer(self, y1):
        """The rest of the network given the first layer's pre-activation ``y1`` [B, hidden_units[0]]."""
        v, p = self.variables, self.prefix
        y = _dropout(self.activation(y1), self.dropout[1], self.training)
        for i in range(1, len(self.hidden_units)):
            W, b = v[f"{p}dnn_layer_{i}_weights"], v[f"{p}dnn_layer_{i}_bias"]
            if ops.narrow_linear_ok(W.shape[1]) and y.shape[1] % 4 == 0:
                y = NarrowLinearFunction.apply(y.contiguous(), W, b)
            else:
                y = torch.addmm(b, y, W)
            y = self.activation(y)
            y = _dropout(y, self.dropout[i + 1], self.training)
        return torch.addmm(v[f"{p}dnn_w0"], y, v[f"{p}dnn_w"])

    def __call__(self, inputs):
        d_in = inputs.shape[1]
        if any(u is None for u in self.hidden_units):
            self.hidden_units = compute_hidden_units_s2(len(self.hidden_units), d_in)
        self._create_weights(d_in)
        v, p = self.variables, self.prefix
        padded = isinstance(inputs, PaddedRows)
        if padded and (self.dropout[0] < 1 and self.training):
            inputs, padded = inputs.view(), False
        y = inputs if padded else _dropout(inputs, self.dropout[0], self.training)
        for i in range(len(self.hidden_units)):
            W, b = v[f"{p}dnn_layer_{i}_weights"], v[f"{p}dnn_layer_{i}_bias"]
            if i == 0 and padded:
                y = FirstLinearFunction.apply(y.buf, W, b, y.d, getattr(y, "fm_back", None))
            elif ops.narrow_linear_ok(W.shape[1]) and y.shape[1] % 4 == 0:
                # narrow hidden layer: batch-reduced backward kernels (csrc/mlp.cu)
                y = NarrowLinearFunction.apply(y.contiguous(), W, b)
            else:
                y = torch.addmm(b, y, W)
            y = self.activation(y)
            y = _dropout(y, self.dropout[i + 1], self.training)
        return torch.addmm(v[f"{p}dnn_w0"], y, v[f"{p}dnn_w"])

    def l2(self):
        v, p = self.variables, self.prefix
        if not self.l2_reg:
            return torch.zeros((), device=DEVICE)
        terms = [self.l2_reg * _l2(v[f"{p}dnn_layer_{i}_weights"]) for i in range(len(self.hidden_units))]
        terms.append(self.l2_reg * _l2(v[f"{p}dnn_w"]))
        return torch.stack(terms).sum()


# --------------------------------------------------------------------------- #
# CIN (A7)
# --------------------------------------------------------------------------- #
class CIN:
    """Compressed Interaction Network (layers.py:631-777): ``inputs [B,m,D] -> [B,1]``.

    ``precision``: "3xtf32" (tcgen05, parity), "tf32" (tcgen05 single pass, fast), "fp32" (CUDA cores).
    """

    display_name = "CompressedInteractionNetwork"

    def __init__(self, variables, cross_layer_units, activation, dropout, l2_reg=0.00001, prefix="", seed=2019,
                 precision="3xtf32"):
        self.variables = variables
        self.cross_layer_units = list(cross_layer_units)
        self.activation = activation
        self.dropout = dropout
        self.l2_reg = l2_reg
        self.prefix = prefix
        self.seed = seed
        self.precision = {"fp32": _C.CIN_FP32_SIMT, "3xtf32": _C.CIN_3XTF32, "tf32": _C.CIN_TF32}[precision] \
            if isinstance(precision, str) else int(precision)
        self.training = True
        assert len(self.cross_layer_units) > 0
        assert len(self.cross_layer_units) + 1 == len(self.dropout)

    def _upsert_variables(self, field_size):
        v, p = self.variables, self.prefix
        field_nums = [field_size]
        final_size = 0
        for i, size in enumerate(self.cross_layer_units):
            if f"{p}cin_filter_{i}" not in v:
                v[f"{p}cin_filter_{i}"] = _param(glorot_normal([1, field_nums[-1] * field_nums[0], size], seed=self.seed))
            if f"{p}cin_bias_{i}" not in v:
                v[f"{p}cin_bias_{i}"] = _param(torch.zeros(size, dtype=torch.float32, device=DEVICE))
            field_nums.append(size // 2)
            final_size += field_nums[-1] if i != len(self.cross_layer_units) - 1 else size
        if f"{p}cin_w" not in v:
            v[f"{p}cin_w"] = _param(glorot_uniform([final_size, 1], seed=int(self.seed) + 1))  # seeded: equal on every rank
        if f"{p}cin_w0" not in v:
            v[f"{p}cin_w0"] = _param(torch.zeros(1, dtype=torch.float32, device=DEVICE))

    def __call__(self, inputs):
        assert inputs.dim() == 3
        B, field_size, D = inputs.shape
        self._upsert_variables(field_size)
        v, p = self.variables, self.prefix
        act = activation_kind(self.activation)
        x0 = _dropout(inputs, self.dropout[0], self.training)
        xk = x0
        finals = []
        n_layers = len(self.cross_layer_units)
        for i, size in enumerate(self.cross_layer_units):
            W = v[f"{p}cin_filter_{i}"][0]  # [m*H_i, N_i]
            last = i == n_layers - 1
            if not self.training or self.dropout[i + 1] is None or self.dropout[i + 1] >= 1:
                # no dropout between the layer and its consumers: split-half + sum-pool ride with the layer kernel
                assert last or size % 2 == 0, "tf.split needs an even layer size"
                xk, pooled = CINLayerPoolFunction.apply(x0, xk, W, v[f"{p}cin_bias_{i}"], act, self.precision,
                                                        0 if last else size // 2)
                finals.append(pooled)
                continue
            feat_map = CINLayerFunction.apply(x0, xk, W, v[f"{p}cin_bias_{i}"], act, self.precision)  # [B,N,D]
            feat_map = _dropout(feat_map, self.dropout[i + 1], self.training)
            if i != n_layers - 1:
                half = size // 2
                assert size == 2 * half, "tf.split needs an even layer size"
                xk, direct = feat_map[:, :half], feat_map[:, half:]  # FIRST half feeds the next layer
            else:
                direct = feat_map
            finals.append(direct.sum(dim=-1))  # sum-pool over D
        result = torch.cat(finals, dim=1)  # [B, sum H]
        return torch.addmm(v[f"{p}cin_w0"], result, v[f"{p}cin_w"])

    def l2(self):
        v, p = self.variables, self.prefix
        if not self.l2_reg:
            return torch.zeros((), device=DEVICE)
        terms = [self.l2_reg * _l2(v[f"{p}cin_filter_{i}"]) for i in range(len(self.cross_layer_units))]
        terms.append(self.l2_reg * _l2(v[f"{p}cin_w"]))
        return torch.stack(terms).sum()


# --------------------------------------------------------------------------- #
# CrossNet (A6)
# --------------------------------------------------------------------------- #
class CrossNet:
    """DCN cross network; call-site contract of DCN.py:135-137,166: ``CrossNet(cross_layer_num,
    cross_layer_l2_reg)(dnn_input) -> logit [B,1]``, ``.l2()``.  The optional ``variables``/``prefix``/``seed``
    keywords follow the other layers."""

    display_name = "CrossNet"

    def __init__(self, cross_layer_num, cross_layer_l2_reg=0.0, variables=None, prefix="", seed=2019):
        self.cross_layer_num = cross_layer_num
        self.l2_reg = cross_layer_l2_reg
        self.variables = variables if variables is not None else {}
        self.prefix = prefix
        self.seed = seed

    @property
    def weights(self):
        p = self.prefix
        return {n: self.variables[n] for n in (f"{p}cross_weights", f"{p}cross_bias", f"{p}cross_w", f"{p}cross_w0")
                if n in self.variables}

    def _upsert_variables(self, d):
        v, p, L = self.variables, self.prefix, self.cross_layer_num
        if f"{p}cross_weights" not in v:
            w = torch.empty(L, d, dtype=torch.float32, device=DEVICE)
            for l in range(L):
                glorot_normal([d, 1], seed=self.seed, out=w[l].view(d, 1))
            v[f"{p}cross_weights"] = _param(w)
        if f"{p}cross_bias" not in v:
            v[f"{p}cross_bias"] = _param(torch.zeros(L, d, dtype=torch.float32, device=DEVICE))
        if f"{p}cross_w" not in v:
            v[f"{p}cross_w"] = _param(glorot_normal([d, 1], seed=self.seed))
        if f"{p}cross_w0" not in v:
            v[f"{p}cross_w0"] = _param(torch.zeros(1, dtype=torch.float32, device=DEVICE))

    def __call__(self, inputs):
        d = inputs.shape[1]
        self._upsert_variables(d)
        v, p = self.variables, self.prefix
        args = (v[f"{p}cross_weights"], v[f"{p}cross_bias"], v[f"{p}cross_w"], v[f"{p}cross_w0"])
        if isinstance(inputs, PaddedRows):
            return CrossFunction.apply(inputs.buf, *args, inputs.d)
        return CrossFunction.apply(inputs, *args, None)

    def l2(self):
        v, p = self.variables, self.prefix
        if not self.l2_reg:
            return torch.zeros((), device=DEVICE)
        return self.l2_reg * (_l2(v[f"{p}cross_weights"]) + _l2(v[f"{p}cross_w"]))


# --------------------------------------------------------------------------- #
# prediction / batch norm
# --------------------------------------------------------------------------- #
class PredictionLayer:
    """(layers.py:780-808): optional global bias, sigmoid for classification, reshape(-1)."""

    display_name = "Prediction"

    def __init__(self, variables, task="classification", use_bias=False, prefix=""):
        self.task = task
        self.use_bias = use_bias
        self.prefix = prefix
        self.variables = variables

    def __call__(self, inputs):
        output = inputs
        if self.use_bias:
            name = f"{self.prefix}global_bias"
            if name not in self.variables:
                self.variables[name] = _param(torch.zeros(1, dtype=torch.float32, device=DEVICE))
            output = output + self.variables[name]
        if self.task == "classification":
            output = torch.sigmoid(output)
        return output.reshape(-1)


class BatchNormalization:
    """(layers.py:26-65): batch statistics only (no moving averages), unused by the live model."""

    def __init__(self, epsilon=1e-3, prefix=""):
        self.epsilon = epsilon
        self.prefix = prefix
        self.weights = {}

    def __call__(self, inputs):
        assert inputs.dim() == 2
        self.units = inputs.shape[1]
        if not self.weights:
            self.weights = {
                f"{self.prefix}scale": _param(torch.ones(self.units, device=inputs.device)),
                f"{self.prefix}beta": _param(torch.zeros(self.units, device=inputs.device)),
            }
        mean = inputs.mean(0)
        var = inputs.var(0, unbiased=False)
        return (inputs - mean) * torch.rsqrt(var + self.epsilon) * self.weights[f"{self.prefix}scale"] + \
            self.weights[f"{self.prefix}beta"]

    @property
    def output_shape(self):
        return -1, self.units
