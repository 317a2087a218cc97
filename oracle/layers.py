"""torch-CPU restatement of recman's layer arithmetic (oracle; test infrastructure).

Every function is a pure function of injected weights (never of seeds: TF's
truncated-normal stream is not reproducible in torch) and works in whatever
dtype its inputs carry (fp32 for parity, fp64 to attribute error).  Citations
are ``path:line`` under the reference tree ``/root/reference``.

PARITY UNPINNED - see ``oracle/__init__.py``.
"""

from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch

__all__ = [
    "leaky_relu_tf",
    "relu",
    "get_activation",
    "tf_dropout",
    "embedding_lookup",
    "embedding_lookup_sqrtn",
    "feat_embedding_layer",
    "l2_loss",
    "fm_layer",
    "cross_net",
    "cin",
    "cin_layer_shapes",
    "dnn",
    "dnn_combiner",
    "sparse_linear",
    "prediction",
    "binary_crossentropy",
    "mean_squared_error",
    "create_loss",
    "deepfm_logit",
    "dcn_logit",
    "xdeepfm_logit",
    "fresh_optimizer_step",
    "calc_fan",
    "glorot_std",
    "glorot_limit",
]


# --------------------------------------------------------------------------- #
# activations / dropout
# --------------------------------------------------------------------------- #
def leaky_relu_tf(x: torch.Tensor, alpha: float = 0.2) -> torch.Tensor:
    """``tf.nn.leaky_relu`` - default slope is 0.2, not torch's 0.01.

    It is the default CIN and DNN activation (tf/hparams/xDeepFM.py:29,33).
    TF computes ``max(alpha*x, x)``.
    """
    return torch.maximum(alpha * x, x)


def relu(x: torch.Tensor) -> torch.Tensor:
    """``tf.nn.relu`` - the DeepFM / DCN default (tf/core/DeepFM.py:39, DCN.py:35)."""
    return torch.clamp_min(x, 0)


def get_activation(name) -> Callable[[torch.Tensor], torch.Tensor]:
    if callable(name):
        return name
    table = {
        "leaky_relu": leaky_relu_tf,
        "relu": relu,
        "identity": lambda x: x,
        "linear": lambda x: x,
        "sigmoid": torch.sigmoid,
        "tanh": torch.tanh,
    }
    return table[name]


def tf_dropout(x: torch.Tensor, keep_prob: float, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``tf.nn.dropout(x, rate=1-keep_prob)`` (layers.py:461,466,589,602,707,740).

    The reference's dropout tuples are KEEP probabilities.  keep_prob == 1 is
    the identity.  Otherwise kept elements are scaled by ``1/keep_prob``; the
    Bernoulli mask must be injected (``mask`` of 0/1) because TF's RNG stream
    cannot be reproduced.
    """
    if keep_prob >= 1:
        return x
    if mask is None:
        raise ValueError("dropout with keep_prob<1 needs an injected mask")
    return x * mask.to(x.dtype) / keep_prob


# --------------------------------------------------------------------------- #
# A1 / A2 / A3 : embedding lookup
# --------------------------------------------------------------------------- #
def embedding_lookup(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """``tf.nn.embedding_lookup(table, ids[:, :1])`` (layers.py:117-128).

    table [V, k]; ids [B] or [B, 1] int64 -> [B, 1, k].  A pure row copy.
    """
    ids = ids.reshape(-1).long()
    return table[ids].unsqueeze(1)


def embedding_lookup_sqrtn(table: torch.Tensor, values: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """``tf.nn.embedding_lookup_sparse(..., combiner="sqrtn")`` (layers.py:144-169).

    The ragged ids of sample b are ``values[offsets[b]:offsets[b+1]]`` (CSR; the
    TF side builds the same thing as a SparseTensor, tf/core/utils.py:117-123).
    With ``sp_weights=None`` every weight is 1, so the result is
    ``sum_j table[id_j] / sqrt(n)``.  Rows with n == 0 yield zeros.
    Rows are accumulated in CSR order (ascending j).
    Returns [B, 1, k].
    """
    B = offsets.numel() - 1
    k = table.shape[1]
    out = torch.zeros(B, k, dtype=table.dtype)
    off = offsets.tolist()
    vals = values.long()
    for b in range(B):
        lo, hi = off[b], off[b + 1]
        n = hi - lo
        if n == 0:
            continue
        acc = torch.zeros(k, dtype=table.dtype)
        for j in range(lo, hi):
            acc = acc + table[vals[j]]
        out[b] = acc / math.sqrt(n) if table.dtype == torch.float64 else acc / torch.tensor(
            math.sqrt(n), dtype=table.dtype
        )
    return out.unsqueeze(1)


def _lookup_sqrtn_autograd(table, values, offsets):
    """Vectorised, differentiable variant of :func:`embedding_lookup_sqrtn`."""
    B = offsets.numel() - 1
    counts = (offsets[1:] - offsets[:-1]).long()
    seg = torch.repeat_interleave(torch.arange(B), counts)
    rows = table[values.long()]
    out = torch.zeros(B, table.shape[1], dtype=table.dtype).index_add(0, seg, rows)
    denom = torch.sqrt(counts.clamp_min(1).to(table.dtype)).unsqueeze(1)
    return (out / denom).unsqueeze(1)


def feat_embedding_layer(
    tables: Sequence[torch.Tensor],
    inputs: Sequence,
    bias_tables: Optional[Sequence[torch.Tensor]] = None,
):
    """``FeatEmbeddingLayer.__call__`` (layers.py:238-261).

    ``inputs[f]`` is either an int64 id tensor [B]/[B,1] (``SparseFeat`` branch)
    or a ``(values, offsets)`` CSR pair (multi-val branch, sqrtn pooling).
    Returns ``(embeds [B, m, k], bias [B, m, 1] | None)``: the per-field results
    concatenated along axis 1 in feature-dictionary order.
    """
    embeds, biases = [], []
    for f, table in enumerate(tables):
        inp = inputs[f]
        if isinstance(inp, (tuple, list)):
            values, offsets = inp
            embeds.append(_lookup_sqrtn_autograd(table, values, offsets))
            if bias_tables is not None:
                biases.append(_lookup_sqrtn_autograd(bias_tables[f], values, offsets))
        else:
            embeds.append(embedding_lookup(table, inp))
            if bias_tables is not None:
                biases.append(embedding_lookup(bias_tables[f], inp))
    e = torch.cat(embeds, dim=1)
    b = torch.cat(biases, dim=1) if bias_tables is not None else None
    return e, b


def l2_loss(w: torch.Tensor) -> torch.Tensor:
    """``tf.nn.l2_loss`` = sum(w**2) / 2 (used by every ``.l2()``, e.g. layers.py:188-193)."""
    return (w * w).sum() / 2


# --------------------------------------------------------------------------- #
# A5 : FM layer
# --------------------------------------------------------------------------- #
def fm_layer(embeddings: torch.Tensor, embedding_bias: torch.Tensor, dropout=(1, 1), masks=(None, None)) -> torch.Tensor:
    """``FMLayer.__call__`` (layers.py:457-478), op for op.

    embeddings [B, m, k], embedding_bias [B, m, 1] -> [B, 1]:
    ``sum_i bias_i + 0.5 * sum_d[(sum_i e_id)^2 - sum_i e_id^2]``.
    """
    assert embeddings.dim() == 3
    embedding_bias = tf_dropout(embedding_bias, dropout[0], masks[0])
    y_first_order = embedding_bias.sum(dim=1)  # [B, 1]
    embeddings = tf_dropout(embeddings, dropout[1], masks[1])
    sum_embeds = embeddings.sum(dim=1, keepdim=True)  # [B, 1, k]
    square_of_sum = sum_embeds * sum_embeds
    square_embeds = embeddings * embeddings
    sum_of_square = square_embeds.sum(dim=1, keepdim=True)
    y_second_order = 0.5 * (square_of_sum - sum_of_square)
    y_second_order = y_second_order.sum(dim=2)  # [B, 1]
    return y_first_order + y_second_order


# --------------------------------------------------------------------------- #
# A6 : DCN cross network (paper-defined; reference has only the call site)
# --------------------------------------------------------------------------- #
def cross_net(
    x: torch.Tensor,
    weights: torch.Tensor,
    biases: torch.Tensor,
    w_out: torch.Tensor,
    w0_out: torch.Tensor,
) -> torch.Tensor:
    """``CrossNet(cross_layer_num, l2)(dnn_input) -> logit`` (call site DCN.py:135-142).

    The reference never defines the class; the arithmetic is eq. (3) of
    arXiv 1708.05123 (cited README.md:6): ``x_{l+1} = x0 * (x_l . w_l) + b_l + x_l``.
    Because the call site adds the result to the other logits
    (DCN.py:140-142) the stack ends in a projection to one logit, written
    like ``DNN``'s ``dnn_w``/``dnn_w0`` head (layers.py:606-609).

    x [B, d]; weights, biases [L, d]; w_out [d, 1]; w0_out [1] -> [B, 1].
    """
    x0 = x
    xl = x
    for l in range(weights.shape[0]):
        s = (xl * weights[l]).sum(dim=1, keepdim=True)  # x_l^T w_l
        xl = x0 * s + biases[l] + xl
    return xl @ w_out + w0_out


# --------------------------------------------------------------------------- #
# A7 : Compressed Interaction Network
# --------------------------------------------------------------------------- #
def cin_layer_shapes(field_size: int, cross_layer_units: Sequence[int]):
    """Filter shapes ``[1, H_{i}*m, N_i]`` and the cin_w length (layers.py:659-691)."""
    field_nums = [field_size]
    final_size = 0
    shapes = []
    for i, size in enumerate(cross_layer_units):
        shapes.append((1, field_nums[-1] * field_nums[0], size))
        field_nums.append(size // 2)
        if i != len(cross_layer_units) - 1:
            final_size += field_nums[-1]
        else:
            final_size += size
    return shapes, final_size


def cin(
    inputs: torch.Tensor,
    filters: Sequence[torch.Tensor],
    biases: Sequence[torch.Tensor],
    cin_w: torch.Tensor,
    cin_w0: torch.Tensor,
    activation: Callable = leaky_relu_tf,
    dropout: Optional[Sequence[float]] = None,
    masks: Optional[Sequence] = None,
    return_pooled: bool = False,
) -> torch.Tensor:
    """``CIN.__call__`` (layers.py:697-760) following the TF op sequence.

    inputs [B, m, D].  ``filters[i]`` is ``cin_filter_i`` of shape
    ``[1, m*H_i, N_i]`` (or ``[m*H_i, N_i]``), ``biases[i]`` is ``[N_i]``.
    The flattened contraction index is X0-field-major, X_i-field-minor
    (reshape at layers.py:722-725).  Split-half (layers.py:742-749): the FIRST
    ``N_i//2`` maps feed the next layer, the SECOND half goes to the output;
    the last layer sends everything to the output.  Output ``[B, 1]``.
    """
    assert inputs.dim() == 3
    B, m, D = inputs.shape
    n_layers = len(filters)
    if dropout is None:
        dropout = [1] * (n_layers + 1)
    if masks is None:
        masks = [None] * (n_layers + 1)
    assert n_layers + 1 == len(dropout)
    inputs = tf_dropout(inputs, dropout[0], masks[0])
    hidden = [inputs]
    finals = []
    field_nums = [m]
    # tf.split(x, D*[1], axis=2) -> list of D tensors [B, H, 1]
    split0 = [inputs[:, :, d : d + 1] for d in range(D)]
    for i in range(n_layers):
        filt = filters[i]
        if filt.dim() == 3:
            filt = filt[0]
        size = filt.shape[1]
        spliti = [hidden[-1][:, :, d : d + 1] for d in range(D)]
        # tf.matmul(split_tensor_0, split_tensor, transpose_b=True): [D, B, m, H_i]
        dot_m = torch.stack([split0[d] @ spliti[d].transpose(1, 2) for d in range(D)], dim=0)
        dot_o = dot_m.reshape(D, -1, field_nums[0] * field_nums[i])
        dot = dot_o.permute(1, 0, 2)  # [B, D, m*H_i]
        feat_map = dot @ filt  # conv1d 1x1 VALID == matmul (layers.py:728-733)
        feat_map = feat_map + biases[i]
        feat_map = activation(feat_map)
        feat_map = feat_map.permute(0, 2, 1)  # [B, N_i, D]
        feat_map = tf_dropout(feat_map, dropout[i + 1], masks[i + 1])
        field_nums.append(size // 2)
        if i != n_layers - 1:
            assert size == 2 * field_nums[-1], "tf.split needs an even layer size"
            next_hidden, direct = feat_map[:, : field_nums[-1]], feat_map[:, field_nums[-1] :]
        else:
            direct, next_hidden = feat_map, None
        finals.append(direct)
        hidden.append(next_hidden)
    result = torch.cat(finals, dim=1).sum(dim=-1)  # [B, sum H]
    if return_pooled:
        return result
    return result @ cin_w + cin_w0


# --------------------------------------------------------------------------- #
# A8 : DNN
# --------------------------------------------------------------------------- #
def dnn_combiner(inputs: Sequence[torch.Tensor]) -> torch.Tensor:
    """``DNNCombiner`` (layers.py:494-501): flatten each input and concat on axis 1."""
    return torch.cat([t.reshape(t.shape[0], -1) for t in inputs], dim=1)


def dnn(
    inputs: torch.Tensor,
    weights: Sequence[torch.Tensor],
    biases: Sequence[torch.Tensor],
    dnn_w: torch.Tensor,
    dnn_w0: torch.Tensor,
    activation: Callable = leaky_relu_tf,
    dropout: Optional[Sequence[float]] = None,
    masks: Optional[Sequence] = None,
) -> torch.Tensor:
    """``DNN.__call__`` (layers.py:576-609): (matmul+bias -> act -> dropout) x n, then the 1-unit head."""
    n = len(weights)
    if dropout is None:
        dropout = [1] * (n + 1)
    if masks is None:
        masks = [None] * (n + 1)
    assert n > 0 and n + 1 == len(dropout)
    y = tf_dropout(inputs, dropout[0], masks[0])
    for i in range(n):
        y = y @ weights[i] + biases[i]
        y = activation(y)
        y = tf_dropout(y, dropout[i + 1], masks[i + 1])
    return y @ dnn_w + dnn_w0


# --------------------------------------------------------------------------- #
# A9 : first-order linear term, prediction, loss
# --------------------------------------------------------------------------- #
def sparse_linear(
    linear_w: torch.Tensor,
    linear_w0: torch.Tensor,
    feat_sizes: Sequence[int],
    inputs: Sequence,
    kinds: Sequence[str],
    extra_weights: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """``(Sparse)LinearCombiner`` + ``(Sparse)LinearLayer`` (layers.py:281-298,330-347,368-386,418-439).

    The reference materialises ``[B, sum(feat_size)]`` (one-hot per sparse
    field, tag-count vector with column 0 zeroed per multi-val field
    (tf/core/utils.py:86-110), raw value per dense field) and multiplies by
    ``linear_w [sum(feat_size), 1]``; here the same product is built explicitly
    so the CUDA k=1 gather can be checked against the matmul form.

    ``kinds[f]`` in {"sparse", "multi", "dense"}; ``inputs[f]`` is ids [B],
    a CSR ``(values, offsets)`` pair, or dense values [B].  ``extra_weights``
    is the inference-time additive override ``W + feat.weights``
    (layers.py:338-345, 426-437).
    """
    W = linear_w
    if extra_weights is not None:
        W = W + extra_weights.reshape(-1, 1).to(W.dtype)
    cols = []
    B = None
    for size, inp, kind in zip(feat_sizes, inputs, kinds):
        if kind == "sparse":
            ids = inp.reshape(-1).long()
            B = ids.numel()
            oh = torch.zeros(B, size, dtype=W.dtype)
            oh[torch.arange(B), ids] = 1
            cols.append(oh)
        elif kind == "multi":
            values, offsets = inp
            B = offsets.numel() - 1
            cnt = torch.zeros(B, size, dtype=W.dtype)
            off = offsets.tolist()
            for b in range(B):
                for j in range(off[b], off[b + 1]):
                    cnt[b, int(values[j])] += 1
            cnt[:, 0] = 0  # flatten_zeros[:, :1] (tf/core/utils.py:109-110)
            cols.append(cnt)
        elif kind == "dense":
            v = inp.reshape(-1, 1).to(W.dtype)
            B = v.shape[0]
            cols.append(v)
        else:
            raise ValueError(kind)
    X = torch.cat(cols, dim=1)
    return X @ W + linear_w0


def prediction(logit: torch.Tensor, task: str = "classification", global_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``PredictionLayer`` (layers.py:796-808): optional bias, sigmoid, reshape(-1)."""
    out = logit
    if global_bias is not None:
        out = out + global_bias
    if task == "classification":
        out = torch.sigmoid(out)
    return out.reshape(-1)


KERAS_EPSILON = 1e-7


def binary_crossentropy(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """``tf.losses.binary_crossentropy`` on probabilities (tf/core/utils.py:192-194).

    TF-2.0-era Keras backend: clip p to [eps, 1-eps] with eps = 1e-7, then
    ``-(y*log(p+eps) + (1-y)*log(1-p+eps))``, mean over the last axis (the
    whole batch here: both arguments are rank 1).
    """
    eps = KERAS_EPSILON
    y_true = y_true.to(y_pred.dtype)
    p = torch.clamp(y_pred, eps, 1 - eps)
    bce = y_true * torch.log(p + eps) + (1 - y_true) * torch.log(1 - p + eps)
    return (-bce).mean(dim=-1)


def mean_squared_error(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """``tf.losses.mean_squared_error`` (tf/core/utils.py:195-196)."""
    y_true = y_true.to(y_pred.dtype)
    return ((y_pred - y_true) ** 2).mean(dim=-1)


def create_loss(y_true, y_pred, task):
    """``create_loss`` (tf/core/utils.py:192-198)."""
    if task == "classification":
        return binary_crossentropy(y_true, y_pred)
    if task == "regression":
        return mean_squared_error(y_true, y_pred)
    raise ValueError(task)


# --------------------------------------------------------------------------- #
# model compositions
# --------------------------------------------------------------------------- #
def deepfm_logit(embeds, bias, linear_logit, dense, dnn_params, activation=relu, use_fm=True, use_deep=True):
    """DeepFM composition (tf/core/DeepFM.py:107-163): add_n([linear, fm, dnn])."""
    logit = linear_logit
    if use_fm:
        logit = logit + fm_layer(embeds, bias)
    if use_deep:
        x = dnn_combiner([embeds] + ([dense] if dense is not None else []))
        logit = logit + dnn(x, *dnn_params, activation=activation)
    return logit


def dcn_logit(embeds, linear_logit, dense, dnn_params, cross_params, activation=relu):
    """DCN composition (tf/core/DCN.py:99-149).

    ``final = add_n([dnn_logit, cn_logit, dnn_logit]) (+ linear_logit)`` - the
    reference sums ``dnn_logit`` TWICE (DCN.py:140-142); replicated.
    ``linear_logit`` may be None (``use_linear=False``).
    """
    x = dnn_combiner([embeds] + ([dense] if dense is not None else []))
    dnn_l = dnn(x, *dnn_params, activation=activation)
    cn_l = cross_net(x, *cross_params)
    logit = dnn_l + cn_l + dnn_l
    if linear_logit is not None:
        logit = logit + linear_logit
    return logit


def xdeepfm_logit(embeds, linear_logit, dense, dnn_params, cin_params, dnn_activation=leaky_relu_tf, cin_activation=leaky_relu_tf):
    """xDeepFM composition (tf/core/xDeepFM.py:47-104): linear + cin + dnn."""
    cin_l = cin(embeds, *cin_params, activation=cin_activation)
    x = dnn_combiner([embeds] + ([dense] if dense is not None else []))
    dnn_l = dnn(x, *dnn_params, activation=dnn_activation)
    return linear_logit + cin_l + dnn_l


# --------------------------------------------------------------------------- #
# optimizer (N1): the reference builds a NEW optimizer every batch
# --------------------------------------------------------------------------- #
def fresh_optimizer_step(param: torch.Tensor, grad: torch.Tensor, optimizer: str, lr: float) -> torch.Tensor:
    """One step of a freshly constructed TF optimizer (tf/core/xDeepFM.py:116-126,
    tf/core/utils.py:201-213): the reference calls ``create_optimizer`` inside
    ``fit_on_batch`` so slot variables never accumulate.  Returns the new param.

    adam (beta1=.9, beta2=.999, eps=1e-7, t=1):
        m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1),
        p -= lr_t * m / (sqrt(v) + eps)
    adagrad (initial accumulator 0.1, eps=1e-7): acc=0.1+g^2, p -= lr*g/(sqrt(acc)+eps)
    gd / momentum (fresh velocity = 0): p -= lr*g
    """
    g = grad
    if optimizer == "adam":
        b1, b2, eps = 0.9, 0.999, 1e-7
        m = (1 - b1) * g
        v = (1 - b2) * g * g
        lr_t = lr * math.sqrt(1 - b2) / (1 - b1)
        return param - lr_t * m / (torch.sqrt(v) + eps)
    if optimizer == "adagrad":
        acc = 0.1 + g * g
        return param - lr * g / (torch.sqrt(acc) + 1e-7)
    if optimizer in ("gd", "momentum", "sgd"):
        return param - lr * g
    raise ValueError(optimizer)


# --------------------------------------------------------------------------- #
# initialisers (tf/core/utils.py:156-189) - shapes/scales only, values injected
# --------------------------------------------------------------------------- #
def calc_fan(weight_shape):
    """``calc_fan`` (tf/core/utils.py:156-165)."""
    if len(weight_shape) == 2:
        fan_in, fan_out = weight_shape
    elif len(weight_shape) in (3, 4):
        in_ch, out_ch = weight_shape[-2:]
        kernel_size = 1
        for s in weight_shape[:-2]:
            kernel_size *= s
        fan_in, fan_out = in_ch * kernel_size, out_ch * kernel_size
    else:
        raise ValueError()
    return fan_in, fan_out


def glorot_std(weight_shape, gain=1.0):
    """std of ``glorot_normal`` (tf/core/utils.py:180-183); the draw is truncated at 2 std."""
    fan_in, fan_out = calc_fan(weight_shape)
    return gain * math.sqrt(2 / (fan_in + fan_out))


def glorot_limit(weight_shape, gain=1.0):
    """bound of ``glorot_uniform`` (tf/core/utils.py:186-189)."""
    fan_in, fan_out = calc_fan(weight_shape)
    return gain * math.sqrt(6 / (fan_in + fan_out))
