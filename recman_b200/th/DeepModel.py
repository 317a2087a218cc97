"""recman.th.DeepModel - sklearn-style estimator base, mirror of recman/tf/core/DeepModel.py:21-228.

Same public surface (``fit / predict / evaluate / get_batch / fit_on_batch / restore``, abstract
``_out`` / ``_loss``) and the same batch loop, including the reference's quirks that affect results:
``len // batch_size + 1`` batches with a possibly EMPTY last batch (DeepModel.py:49,188 - skipped here
instead of being fed to the kernels), evaluation with ``training=True`` (DeepModel.py:103-111) and a NEW
optimizer for every batch (xDeepFM.py:116-126), which makes every step a stateless first step.

What differs is underneath: batches are packed once into pinned host buffers and copied in one piece per
dtype (``DataInputs.load``), the forward/backward runs on the sm_100a kernels, embedding gradients stay
sparse and the optimizer touches only the rows that received gradient.
"""

from __future__ import annotations

import logging
from abc import ABC, abstractmethod
from time import time
from typing import Dict, Optional

import numpy as np
import torch

from .. import _C, ops
from ..autograd import FmBack, FrontEndFunction, dense_table_grad, pop_sparse_grads
from .input import DataInputs, FeatureDictionary
from .layers import FeatEmbeddingLayer, LinearLayer, PaddedRows

log = logging.getLogger(__name__)

try:
    from sklearn.base import BaseEstimator, TransformerMixin
except Exception:  # pragma: no cover

    class BaseEstimator:  # type: ignore
        pass

    class TransformerMixin:  # type: ignore
        pass


KERAS_EPSILON = 1e-7


def binary_crossentropy(y_true, y_pred):
    """tf.losses.binary_crossentropy on probabilities (tf/core/utils.py:192-194): clip to [eps, 1-eps], log(p+eps)."""
    eps = KERAS_EPSILON
    p = torch.clamp(y_pred, eps, 1 - eps)
    y = y_true.to(p.dtype)
    return (-(y * torch.log(p + eps) + (1 - y) * torch.log(1 - p + eps))).mean(dim=-1)


def create_loss(y_true, y_pred, task):
    if task == "classification":
        return binary_crossentropy(y_true, y_pred)
    if task == "regression":
        return ((y_pred - y_true.to(y_pred.dtype)) ** 2).mean(dim=-1)
    raise ValueError()


def get_linear_features(feat_dict, linear_feats):
    """tf/core/utils.py:26-35."""
    if linear_feats:
        return [feat_dict[name] for name in linear_feats.split(",")]
    return feat_dict.sparse_feats + feat_dict.sparse_val_feats + feat_dict.multi_val_csv_feats + feat_dict.dense_feats


def _shuffle(n, seed):
    return np.random.RandomState(seed).permutation(n)


def _take(X, idx):
    if hasattr(X, "iloc"):
        return X.iloc[idx]
    if isinstance(X, dict):
        return {k: v[idx] for k, v in X.items()}
    return X[idx]


class DeepModel(BaseEstimator, TransformerMixin, ABC):
    def __init__(self, feat_dict: FeatureDictionary, hparams: dict, metrics, epoch, batch_size=64, random_seed=2019,
                 task="classification"):
        assert task in ["classification", "regression"], \
            "target can be either 'classification' for classification task or 'regression' for regression task"
        self.task = task
        self.feat_dict = feat_dict
        self.hparams = hparams
        self.epoch = epoch
        self.batch_size = batch_size
        self.random_seed = random_seed
        self.metrics = metrics
        self.variables: Dict[str, torch.nn.Parameter] = dict()
        self.device = torch.device("cuda")
        self._emb_status = None
        self.samples_seen = 0
        self.shard = None  # th.dist.ShardPlan when the tables are row-sharded over ranks (th.dist.shard_model)

    # ------------------------------------------------------------------ front end shared by the models
    def _embedding_layer(self, use_bias, l2_mode="dense") -> FeatEmbeddingLayer:
        layer = FeatEmbeddingLayer(
            self.variables, self.feat_dict, self.hparams["embedding_size"], self.hparams.get("embedding_l2_reg", 0.0),
            use_bias=use_bias, seed=self.random_seed,
        )
        layer.status = self._status()
        # the layout holds small device tensors built from host lists: build once per model (also keeps the
        # step free of host->device copies, a requirement for CUDA-graph capture)
        key = ("layout", use_bias, id(self.shard))
        cache = self.__dict__.setdefault("_cache", {})
        if key in cache:
            layer._layout = cache[key]
        if self.shard is not None:
            if layer.l2_reg and l2_mode != "touched":
                raise NotImplementedError("row-sharded tables: use embedding_l2_reg=0 or embedding_l2_mode='touched'")
            layer.set_shard(self.shard)
        if l2_mode == "touched":  # scale mode: regularise only rows that were looked up (SURVEY hard part 5)
            layer._upsert_variables()
            self.variables[layer.table_name].rm_l2_touched = float(layer.l2_reg)
            layer.l2_reg = 0.0
        if key not in cache:
            cache[key] = layer.layout()
        return layer

    def _status(self):
        if self._emb_status is None:
            self._emb_status = ops.new_status(self.device)
        return self._emb_status

    def check_ids(self):
        """Raises if any id seen so far was outside its table (host sync; call it outside hot loops)."""
        st = 0 if self._emb_status is None else int(self._emb_status.item())
        if st & 4:
            raise _C.RecmanB200Error("row-sharded tables: a rank owned more ids than the plan capacity "
                                     "(ShardPlan.slack); gradients of the overflowing ids were dropped")
        if st:
            raise _C.RecmanB200Error("an embedding id was outside its table (rows were zero-filled)")

    def _fused_front_end(self, layer: FeatEmbeddingLayer, inputs: DataInputs, linear: Optional[LinearLayer],
                         want_fm: bool):
        """One launch: gather -> [embeds | dense] row buffer (+ FM logit, + first-order logit).

        Eligible when every embedding feature is a one-hot SparseFeat and k % 4 == 0, k <= 128.
        Returns (PaddedRows, fm_logit | None, lin_logit | None) or None when not eligible.
        """
        k = layer.embedding_size
        if not (layer.all_one_hot and k % 4 == 0 and k <= 128 and inputs.sparse_ids is not None):
            if self.shard is not None:
                # the unfused layers know nothing about the exchange: they would index the LOCAL shard with global ids
                raise NotImplementedError(
                    "row-sharded tables need the fused front end: every embedding feature one-hot (no multi-valued "
                    "fields), embedding_size % 4 == 0 and <= 128")
            return None
        if "scal_storage" in self.__dict__.get("_cache", {}):
            raise RuntimeError("this model's k=1 tables are interleaved for the fused tower kernels; the separate kernels "
                               "cannot read that layout (hparams['tower'] must not change after the first forward)")
        layer._upsert_variables()
        table = self.variables[layer.table_name]
        bias_table = self.variables[layer.bias_name] if (layer.use_bias and want_fm) else None
        lin_table = lin_dense = None
        dense = inputs.dense
        n_dense = 0 if dense is None else dense.shape[1]
        if linear is not None:
            # the fused kernel indexes linear_w with the embedding table's row numbers: needs the same layout,
            # i.e. linear features == [all sparse feats in order] + [all dense feats in order]
            feats = linear.linear_feats
            want = self.feat_dict.sparse_feats + self.feat_dict.dense_feats
            if [f.name for f in feats] != [f.name for f in want]:
                if self.shard is not None:
                    raise NotImplementedError("row-sharded tables: linear features must be [sparse..., dense...]")
                return None
            if self.shard is not None:
                linear.total = self.shard.total_local + n_dense  # id rows are sharded like the tables, tail replicated
                linear.alloc = self.shard.alloc
            linear._upsert_variables()
            W = linear.effective_weight()
            if self.shard is not None and W is not self.variables[f"{linear.prefix}linear_w"]:
                raise NotImplementedError("row-sharded tables: inference-time feature weights are not supported")
            flat = W.reshape(-1)
            lin_table = flat[: layer.total_rows]
            lin_dense = flat[layer.total_rows :] if n_dense else None
        lay = layer.layout()
        ids = inputs.sparse_ids
        W_lin = None
        if linear is not None:
            W_lin = self.variables[f"{linear.prefix}linear_w"]
            lin_table = lin_table.detach()
            lin_dense = None if lin_dense is None else lin_dense.detach().contiguous()
        # N1: inside a training step the sparse update is applied by the backward kernel itself (no summed-gradient
        # round trip) when nothing else needs the embedding gradients: no dense L2 on the tables, no touched-rows L2
        fused = getattr(self, "_fused_opt", None)
        if fused is not None and (layer.l2_reg or (linear is not None and linear.l2_reg)
                                  or getattr(table, "rm_l2_touched", 0.0)):
            fused = None
        # DeepFM, opt-in (hparams["fuse_fm_backward"]): let the first MLP layer's input-gradient kernel add the FM backward
        # term and emit the row gradients.  Measured at C5 it is a net loss (the epilogue's x / S reads cost 0.17 ms, the
        # leaner reduce saves 0.10 ms), so the separate kernels stay the default.
        fm_back = FmBack() if (want_fm and torch.is_grad_enabled() and self.hparams.get("fuse_fm_backward", False)) else None
        if self.shard is not None:
            from .dist import P2PFrontEndFunction, ShardedFrontEndFunction

            fn = P2PFrontEndFunction if self.shard.peer is not None else ShardedFrontEndFunction
            if self.shard.peer is None:
                fm_back = None
            x, fm, lin = fn.apply(table, bias_table, W_lin, lin_table, lin_dense, self.shard, self._status(), ids,
                                  dense, fused, fm_back)
        else:
            x, fm, lin = FrontEndFunction.apply(table, bias_table, W_lin, lin_table, lin_dense, lay.runs[0].offsets,
                                                layer.total_rows, self._status(), ids, dense, fused, fm_back)
        d = lay.m * k + n_dense
        if linear is not None:
            lin = lin + self.variables[f"{linear.prefix}linear_w0"]
        rows = PaddedRows(x, d)
        if fm_back is not None and fm.requires_grad:
            rows.fm_back = fm_back  # the model hooks fm_back.on_g_fm onto its final logit (see DeepFM._out)
        return rows, (fm if want_fm else None), (lin if linear is not None else None)

    # ------------------------------------------------------------------ fused DeepFM tower (front end + first DNN layer)
    def _tower_front_end(self, layer: FeatEmbeddingLayer, linear: LinearLayer, dnn, inputs: DataInputs, training: bool,
                         add_w0: bool = True):
        """One kernel for gather + FM + first-order term + the first DNN layer (``rm_tower_fwd``), and - inside a
        training step whose update can be fused - one sorted pass for the whole sparse backward + optimizer
        (``rm_tower_bwd_update``).  Returns (y1 pre-activation, fm_logit, lin_logit) or None when not eligible:
        every embedding feature one-hot, k in {32, 64}, first hidden layer <= 64 wide, linear features ==
        [sparse..., dense...], single GPU.  The two k=1 tables (``feat_bias_table``, ``linear_w``) then share ONE
        interleaved [rows + n_dense, 2] storage, so an id costs one 8-byte lookup instead of two sector fetches."""
        if not self.hparams.get("tower", True) or inputs.sparse_ids is None:
            return None
        sharded = self.shard is not None
        if sharded and self.shard.peer is None:
            return None  # the all-to-all exchange keeps the separate kernels
        k = layer.embedding_size
        if not layer.all_one_hot or not layer.use_bias:
            return None
        lay = layer.layout()
        dense = inputs.dense
        n_dense = 0 if dense is None else dense.shape[1]
        hidden0 = dnn.hidden_units[0]
        if hidden0 is None or not ops.tower_supported(lay.m, k, n_dense, int(hidden0)):
            return None
        if sharded and not (ops.tower_bwd_supported(k, int(hidden0)) and self._fused_opt_config() is not None
                            and not layer.l2_reg and not linear.l2_reg
                            and self.hparams.get("embedding_l2_mode", "dense") != "touched"):
            return None  # sharded training runs only the in-kernel update: everything else keeps the separate kernels
        want = self.feat_dict.sparse_feats + self.feat_dict.dense_feats
        if [f.name for f in linear.linear_feats] != [f.name for f in want]:
            return None
        total = layer.total_rows
        bname, wname = layer.bias_name, f"{linear.prefix}linear_w"
        cache = self.__dict__.setdefault("_cache", {})
        if "scal_storage" not in cache:
            if bname in self.variables or wname in self.variables:
                return None  # the k=1 tables already exist in the separate layout
            st = (self.shard.alloc((total + n_dense, 2)) if sharded
                  else torch.zeros(total + n_dense, 2, dtype=torch.float32, device=self.device))
            self.variables[bname] = torch.nn.Parameter(st[:total, 0], requires_grad=True)
            self.variables[wname] = torch.nn.Parameter(st[:, 1:2], requires_grad=True)
            cache["scal_storage"] = st
        scal = cache["scal_storage"]
        layer._upsert_variables()
        linear._upsert_variables()
        table = self.variables[layer.table_name]
        bias_param, W_lin = self.variables[bname], self.variables[wname]
        if bias_param.data_ptr() != scal.data_ptr() or W_lin.data_ptr() != scal.data_ptr() + 4:
            return None  # a parameter was re-bound: the interleaved storage is no longer what the model trains
        d = lay.m * k + n_dense
        W1, b1 = dnn.first_layer(d)
        scal_fwd = scal
        if not training:  # inference-time per-id weights (layers.py:338-345, 426-437) are added to linear_w
            extra = np.concatenate([np.asarray(f.weights, dtype=np.float32).reshape(-1) for f in linear.linear_feats])
            if np.any(extra != 0):
                if sharded:
                    raise NotImplementedError("row-sharded tables: inference-time feature weights are not supported")
                scal_fwd = scal.clone()
                scal_fwd[:, 1] += torch.from_numpy(extra).to(scal.device)
        fused = getattr(self, "_fused_opt", None)
        if fused is not None and (layer.l2_reg or linear.l2_reg or getattr(table, "rm_l2_touched", 0.0)):
            fused = None
        if fused is not None:  # + which variant(s) of the backward kernel to launch (compile_step sets a hint)
            fused = tuple(fused[:2]) + (int(getattr(self, "_bk_variant", 0)),)
        # side channel head kernel -> tower backward of this step (autograd.TowerSide); the model's loss hands it to
        # HeadFunction when the fused head runs
        from ..autograd import TowerSide

        side = self._tower_side = TowerSide() if torch.is_grad_enabled() else None
        if sharded:
            from .dist import P2PTowerFunction

            if scal_fwd is not scal:
                raise NotImplementedError("row-sharded tables: inference-time feature weights are not supported")
            y1, fm, lin = P2PTowerFunction.apply(table, scal, bias_param, W_lin, W1, b1, self.shard, self._status(),
                                                 inputs.sparse_ids, dense, fused, torch.is_grad_enabled(), side)
            if add_w0:
                lin = lin + self.variables[f"{linear.prefix}linear_w0"]
            return y1, fm, lin
        from ..autograd import TowerFunction

        y1, fm, lin = TowerFunction.apply(table, scal, scal_fwd, bias_param, W_lin, W1, b1, lay.runs[0].offsets, total,
                                          self._status(), inputs.sparse_ids, dense, fused, torch.is_grad_enabled(), side)
        if add_w0:
            lin = lin + self.variables[f"{linear.prefix}linear_w0"]
        return y1, fm, lin

    # ------------------------------------------------------------------ reference API
    def predict(self, X, training=False, batch_number_to_show_progress=50):
        n = len(X) if not isinstance(X, dict) else len(next(iter(X.values())))
        dummy_y = np.ones(n, dtype=np.float32)
        outs = []
        total_batch = n // self.batch_size + 1
        enc = self._encode_once(X, dummy_y) if n > 0 else None  # one-hot + dense: encode once, prefetch pinned slices
        with torch.no_grad():
            if enc is not None:
                from .input import HostPrefetcher

                ids, dense, yy = enc
                bs = self.batch_size

                def source(i):
                    lo, hi = min(i * bs, n), min((i + 1) * bs, n)
                    return ids[lo:hi], (None if dense is None else dense[lo:hi]), yy[lo:hi]

                pf = HostPrefetcher(self.feat_dict, source, self.device)
                for batch_index in range(total_batch):
                    if min(batch_index * bs, n) == min((batch_index + 1) * bs, n):
                        continue
                    outs.append(self._out(pf.get(batch_index), training=training).reshape(-1))
                    if batch_index % batch_number_to_show_progress == 0:
                        log.info(f"Predict: {(batch_index + 1)}/{total_batch} has been completed")
                torch.cuda.synchronize(self.device)
            else:
                for batch_index in range(total_batch):
                    x_batch, y_batch = self.get_batch(X, dummy_y, self.batch_size, batch_index)
                    if len(y_batch) == 0:  # the reference feeds this empty batch to TF; nothing to compute
                        continue
                    inputs = DataInputs(self.device).load(self.feat_dict, x_batch, y_batch)
                    outs.append(self._out(inputs, training=training).reshape(-1))
                    if batch_index % batch_number_to_show_progress == 0:
                        log.info(f"Predict: {(batch_index + 1)}/{total_batch} has been completed")
        log.info(f"Predict: {total_batch}/{total_batch} has been completed")
        if not outs:
            return np.zeros((0,), dtype=np.float32)
        return torch.cat(outs).cpu().numpy()

    def evaluate(self, X, y, training=False, batch_number_to_show_progress=50):
        pred = self.predict(X, training, batch_number_to_show_progress)
        return [metric(y, pred) for metric in self.metrics]

    @staticmethod
    def get_batch(X, y, batch_size, index):
        start = index * batch_size
        end = start + batch_size
        end = end if end < len(y) else len(y)
        if isinstance(X, dict):
            return {k: v[start:end] for k, v in X.items()}, y[start:end]
        return X[start:end], y[start:end]

    # ------------------------------------------------------------------ state
    def _table_layers(self):
        """(embedding layer, linear layer | None) of the last forward, for the reference-named views."""
        return getattr(self, "embeddings", None), getattr(self, "linear", None)

    def state_dict(self, reference_names: bool = True) -> Dict[str, torch.Tensor]:
        """Every variable under the reference's names.  The fused tables are exported as the per-feature variables
        the reference creates (layers.py:95-110): ``{feat}_feat_embed`` [V_f, k] and ``{feat}_feat_bias`` [V_f, 1]
        (row-range views of ``feat_embed_table`` / ``feat_bias_table``); everything else keeps its name.  With
        ``reference_names=False`` the fused tables are exported as they are stored."""
        out = {}
        emb, _ = self._table_layers()
        fused = set()
        if reference_names and emb is not None and self.shard is None and emb.table_name in self.variables:
            out.update(emb.feature_tables())
            fused = {emb.table_name, emb.bias_name}
        for name, p in self.variables.items():
            if name not in fused:
                out[name] = p.data
        return out

    def load_state_dict(self, state: Dict[str, torch.Tensor], strict=True):
        """Accepts both the reference-named per-feature tables and the fused tables.  Returns (missing, unexpected);
        with ``strict`` (default) either raises."""
        emb, _ = self._table_layers()
        views = emb.feature_tables() if (emb is not None and self.shard is None and emb.table_name in self.variables) else {}
        seen, unexpected = set(), []
        for name, t in state.items():
            if name in self.variables:
                self.variables[name].data.copy_(t.to(self.device).reshape(self.variables[name].shape))
                seen.add(name)
            elif name in views:
                views[name].copy_(t.to(self.device).reshape(views[name].shape))
                seen.add(name)
            else:
                unexpected.append(name)
        covered = set(seen)
        if emb is not None and views:
            per_table = [n for n in views if n.endswith("_feat_embed")]
            if per_table and all(n in seen for n in per_table):
                covered.add(emb.table_name)
            per_bias = [n for n in views if n.endswith("_feat_bias")]
            if per_bias and all(n in seen for n in per_bias):
                covered.add(emb.bias_name)
        missing = [n for n in self.variables if n not in covered]
        if strict and (missing or unexpected):
            raise KeyError(f"load_state_dict: missing {missing}, unexpected {unexpected}")
        return missing, unexpected

    def save(self, path="ckpt_model.pt"):
        """torch.save of ``state_dict()``.  With row-sharded tables every rank holds different rows: the file gets a
        ``.rank{r}`` suffix and keeps the fused (local-shard) layout."""
        if self.shard is not None:
            path = f"{path}.rank{self.shard.rank}"
        torch.save({k: v.cpu() for k, v in self.state_dict(reference_names=self.shard is None).items()}, path)

    def restore(self, path="ckpt_model.pt", example=None, strict=True):
        """DeepModel.restore (DeepModel.py:83-86) with torch.save files instead of tf.train.Checkpoint.

        Variables are created lazily by the first forward: on a fresh model pass ``example`` (a DataFrame / dict of
        columns, or DataInputs) so that they exist before the file is loaded; restoring into a model that has no
        variables raises instead of silently loading nothing."""
        if self.shard is not None:
            path = f"{path}.rank{self.shard.rank}"
        if example is not None and not self.variables:
            inputs = example if isinstance(example, DataInputs) else DataInputs(self.device).load(
                self.feat_dict, example, np.zeros(len(next(iter(example.values()))) if isinstance(example, dict)
                                                  else len(example), dtype=np.float32))
            with torch.no_grad():
                self._out(inputs, training=False)
        if not self.variables:
            raise RuntimeError("restore(): the model has no variables yet - run a forward first or pass example=...")
        return self.load_state_dict(torch.load(path), strict=strict)

    # ------------------------------------------------------------------ training
    @abstractmethod
    def _out(self, inputs, training=True):
        pass

    @abstractmethod
    def _loss(self, inputs):
        pass

    def fit_on_batch(self, X, y):
        """One optimisation step (xDeepFM.py:116-126): encode, forward, backward, fresh-optimizer update."""
        inputs = X if isinstance(X, DataInputs) else DataInputs(self.device).load(self.feat_dict, X, y)
        self.samples_seen += inputs.batch_size
        if self.shard is not None and not getattr(self, "_replicas_synced", False):
            self._sync_replicated(inputs)
        if getattr(self, "_graph", None) is not None:
            loss = self._graph_step(inputs)
            if loss is not None:
                return loss
        # eager: with sharded tables every rank's loss is a mean over its local batch; 1/W makes the summed
        # gradients those of the global-batch mean (and counts the replicated L2 terms once)
        return self._eager_step(inputs)

    def _sync_replicated(self, inputs: DataInputs):
        """Row-sharded mode: the dense (replicated) parameters must start equal on every rank - only their gradients
        are all-reduced afterwards.  Creates the variables with one forward and broadcasts rank 0's values."""
        import torch.distributed as dist

        with torch.no_grad():
            self._out(inputs, training=False)
        emb, lin = getattr(self, "embeddings", None), getattr(self, "linear", None)
        sharded = {emb.table_name, emb.bias_name} if emb is not None else set()
        lin_name = f"{lin.prefix}linear_w" if lin is not None else None
        for name, p in self.variables.items():
            if name in sharded:
                continue
            t = p.data
            if name == lin_name:  # id rows are sharded like the tables, the dense tail is replicated
                t = t.reshape(-1)[self.shard.total_local:]
                if t.numel() == 0:
                    continue
            buf = t.contiguous()
            dist.broadcast(buf, src=dist.get_global_rank(self.shard.group, 0) if self.shard.group is not None else 0,
                           group=self.shard.group)
            if buf.data_ptr() != t.data_ptr():
                t.copy_(buf)
        self._replicas_synced = True

    # ------------------------------------------------------------------ CUDA-graphed step (N1)
    def compile_step(self, example: DataInputs, warmup: int = 3):
        """Capture one whole training step (forward, backward, K2, optimizer) into a CUDA graph.

        The step contains no host synchronisation (K2's unique-row count stays on the device) and every buffer it
        touches is allocated from the graph's private pool, so the launch-bound eager sequence (~150 launches for
        DeepFM) becomes one ``cudaGraphLaunch``.  ``fit_on_batch`` then copies the batch into the static input
        buffers and replays.  Not available with row-sharded tables yet (the exchange has a host-side split-size
        sync).  The ``warmup`` eager steps are real optimisation steps on ``example``.
        """
        if self.shard is not None and self.shard.peer is None:
            raise NotImplementedError("compile_step with the all-to-all exchange (split sizes need a host sync); "
                                      "use peer-memory sharding")
        from .. import ops as _ops

        if self.shard is not None and not getattr(self, "_replicas_synced", False):
            self._sync_replicated(example)
        st = DataInputs.from_tensors(
            self.feat_dict, example.sparse_ids.clone(), None if example.dense is None else example.dense.clone(),
            example["y"].clone())
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._eager_step(st)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        # fused tower backward: the eager steps launch both kernel variants and the plan's device flag picks one; the
        # captured step launches only the one the warm-up batch needed (hot rows under skewed ids or not).  Either variant
        # is correct for any batch - a later change of the id distribution costs speed, not results.
        self._bk_variant = 0
        for p_ in self.variables.values():
            flag = getattr(p_, "rm_hot_flag", None)
            if flag is not None:
                self._bk_variant = 2 if int(flag.item()) else 1
                break
        graph = torch.cuda.CUDAGraph()
        before = _ops.launch_count()
        try:
            with torch.cuda.graph(graph):
                loss = self._eager_step(st)
        finally:
            self._bk_variant = 0
        self._graph_launches = _ops.launch_count() - before
        self._graph, self._graph_inputs, self._graph_loss = graph, st, loss
        return self

    def _fused_opt_config(self):
        """(optimizer kind, lr) when the embedding update may be fused into the backward kernel, else None."""
        if not self.hparams.get("fuse_sparse_update", True):
            return None
        opt = self.hparams.get("optimizer", "adam")
        if not isinstance(opt, str) or opt not in _C.OPT_KINDS:
            return None
        return _C.OPT_KINDS[opt], float(self.hparams.get("learning_rate", 0.001))

    def _eager_step(self, inputs: DataInputs):
        self._fused_opt = self._fused_opt_config()
        # one GPU: this method backpropagates d(loss) = 1 itself, so a loss kernel that already holds its gradients
        # (HeadFunction) need not scale them
        self._unit_loss_grad = self.shard is None
        try:
            loss = self._loss(inputs)
            if self.shard is not None:
                (loss / self.shard.world).backward()
            else:
                loss.backward()
        finally:
            self._fused_opt = None
            self._unit_loss_grad = False
        self.optimizer_step()
        return loss.detach()

    def _graph_step(self, inputs: DataInputs):
        st = self._graph_inputs
        if inputs is not st:
            if inputs.sparse_ids.shape != st.sparse_ids.shape:
                return None  # ragged last batch: run it eagerly
            st.sparse_ids.copy_(inputs.sparse_ids, non_blocking=True)
            if st.dense is not None:
                st.dense.copy_(inputs.dense, non_blocking=True)
            st["y"].copy_(inputs["y"], non_blocking=True)
        self._graph.replay()
        return self._graph_loss

    def optimizer_step(self):
        opt = self.hparams.get("optimizer", "adam")
        if not isinstance(opt, str) or opt not in _C.OPT_KINDS:
            raise ValueError(f"optimizer {opt!r}: the kernels implement adam / adagrad / gd / momentum")
        kind = _C.OPT_KINDS[opt]
        lr = float(self.hparams.get("learning_rate", 0.001))
        if self.shard is not None:
            from .dist import allreduce_dense

            dense = []
            for p in self.variables.values():
                if not getattr(p, "rm_sparse_grads", None) and p.grad is not None:
                    dense.append(p.grad)
                tail = getattr(p, "rm_dense_tail", None)
                if tail is not None:
                    dense.append(tail[1])
            allreduce_dense(dense, self.shard.group)
        dense_pairs = []  # (parameter, gradient) of every dense update: one multi-tensor launch at the end

        def add_dense(pv, g):
            if pv.is_contiguous():
                dense_pairs.append((pv, g.contiguous()))
            else:  # a strided view of the interleaved k=1 storage (tower layout): packed copy, update, write back
                ops.dense_opt_step(pv, g.reshape(pv.shape), kind, lr, 0.0)

        for name, p in self.variables.items():
            sparse = pop_sparse_grads(p)
            tail = getattr(p, "rm_dense_tail", None)
            p.rm_dense_tail = None
            if sparse and p.grad is None:
                l2 = float(getattr(p, "rm_l2_touched", 0.0))
                for sg in sparse:
                    ops.sparse_opt_step(p.data, sg, kind, lr, l2)
                if tail is not None:
                    first, g = tail
                    add_dense(p.data.reshape(-1)[first:], g)
            elif sparse:
                # a dense part exists (the reference's whole-table L2, layers.py:188-193): densify and update all rows
                p.rm_sparse_grads = sparse
                g = dense_table_grad(p, consume=True)
                if tail is not None:
                    first, gt = tail
                    g.reshape(-1)[first:] += gt
                ops.dense_opt_step(p.data, g.contiguous(), kind, lr, 0.0)
            elif p.grad is not None:
                add_dense(p.data, p.grad)
            elif tail is not None:  # id rows already updated by the fused backward kernel: only the dense tail is left
                first, g = tail
                add_dense(p.data.reshape(-1)[first:], g)
            p.grad = None
        ops.dense_opt_step_multi(dense_pairs, kind, lr, 0.0)

    # ------------------------------------------------------------------ N2: encode once, prefetch batches
    def _encode_once(self, X, y):
        """Whole training frame -> pinned host arrays (ids int64 [n, m], dense float32 [n, nd] | None, y float32 [n]).

        The reference re-encodes every mini-batch with pandas / LabelEncoder (inputs.py:128-139, DeepModel.py:188); the
        encoders are row-wise, so encoding the frame once and slicing gives the same batches.  Only for models whose
        features are one-hot sparse + dense (multi-valued CSV fields keep the per-batch path); returns None otherwise."""
        fd = self.feat_dict
        if fd.multi_val_csv_feats or not fd.sparse_feats or self.device.type != "cuda":
            return None
        from .input import _column

        ids = np.concatenate([f(_column(X, f.name)) for f in fd.sparse_feats], axis=1).astype(np.int64)
        dense = None
        if fd.dense_feats:
            dense = np.concatenate([f(_column(X, f.name)) for f in fd.dense_feats], axis=1).astype(np.float32)
        yv = np.asarray(y, dtype=np.float32).reshape(-1)
        pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return pin(ids), pin(dense), pin(yv)

    @staticmethod
    def _permute_encoded(enc, perm):
        idx = torch.from_numpy(np.asarray(perm, dtype=np.int64))
        out = []
        for t in enc:
            if t is None:
                out.append(None)
            else:
                buf = torch.empty_like(t).pin_memory()
                torch.index_select(t, 0, idx, out=buf)
                out.append(buf)
        return tuple(out)

    def _fit_epoch_prefetched(self, enc, total_batch, show_every):
        from .input import HostPrefetcher

        ids, dense, y = enc
        n, bs = y.shape[0], self.batch_size

        def source(i):
            lo, hi = min(i * bs, n), min((i + 1) * bs, n)
            return ids[lo:hi], (None if dense is None else dense[lo:hi]), y[lo:hi]

        pf = HostPrefetcher(self.feat_dict, source, self.device)
        for i in range(total_batch):
            if min(i * bs, n) == min((i + 1) * bs, n):
                continue  # the reference's empty trailing batch
            self.fit_on_batch(pf.get(i), None)
            if i % show_every == 0:
                log.info(f"Fit: {(i + 1)}/{total_batch} has been completed")
        torch.cuda.synchronize(self.device)  # the pinned epoch buffers are about to be replaced

    def _eval_at_epoch(self, X_train, y_train, X_valid=None, y_valid=None, start_time=None, epoch=0,
                       batch_number_to_show_progress=50):
        start_time = time() if start_time is None else start_time
        has_valid = X_valid is not None and y_valid is not None
        train_res = self.evaluate(X_train, y_train, training=True,
                                  batch_number_to_show_progress=batch_number_to_show_progress)
        valid_res = self.evaluate(X_valid, y_valid, training=True) if has_valid else None
        log.info("[%d] train-result=%s%s [%.1f s]" % (
            epoch, str([(str(f), round(float(r), 4)) for f, r in zip(self.metrics, train_res)]),
            (", valid-result=%s" % str([(str(f), round(float(r), 4)) for f, r in zip(self.metrics, valid_res)]))
            if has_valid else "", time() - start_time))
        return train_res, valid_res

    def fit(self, X_train, y_train, X_valid=None, y_valid=None, random_seed_for_mini_batch=True, tb_logger=None,
            epoch_callback=None, show_progress=False, batch_number_to_show_progress=50):
        assert X_train is not None or y_train is not None
        y_train = np.asarray(y_train)
        eval_results = self._eval_at_epoch(X_train, y_train, X_valid, y_valid, time())
        self.history = [eval_results]
        enc = self._encode_once(X_train, y_train)  # one-hot + dense models: encode once, then pinned slices + prefetch
        for epoch in range(1, self.epoch + 1):
            t0 = time()
            seed = np.random.randint(1, 2019) if random_seed_for_mini_batch else self.random_seed
            perm = _shuffle(len(y_train), seed)
            X_train, y_train = _take(X_train, perm), y_train[perm]
            total_batch = len(y_train) // self.batch_size + 1
            if enc is not None:
                enc = self._permute_encoded(enc, perm)
                self._fit_epoch_prefetched(enc, total_batch, batch_number_to_show_progress)
            else:
                for i in range(total_batch):
                    Xi_batch, y_batch = self.get_batch(X_train, y_train, self.batch_size, i)
                    if len(y_batch) == 0:
                        continue
                    self.fit_on_batch(Xi_batch, y_batch)
                    if i % batch_number_to_show_progress == 0:
                        log.info(f"Fit: {(i + 1)}/{total_batch} has been completed")
            log.info(f"Fit: {total_batch}/{total_batch} has been completed [{time() - t0:.1f} s]")
            eval_results = self._eval_at_epoch(X_train, y_train, X_valid, y_valid, time(), epoch,
                                               batch_number_to_show_progress)
            self.history.append(eval_results)
            if epoch_callback:
                epoch_callback(model=self, eval_results=eval_results, df_all=X_train[:1])
        self.check_ids()
        return self
