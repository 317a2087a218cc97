"""recman.th.DeepFM - the class the reference stub reserves (recman/th/DeepFM.py:12-13), built to the
composition and keyword vocabulary of the stale TF1 shell recman/tf/core/DeepFM.py:30-183:

    embeddings + bias (use_bias=True) -> LinearLayer ; FMLayer(embeds, bias) ; DNN(flatten(embeds) ++ dense)
    -> add_n([linear, fm, dnn]) -> sigmoid ;  loss = logloss + L2(embeds, linear, dnn)
"""

from __future__ import annotations

import torch

from ..autograd import GradTap
from .DeepModel import DeepModel, create_loss
from .input import DataInputs, FeatureDictionary
from .layers import DNN, DNNCombiner, FMLayer, LinearCombiner, LinearLayer, PredictionLayer, relu


class DeepFM(DeepModel):
    def __init__(
        self,
        feat_dict: FeatureDictionary,
        embedding_size=8,
        embedding_l2_reg=0.00001,
        linear_l2_reg=0.00001,
        fm_dropout=(1.0, 1.0),
        deep_hidden_units=(32, 32),
        deep_dropout=(0.8, 0.8, 0.8),
        deep_l2_reg=0.00001,
        deep_activation=relu,
        epoch=10,
        batch_size=64,
        learning_rate=0.001,
        optimizer="adam",
        random_seed=2019,
        use_fm=True,
        use_deep=True,
        loss_type="logloss",
        eval_metric=(),
        what_means_greater=None,
        use_interactive_session=False,
        log_dir="./logs",
        embedding_l2_mode="dense",
    ):
        assert use_fm or use_deep
        assert loss_type in ["logloss", "mse"], \
            "loss_type can be either 'logloss' for classification task or 'mse' for regression task"
        hparams = dict(
            embedding_size=embedding_size, embedding_l2_reg=embedding_l2_reg, linear_l2_reg=linear_l2_reg,
            fm_dropout=tuple(fm_dropout), deep_hidden_units=tuple(deep_hidden_units), deep_dropout=tuple(deep_dropout),
            deep_l2_reg=deep_l2_reg, deep_activation=deep_activation, learning_rate=learning_rate, optimizer=optimizer,
            embedding_l2_mode=embedding_l2_mode,
        )
        DeepModel.__init__(self, feat_dict, hparams, eval_metric, epoch, batch_size, random_seed,
                           task="classification" if loss_type == "logloss" else "regression")
        self.use_fm = use_fm
        self.use_deep = use_deep
        self.loss_type = loss_type

    def _tower_layers(self, inputs, training):
        """The layers of the fused tower path and its raw outputs (y1, fm, lin without w0), or None when the model /
        hyper-parameters are outside what the fused kernels cover."""
        hp = self.hparams
        # a per-model decision, NOT per call: the tower keeps the two k=1 tables interleaved in one array, a layout the
        # separate kernels do not read - a model whose training step cannot use the tower (FM dropout, dropout on the DNN
        # input) must not use it for inference either
        if not (self.use_fm and self.use_deep and all(p >= 1 for p in hp["fm_dropout"]) and hp["deep_dropout"][0] >= 1):
            return None
        self.embeddings = self._embedding_layer(use_bias=True, l2_mode=hp["embedding_l2_mode"])
        linear_feats = list(self.feat_dict.values())
        fused_feats = self.feat_dict.sparse_feats + self.feat_dict.dense_feats
        self.linear = LinearLayer(self.variables, fused_feats if len(fused_feats) == len(linear_feats) else linear_feats,
                                  hp["linear_l2_reg"], training=training)
        self.dnn = DNN(self.variables, hp["deep_hidden_units"],
                       hp["deep_dropout"] if training else (1.0,) * len(hp["deep_dropout"]),
                       hp["deep_activation"], hp["deep_l2_reg"])
        self.dnn.training = training
        return self._tower_front_end(self.embeddings, self.linear, self.dnn, inputs, training, add_w0=False)

    def _head_args(self, training):
        """(w0, W2, b2, w3, b3, act kind) when everything after the first DNN layer fits the fused head kernel:
        two hidden layers of 32, identity / relu / leaky_relu, no dropout."""
        from .. import _C, ops
        from .layers import activation_kind

        if not self.hparams.get("fused_head", True):
            return None
        hu = self.dnn.hidden_units
        if len(hu) != 2 or not ops.deepfm_head_supported(int(hu[0]), int(hu[1])):
            return None
        if training and any(p < 1 for p in self.hparams["deep_dropout"][1:]):
            return None
        try:
            act = activation_kind(self.hparams["deep_activation"])
        except ValueError:
            return None
        v, p = self.variables, self.dnn.prefix
        return (v[f"{self.linear.prefix}linear_w0"], v[f"{p}dnn_layer_1_weights"], v[f"{p}dnn_layer_1_bias"],
                v[f"{p}dnn_w"], v[f"{p}dnn_w0"], act)

    def _out(self, inputs: DataInputs, training=True, _tower=None):
        hp = self.hparams
        tower = _tower if _tower is not None else self._tower_layers(inputs, training)
        if tower is not None:
            # one kernel: gather + FM + first-order + first DNN layer (tcgen05); the row buffer is never written
            y1, fm_logit, lin_raw = tower
            head = self._head_args(training)
            if head is not None and not torch.is_grad_enabled():
                from .. import ops

                w0, W2, b2, w3, b3, act = head
                logit, pred = ops.deepfm_head(y1, fm_logit.reshape(-1), lin_raw.reshape(-1), w0.data, W2.data, b2.data,
                                              w3.data.reshape(-1), b3.data, None, act,
                                              0 if self.task == "classification" else 1)
                self.final_logit = logit.reshape(-1, 1)
                return pred
            linear_logit = lin_raw + self.variables[f"{self.linear.prefix}linear_w0"]
            final_logit = linear_logit + fm_logit + self.dnn.from_first_layer(y1)
            self.final_logit = final_logit.detach()
            return PredictionLayer(self.variables, self.task, use_bias=False)(final_logit)

        self.embeddings = self._embedding_layer(use_bias=True, l2_mode=hp["embedding_l2_mode"])
        linear_feats = list(self.feat_dict.values())  # LinearCombiner(self.feat_dict), DeepFM.py:122
        fused_feats = self.feat_dict.sparse_feats + self.feat_dict.dense_feats
        self.linear = LinearLayer(self.variables, fused_feats if len(fused_feats) == len(linear_feats) else linear_feats,
                                  hp["linear_l2_reg"], training=training)
        fm_dropout = hp["fm_dropout"] if training else (1.0,) * len(hp["fm_dropout"])
        fm_identity = all(p >= 1 for p in fm_dropout)

        # FM dropout (a legal reference hyper-parameter, DeepFM.py:36) needs the [B,m,k] block and the bias block
        # before the reduction: the unfused layers handle it
        fused = None
        if fm_identity or not self.use_fm or self.shard is not None:
            fused = self._fused_front_end(self.embeddings, inputs, self.linear, want_fm=self.use_fm and fm_identity)
        if fused is not None:
            rows, fm_logit, linear_logit = fused
            m, k = len(self.embeddings.feats), hp["embedding_size"]
            if self.use_fm and fm_logit is None:
                raise NotImplementedError("row-sharded tables: fm_dropout < 1 is not supported")
            dnn_input = rows
        else:
            feat_embeds, feat_bias = self.embeddings(inputs)
            linear_logit = self.linear(LinearCombiner(self.linear.linear_feats)(inputs))
            fm_logit = None
            if self.use_fm:
                fm = FMLayer(fm_dropout)
                fm.training = training
                fm_logit = fm(feat_embeds, feat_bias)
            dnn_input = DNNCombiner()([feat_embeds] + inputs.dense_inputs(self.feat_dict)) if self.use_deep else None

        final_logit = linear_logit
        if self.use_fm:
            final_logit = final_logit + fm_logit
        self.dnn = None
        if self.use_deep:
            self.dnn = DNN(self.variables, hp["deep_hidden_units"],
                           hp["deep_dropout"] if training else (1.0,) * len(hp["deep_dropout"]),
                           hp["deep_activation"], hp["deep_l2_reg"])
            self.dnn.training = training
            final_logit = final_logit + self.dnn(dnn_input)
        fm_back = getattr(dnn_input, "fm_back", None) if self.use_deep else None
        if fm_back is not None and final_logit.requires_grad:
            # d(loss)/d(fm_logit) == d(loss)/d(final_logit) (plain sum); a tap on the final logit runs at the very start of
            # the backward pass, i.e. before autograd walks into the MLP, whose first layer then fuses the FM backward
            final_logit = GradTap.apply(final_logit, fm_back)
        self.final_logit = final_logit.detach()  # detached: keeping the graph alive would pin its grad accumulators
        return PredictionLayer(self.variables, self.task, use_bias=False)(final_logit)

    def _l2_terms(self):
        hp = self.hparams
        terms = []
        if self.embeddings.l2_reg:
            terms.append(self.embeddings.l2())
        if self.linear.l2_reg:
            terms.append(self.linear.l2())
        if self.use_deep and self.dnn is not None and self.dnn.l2_reg:
            terms.append(self.dnn.l2())
        return terms

    def _loss(self, inputs):
        tower = self._tower_layers(inputs, True) if torch.is_grad_enabled() else None
        if tower is not None:
            head = self._head_args(True)
            if head is not None:
                # second DNN layer .. loss, forward and backward, in one kernel
                from ..autograd import HeadFunction

                y1, fm_logit, lin_raw = tower
                w0, W2, b2, w3, b3, act = head
                loss, logit, _ = HeadFunction.apply(y1, fm_logit, lin_raw, w0, W2, b2, w3, b3, inputs.y, act,
                                                    0 if self.task == "classification" else 1, inputs.dense,
                                                    getattr(self, "_tower_side", None),
                                                    getattr(self, "_unit_loss_grad", False))
                self.final_logit = logit.reshape(-1, 1)
                for t in self._l2_terms():
                    loss = loss + t
                return loss
        loss = create_loss(inputs.y, self._out(inputs, _tower=tower), task=self.task)
        for t in self._l2_terms():
            loss = loss + t
        return loss
