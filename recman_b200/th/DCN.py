"""recman.th.DCN - Deep & Cross network to the composition of the stale shell recman/tf/core/DCN.py:28-169:

    dnn_input = flatten(embeds) ++ dense ; dnn_logit = DNN(dnn_input) ; cn_logit = CrossNet(L, l2)(dnn_input)
    final = add_n([dnn_logit, cn_logit, dnn_logit]) (+ linear_logit)      <- dnn_logit is summed TWICE (DCN.py:140-142)

``CrossNet`` itself is absent from the reference (DCN.py:7,135); its arithmetic is arXiv 1708.05123 eq. 3.
"""

from __future__ import annotations

from .DeepModel import DeepModel, create_loss
from .input import DataInputs, FeatureDictionary
from .layers import DNN, CrossNet, DNNCombiner, LinearCombiner, LinearLayer, PaddedRows, PredictionLayer, relu


class DCN(DeepModel):
    def __init__(
        self,
        feat_dict: FeatureDictionary,
        embedding_size=8,
        embedding_l2_reg=0.00001,
        linear_l2_reg=0.00001,
        deep_hidden_units=(32, 32),
        deep_dropout=(0.6, 0.6, 0.6),
        deep_activation=relu,
        deep_l2_reg=0.0,
        cross_layer_num=3,
        cross_layer_l2_reg=0.0,
        epoch=10,
        batch_size=64,
        learning_rate=0.001,
        optimizer="adam",
        random_seed=2019,
        use_linear=True,
        loss_type="logloss",
        eval_metric=(),
        what_means_greater=None,
        use_interactive_session=False,
        log_dir="./logs",
        embedding_l2_mode="dense",
    ):
        hparams = dict(
            embedding_size=embedding_size, embedding_l2_reg=embedding_l2_reg, linear_l2_reg=linear_l2_reg,
            deep_hidden_units=tuple(deep_hidden_units), deep_dropout=tuple(deep_dropout),
            deep_activation=deep_activation, deep_l2_reg=deep_l2_reg, cross_layer_num=cross_layer_num,
            cross_layer_l2_reg=cross_layer_l2_reg, learning_rate=learning_rate, optimizer=optimizer,
            embedding_l2_mode=embedding_l2_mode,
        )
        DeepModel.__init__(self, feat_dict, hparams, eval_metric, epoch, batch_size, random_seed,
                           task="classification" if loss_type == "logloss" else "regression")
        self.use_linear = use_linear
        self.loss_type = loss_type

    def _out(self, inputs: DataInputs, training=True):
        hp = self.hparams
        # the shell builds FeatEmbeddingLayer with its default use_bias=True but never uses the bias (DCN.py:107-117)
        self.embeddings = self._embedding_layer(use_bias=False, l2_mode=hp["embedding_l2_mode"])
        self.linear = None
        if self.use_linear:
            all_feats = list(self.feat_dict.values())
            fused_feats = self.feat_dict.sparse_feats + self.feat_dict.dense_feats
            self.linear = LinearLayer(self.variables, fused_feats if len(fused_feats) == len(all_feats) else all_feats,
                                      hp["linear_l2_reg"], training=training)
        fused = self._fused_front_end(self.embeddings, inputs, self.linear, want_fm=False)
        if fused is not None:
            dnn_input, _, linear_logit = fused
        else:
            feat_embeds, _ = self.embeddings(inputs)
            linear_logit = self.linear(LinearCombiner(self.linear.linear_feats)(inputs)) if self.use_linear else None
            dnn_input = DNNCombiner()([feat_embeds] + inputs.dense_inputs(self.feat_dict))

        self.dnn = DNN(self.variables, hp["deep_hidden_units"],
                       hp["deep_dropout"] if training else (1.0,) * len(hp["deep_dropout"]),
                       hp["deep_activation"], hp["deep_l2_reg"])
        self.dnn.training = training
        dnn_logit = self.dnn(dnn_input)
        self.cross_net = CrossNet(hp["cross_layer_num"], hp["cross_layer_l2_reg"], variables=self.variables,
                                  seed=self.random_seed)
        cn_logit = self.cross_net(dnn_input)
        final_logit = dnn_logit + cn_logit + dnn_logit
        if self.use_linear:
            final_logit = final_logit + linear_logit
        self.final_logit = final_logit.detach()  # detached: keeping the graph alive would pin its grad accumulators
        return PredictionLayer(self.variables, self.task, use_bias=False)(final_logit)

    def _loss(self, inputs):
        loss = create_loss(inputs.y, self._out(inputs), task=self.task)
        loss = loss + self.embeddings.l2()
        if self.use_linear:
            loss = loss + self.linear.l2()
        return loss + self.dnn.l2() + self.cross_net.l2()
