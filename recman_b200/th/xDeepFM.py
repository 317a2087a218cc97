"""recman.th.xDeepFM - mirror of recman/tf/core/xDeepFM.py:20-126 (the reference's only live model)."""

from __future__ import annotations

import torch

from .DeepModel import DeepModel, create_loss, get_linear_features
from .hparams import xDeepFM as HyperParams
from .input import DataInputs, FeatureDictionary
from .layers import (
    CIN,
    DNN,
    DNNCombiner,
    PredictionLayer,
    SparseLinearCombiner,
    SparseLinearLayer,
)


class xDeepFM(DeepModel):
    """xDeepFM (arXiv 1803.05170): sigmoid(linear + CIN + DNN).

    ``hparams`` is the reference's plain dict (hparams/xDeepFM.py:7-18 + learning_rate, optimizer).  Extra,
    optional keys: ``cin_precision`` ("3xtf32" parity default | "tf32" | "fp32"), ``embedding_l2_mode``
    ("dense" = reference semantics | "touched").
    """

    def __init__(self, feat_dict: FeatureDictionary, hparams: dict, task="classification", metrics=(), epoch=10,
                 batch_size=64, random_seed=2019):
        full = HyperParams().defaults()
        full.update(hparams or {})
        DeepModel.__init__(self, feat_dict=feat_dict, hparams=full, epoch=epoch, batch_size=batch_size,
                           random_seed=random_seed, metrics=metrics, task=task)

    def _out(self, inputs: DataInputs, training=True):
        hp = self.hparams
        self.embeddings = self._embedding_layer(use_bias=False, l2_mode=hp.get("embedding_l2_mode", "dense"))
        linear_feats = get_linear_features(self.feat_dict, hp[HyperParams.LinearFeatures])
        self.linear = SparseLinearLayer(self.variables, linear_feats, hp[HyperParams.LinearL2Reg], training=training)
        k = hp[HyperParams.EmbeddingSize]
        m = len(self.embeddings.feats)

        fused = self._fused_front_end(self.embeddings, inputs, self.linear, want_fm=False)
        if fused is not None:
            rows, _, linear_logit = fused
            feat_embeds = rows.buf[:, : m * k].unflatten(1, (m, k))  # view into the row buffer
            dnn_input = rows
        else:
            feat_embeds, _ = self.embeddings(inputs)
            linear_logit = self.linear(SparseLinearCombiner(linear_feats)(inputs))
            dnn_input = DNNCombiner()([feat_embeds] + inputs.dense_inputs(self.feat_dict))

        self.cin = CIN(
            self.variables, hp[HyperParams.CinCrossLayerUnits], hp[HyperParams.CinActivation],
            hp[HyperParams.CinDropOut] if training else [1] * len(hp[HyperParams.CinDropOut]),
            hp[HyperParams.CinL2Reg], seed=self.random_seed, precision=hp.get("cin_precision", "3xtf32"),
        )
        cin_logit = self.cin(feat_embeds)

        self.dnn = DNN(
            self.variables, hp[HyperParams.DeepHiddenUnits],
            hp[HyperParams.DeepDropOut] if training else [1] * len(hp[HyperParams.DeepDropOut]),
            hp[HyperParams.DeepActivation], hp[HyperParams.DeepL2Reg],
        )
        dnn_logit = self.dnn(dnn_input)
        final_logit = linear_logit + cin_logit + dnn_logit
        self.final_logit = final_logit.detach()  # detached: keeping the graph alive would pin its grad accumulators
        return PredictionLayer(self.variables, self.task)(final_logit)

    def _loss(self, inputs):
        loss = create_loss(inputs.y, self._out(inputs), task=self.task)
        return loss + sum(layer.l2() for layer in [self.embeddings, self.linear, self.dnn, self.cin])
