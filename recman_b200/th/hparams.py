"""Hyper-parameter vocabulary: key names and defaults of recman/tf/hparams/xDeepFM.py:7-34 and
BaseHyperParameters.py:67-100, without the TensorBoard plugin.  Models receive a plain dict."""

import itertools

from .layers import leaky_relu


class HParam:
    def __init__(self, name, default_value):
        assert name
        self._name = name
        self._default_value = default_value
        self._domain = [default_value]

    def __call__(self, domain=None):
        if domain is None:
            domain = [self._default_value]
        self._domain = list(getattr(domain, "values", domain))
        return self

    @property
    def name(self):
        return self._name

    @property
    def hp_domain(self):
        return self._domain

    @property
    def default_value(self):
        return self._default_value


class Discrete:
    """Stand-in for tensorboard.plugins.hparams.api.Discrete."""

    def __init__(self, values):
        self.values = list(values)


class BaseHyperParameters(dict):
    LearningRate = "learning_rate"
    Optimizer = "optimizer"

    def __init__(self):
        dict.__init__(self)
        self.add_param(self.LearningRate, 0.001)
        self.add_param(self.Optimizer, "adam")

    def add_param(self, name, default_val):
        self[name] = HParam(name, default_val)()

    def grid_search(self, print_hp=False):
        axes = [[(p.name, v) for v in p.hp_domain] for p in self.values()]
        for bags in itertools.product(*axes):
            d = dict(bags)
            if print_hp:
                print(d)
            yield d

    def defaults(self):
        return {p.name: p.default_value for p in self.values()}


class xDeepFM(BaseHyperParameters):
    EmbeddingSize = "embedding_size"
    EmbeddingL2Reg = "embedding_l2_reg"
    LinearL2Reg = "linear_l2_reg"
    LinearFeatures = "linear_features"
    DeepHiddenUnits = "deep_hidden_units"
    DeepDropOut = "deep_dropout"
    DeepActivation = "deep_activation"
    DeepL2Reg = "deep_l2_reg"
    CinCrossLayerUnits = "cin_cross_layer_units"
    CinDropOut = "cin_dropout"
    CinActivation = "cin_activation"
    CinL2Reg = "cin_l2_reg"

    def __init__(self):
        BaseHyperParameters.__init__(self)
        self.add_param(self.EmbeddingSize, 8)
        self.add_param(self.EmbeddingL2Reg, 0.00001)
        self.add_param(self.LinearL2Reg, 0.00001)
        self.add_param(self.LinearFeatures, [])
        self.add_param(self.DeepHiddenUnits, (32, 32))
        self.add_param(self.DeepDropOut, (0.8, 0.8, 0.8))
        self.add_param(self.DeepActivation, leaky_relu)
        self.add_param(self.DeepL2Reg, 0.00001)
        self.add_param(self.CinCrossLayerUnits, [100, 100, 100])
        self.add_param(self.CinDropOut, [1, 1, 1, 1])
        self.add_param(self.CinActivation, leaky_relu)
        self.add_param(self.CinL2Reg, 0.00001)
