"""recman.th - torch-side drop-in namespace (the reference reserves it with an empty stub, recman/th/)."""
from . import hparams, input, layers, metric  # noqa: F401
from .BestModelFinder import BestModelFinder  # noqa: F401
from .DCN import DCN  # noqa: F401
from .DeepFM import DeepFM  # noqa: F401
from .DeepModel import DeepModel  # noqa: F401
from .input import (  # noqa: F401
    DataInputs,
    DenseFeat,
    FeatureDictionary,
    MultiValCsvFeat,
    MultiValFeat,
    SparseFeat,
)
from .xDeepFM import xDeepFM  # noqa: F401
