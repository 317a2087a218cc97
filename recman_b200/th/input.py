"""Feature schema and batch encoding - the ``recman.th.input`` module the reference stub imports
(recman/th/DeepFM.py:9) and the torch-side mirror of recman/tf/inputs.py.

Same vocabulary and semantics as the reference (``SparseFeat`` tables have ``feat_size + 1`` rows with
row 0 = unknown ``"-----"``, inputs.py:116-126,166; ``MultiValCsvFeat`` tag ids start at 1 with 0 = unknown,
inputs.py:380-396; ``DenseFeat`` values go through an sklearn scaler, inputs.py:308-316), but the per-batch
encoding is vectorised (``np.searchsorted`` instead of a pandas ``isin`` + ``LabelEncoder.transform`` per
feature per batch, inputs.py:128-139) and ``DataInputs.load`` packs every batch into three contiguous pinned
host buffers - int64 ids ``[B, m_sparse]``, float32 dense ``[B, n_dense]``, CSR for multi-valued fields - that
travel to the GPU in one copy each.
"""

from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

try:  # pandas is only needed for DataFrame inputs
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None

__all__ = [
    "FeatureDictionary",
    "DataInputs",
    "ResilientLabelEncoder",
    "SparseFeat",
    "DenseFeat",
    "MultiValCsvFeat",
    "MultiValFeat",
    "MultiValSparseFeat",
    "SparseValueFeat",
    "SequenceFeat",
]


def _column(X, name):
    col = X[name]
    if pd is not None and isinstance(col, pd.Series):
        return col.to_numpy()
    if isinstance(col, torch.Tensor):
        return col.cpu().numpy()
    return np.asarray(col)


class ResilientLabelEncoder:
    """Label encoder with a reserved class 0 for unseen values (inputs.py:116-146)."""

    def __init__(self, null_val="-----"):
        self.null_val = null_val
        self.classes_ = None
        self._sorted = None

    def fit(self, X, y=None):
        vals = np.unique(np.asarray(X if not (pd is not None and isinstance(X, pd.Series)) else X.to_numpy()))
        self._sorted = vals
        # class list as the reference exposes it: null first, then the sorted classes
        self.classes_ = np.concatenate((np.array([self.null_val], dtype=object), vals.astype(object)))
        return self

    def transform(self, X):
        x = np.asarray(X if not (pd is not None and isinstance(X, pd.Series)) else X.to_numpy())
        if self._sorted is None:
            raise RuntimeError("encoder is not fitted")
        try:
            x = x.astype(self._sorted.dtype, copy=False)
        except (TypeError, ValueError):
            x = x.astype(object)
        pos = np.searchsorted(self._sorted, x)
        pos_c = np.minimum(pos, len(self._sorted) - 1)
        known = self._sorted[pos_c] == x
        return np.where(known, pos_c + 1, 0).astype(np.int64).reshape(-1, 1)

    def fit_transform(self, X, y=None):
        return self.fit(X, y).transform(X)

    def inverse_transform(self, y):
        y = np.asarray(y).reshape(-1)
        return self.classes_[y]


class _FeatBase:
    kind = "base"

    def __str__(self):
        return f"{type(self).__name__}({self.name}, {self.feat_size}, {self.dtype})"

    __repr__ = __str__


class SparseFeat(_FeatBase):
    """Single categorical field -> one id per sample (inputs.py:148-211)."""

    kind = "sparse"

    def __init__(self, name, feat_size, weights=None, dtype=torch.int64, encoder=None, description=None):
        self.name = name
        self.dtype = dtype
        self.description = description
        # encoder=False disables encoding (inputs are already ids in [0, feat_size])
        self.encoder = ResilientLabelEncoder() if encoder is None else encoder
        self.feat_size = feat_size + 1  # +1: row 0 is the unknown class
        self._weights = weights
        self._weights_cache = None

    @property
    def weights(self):
        """Inference-time additive first-order weights per id (layers.py:426-437)."""
        if self._weights:
            if self._weights_cache is None:
                keys = list(self._weights.keys())
                ids = self.encoder.transform(keys).reshape(-1) if self.encoder else np.asarray(keys)
                w = np.zeros((self.feat_size,))
                for idx, val in zip(ids, self._weights.values()):
                    w[int(idx)] = val
                self._weights_cache = w
            return self._weights_cache
        return np.zeros((self.feat_size,))

    def set_weights(self, val):
        self._weights = val
        self._weights_cache = None

    def get_shape(self, for_tf=True):
        return (None if for_tf else -1), 1

    def initialize(self, X):
        if self.encoder:
            self.encoder.fit(X)

    def __call__(self, x):
        if self.encoder:
            return self.encoder.transform(x)
        return np.asarray(x).astype(np.int64).reshape(-1, 1)

    def decode(self, x):
        return self.encoder.inverse_transform(x) if self.encoder else x


class SparseValueFeat(SparseFeat):
    """(id, value) field.  The reference's lookup for it is shape-invalid (layers.py:142); not on the hot path."""

    kind = "sparse_value"

    def __call__(self, x):
        raise NotImplementedError("SparseValueFeat is broken in the reference (layers.py:142) and out of scope")


class DenseFeat(_FeatBase):
    """Numeric field, scaled (inputs.py:281-322)."""

    kind = "dense"

    def __init__(self, name, weights=None, dtype=torch.float32, scaler=None, description=None):
        from sklearn.preprocessing import StandardScaler

        self.name = name
        self.dtype = dtype
        self.description = description
        self.scaler = StandardScaler() if scaler is None else scaler  # scaler=False: raw values
        self.feat_size = 1
        self._weights = weights

    @property
    def weights(self):
        return [self._weights if self._weights is not None else 0]

    def get_shape(self, for_tf=True):
        return (None if for_tf else -1), 1

    def initialize(self, X):
        if self.scaler:
            self.scaler.fit(np.asarray(X, dtype=np.float64).reshape(-1, 1))

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float32)
        if self.scaler:
            x = self.scaler.transform(x.reshape(-1, 1))
        return np.asarray(x, dtype=np.float32).reshape(-1, 1)


class MultiValCsvFeat(_FeatBase):
    """Multi-valued field given as ``"a|b|d"`` strings (inputs.py:380-425).

    The reference parses the strings on device with tf.io.decode_csv + a StaticHashTable
    (tf/core/utils.py:70-83); here they are parsed on the host into CSR ``(values, offsets)``.
    Tag ids are 1-based, unknown tags map to 0, an empty string yields an empty row.
    """

    kind = "multi"

    def __init__(self, name, tags=(), weights=None, dtype=str, description=None, sep="|"):
        self.name = name
        self.dtype = dtype
        self.description = description
        self.tags = tags
        self.sep = sep
        self.tag_hash_table = dict((tag, idx + 1) for idx, tag in enumerate(self.tags))
        self.feat_size = len(self.tags) + 1
        self._weights = weights
        self._weights_cache = None

    def get_shape(self, for_tf=True):
        return (None if for_tf else -1), 1

    def initialize(self, X):
        pass

    def set_weights(self, val):
        self._weights = val
        self._weights_cache = None

    @property
    def weights(self):
        if self._weights:
            if self._weights_cache is None:
                self._weights_cache = np.zeros((self.feat_size,))
                for tag, weight in self._weights.items():
                    if tag in self.tag_hash_table:
                        self._weights_cache[self.tag_hash_table[tag]] = weight
            return self._weights_cache
        return np.zeros((self.feat_size,))

    def __call__(self, x):
        """-> (values int64 [nnz], offsets int64 [B+1])"""
        table = self.tag_hash_table
        values: List[int] = []
        offsets = [0]
        for s in np.asarray(x).reshape(-1):
            if isinstance(s, str) and s != "":
                for tok in s.split(self.sep):
                    values.append(table.get(tok, 0))
            offsets.append(len(values))
        return np.asarray(values, dtype=np.int64), np.asarray(offsets, dtype=np.int64)


MultiValFeat = MultiValCsvFeat  # the name recman/th/DeepFM.py:9 imports


class MultiValSparseFeat(_FeatBase):
    """Present in the reference vocabulary but ``to_sparse_tensor`` raises for it (tf/core/utils.py:113-117)."""

    kind = "multi_sparse"

    def __init__(self, *a, **k):
        raise NotImplementedError("MultiValSparseFeat has no working lookup in the reference (utils.py:113-117)")


class SequenceFeat(_FeatBase):
    def __init__(self, *a, **k):
        raise NotImplementedError("Sequence feature is not yet completed")  # same as inputs.py:443


class FeatureDictionary(OrderedDict):
    """Ordered name -> feature map (inputs.py:8-43)."""

    @property
    def embedding_feats(self):
        return [f for f in self.values() if not isinstance(f, DenseFeat)]

    @property
    def sparse_feats(self):
        return [f for f in self.values() if type(f) is SparseFeat]

    @property
    def sparse_val_feats(self):
        return [f for f in self.values() if isinstance(f, SparseValueFeat)]

    @property
    def dense_feats(self):
        return [f for f in self.values() if isinstance(f, DenseFeat)]

    @property
    def multi_val_csv_feats(self):
        return [f for f in self.values() if isinstance(f, MultiValCsvFeat)]

    @property
    def multi_val_sparse_feats(self):
        return []

    @property
    def sequence_feats(self):
        return []

    def initialize(self, X):
        for feat in self.values():
            feat.initialize(_column(X, feat.name))


class DataInputs(dict):
    """One encoded mini-batch on the device (mirror of inputs.py:46-93).

    Keys are feature names (plus ``"y"``) like the reference; additionally the packed forms the kernels
    consume are kept: ``sparse_ids`` int64 [B, m_sparse] (feature-dictionary order of the SparseFeats),
    ``dense`` float32 [B, n_dense], and per multi-valued feature a CSR ``(values, offsets)`` pair.
    ``h2d_bytes`` counts what crossed PCIe.
    """

    def __init__(self, device: Optional[str] = None):
        super().__init__()
        self.device = torch.device(device if device is not None else "cuda")
        self.sparse_ids: Optional[torch.Tensor] = None
        self.dense: Optional[torch.Tensor] = None
        self.csr: Dict[str, tuple] = {}
        self.h2d_bytes = 0
        self.batch_size = 0

    def _to_device(self, arr: np.ndarray) -> torch.Tensor:
        arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:  # pandas may hand out read-only views
            arr = arr.copy()
        t = torch.from_numpy(arr)
        if self.device.type == "cuda":
            t = t.pin_memory().to(self.device, non_blocking=True)
        self.h2d_bytes += t.numel() * t.element_size()
        return t

    def load(self, feat_dict: FeatureDictionary, X, y=None):
        sparse = feat_dict.sparse_feats
        dense = feat_dict.dense_feats
        B = None
        if sparse:
            ids = np.concatenate([f(_column(X, f.name)) for f in sparse], axis=1).astype(np.int64)
            B = ids.shape[0]
            self.sparse_ids = self._to_device(ids)
            for j, f in enumerate(sparse):
                self[f.name] = self.sparse_ids[:, j : j + 1]
        if dense:
            dv = np.concatenate([f(_column(X, f.name)) for f in dense], axis=1).astype(np.float32)
            B = dv.shape[0]
            self.dense = self._to_device(dv)
            for j, f in enumerate(dense):
                self[f.name] = self.dense[:, j : j + 1]
        for f in feat_dict.multi_val_csv_feats:
            values, offsets = f(_column(X, f.name))
            B = offsets.shape[0] - 1
            pair = (self._to_device(values), self._to_device(offsets))
            self.csr[f.name] = pair
            self[f.name] = pair
        self.batch_size = 0 if B is None else B
        if y is not None:
            self["y"] = self._to_device(np.asarray(y, dtype=np.float32).reshape(-1))
        return self

    @classmethod
    def from_tensors(cls, feat_dict, sparse_ids, dense=None, y=None, csr=None):
        """Wrap tensors that are already on the device (synthetic benchmarks)."""
        self = cls(device=str(sparse_ids.device))
        self.sparse_ids = sparse_ids
        self.dense = dense
        self.batch_size = sparse_ids.shape[0]
        for j, f in enumerate(feat_dict.sparse_feats):
            self[f.name] = sparse_ids[:, j : j + 1]
        if dense is not None:
            for j, f in enumerate(feat_dict.dense_feats):
                self[f.name] = dense[:, j : j + 1]
        for name, pair in (csr or {}).items():
            self.csr[name] = pair
            self[name] = pair
        if y is not None:
            self["y"] = y
        return self

    @property
    def y(self):
        return self["y"]

    def dense_inputs(self, feat_dict):
        return [self[f.name] for f in feat_dict.dense_feats]

    def sparse_inputs(self, feat_dict):
        return [self[f.name] for f in feat_dict.sparse_feats]

    def embedding_inputs(self, feat_dict):
        return [self[f.name] for f in feat_dict.embedding_feats]

    def multi_val_csv_inputs(self, feat_dict):
        return [self[f.name] for f in feat_dict.multi_val_csv_feats]


class HostPrefetcher:
    """Double-buffered host->device input pipeline: while step i runs, the pinned host buffers of step
    i+1 are already on their way over PCIe on a copy stream, so the GPU is not input-starved.

    ``source(i)`` returns the pinned host tensors ``(sparse_ids int64 [B,m], dense float32 [B,n] | None, y float32 [B])``
    of step i.  ``get(i)`` returns the device ``DataInputs`` of step i (copied on the copy stream, the current stream
    waits for the copy) and starts the copy of step i+1.  Every step's copy happens, one step early.
    """

    def __init__(self, feat_dict: "FeatureDictionary", source, device=None):
        self.feat_dict = feat_dict
        self.source = source
        self.device = torch.device(device if device is not None else "cuda")
        self.stream = torch.cuda.Stream(device=self.device)
        self._pending = None  # (step index, device tensors, copy-done event)
        self.h2d_bytes = 0

    def _stage(self, i):
        host = self.source(i)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            dev = tuple(None if t is None else t.to(self.device, non_blocking=True) for t in host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        for t in dev:
            if t is not None:
                t.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream
                self.h2d_bytes += t.numel() * t.element_size()
        return i, dev, ev

    def get(self, i) -> "DataInputs":
        if self._pending is None or self._pending[0] != i:
            self._pending = self._stage(i)
        _, (ids, dense, y), ev = self._pending
        self._pending = self._stage(i + 1)
        torch.cuda.current_stream(self.device).wait_event(ev)
        return DataInputs.from_tensors(self.feat_dict, ids, dense, y)

