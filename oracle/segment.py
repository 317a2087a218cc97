"""numpy oracle for A4: the sparse embedding-gradient scatter-add (test infrastructure).

The reference has no line for this - it is TensorFlow's autodiff of
``tf.nn.embedding_lookup`` (layers.py:118-128): an ``IndexedSlices`` gradient
whose duplicate rows are summed.  Restated here as: stable-sort the
``(table, row)`` keys, then sum every run of equal keys IN ASCENDING POSITION
ORDER with plain fp32 adds, which is also the summation order the CUDA kernel
commits to (that makes the result deterministic and, for the unfused kernel,
bit-reproducible on the CPU).

PARITY UNPINNED - see ``oracle/__init__.py``.
"""

from __future__ import annotations

import numpy as np

__all__ = ["global_rows", "segment_sum_sorted", "dense_table_grad", "shard_route", "csr_expand"]


def global_rows(ids: np.ndarray, table_offsets: np.ndarray) -> np.ndarray:
    """ids [B, m] (row within table f) -> global row ``table_offsets[f] + ids[b, f]``, int64 [B, m]."""
    ids = np.asarray(ids, dtype=np.int64)
    return ids + np.asarray(table_offsets, dtype=np.int64)[None, : ids.shape[1]]


SEG_LONG = 16   # segments longer than this are summed chunk-wise (recman_b200/csrc/scatter.cu: SEG_LONG)
SEG_CHUNK = 32  # positions per chunk (SEG_CHUNK)


def segment_sum_sorted(keys: np.ndarray, grads: np.ndarray, scale: np.ndarray | None = None, chunked: bool = True):
    """Deterministic scatter-add.

    keys  [N] int64 global rows (position p = b*m + f for one-hot fields)
    grads [N, k] float32 - gradient row of position p
    scale [N] optional per-position multiplier (1/sqrt(n) of sqrtn pooling)

    Returns ``(unique_rows [U] int64 ascending, sums [U, k] float32,
    order [N] int64, seg_start [U+1] int64)`` where ``order`` is the stable
    sort permutation and ``sums[u] = sum over order[seg_start[u]:seg_start[u+1]]``
    accumulated sequentially in that order in float32.  With ``chunked`` (the order the CUDA kernels commit to), a
    segment longer than ``SEG_LONG`` positions - a hot row under skewed ids - is summed as: every ``SEG_CHUNK``
    consecutive positions sequentially into a partial, then the partials sequentially in chunk order.
    """
    keys = np.asarray(keys, dtype=np.int64).reshape(-1)
    grads = np.asarray(grads)
    N = keys.shape[0]
    k = grads.shape[1] if grads.ndim == 2 else 1
    g = grads.reshape(N, k)
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    if N == 0:
        return (np.zeros(0, np.int64), np.zeros((0, k), g.dtype), order, np.zeros(1, np.int64))
    boundary = np.concatenate(([True], sk[1:] != sk[:-1]))
    seg_start = np.flatnonzero(boundary)
    uniq = sk[seg_start]
    seg_start = np.concatenate((seg_start, [N])).astype(np.int64)
    sums = np.zeros((uniq.shape[0], k), dtype=g.dtype)
    gs = g[order]
    if scale is not None:
        gs = (gs * np.asarray(scale, dtype=g.dtype)[order][:, None]).astype(g.dtype)
    # sequential accumulation in sorted order: vectorised over segments, stepping
    # through the j-th element of every segment at once (keeps fp32 add order).
    lens = np.diff(seg_start)
    short = lens <= SEG_LONG if chunked else np.ones_like(lens, dtype=bool)
    max_short = int(lens[short].max()) if short.any() else 0
    for j in range(max_short):
        live = np.flatnonzero(short & (lens > j))
        sums[live] = (sums[live] + gs[seg_start[live] + j]).astype(g.dtype)
    for u in np.flatnonzero(~short):  # long segments: chunk partials, then the partials in chunk order
        total = np.zeros(k, dtype=g.dtype)
        for c0 in range(int(seg_start[u]), int(seg_start[u + 1]), SEG_CHUNK):
            part = np.zeros(k, dtype=g.dtype)
            for j in range(c0, min(c0 + SEG_CHUNK, int(seg_start[u + 1]))):
                part = (part + gs[j]).astype(g.dtype)
            total = (total + part).astype(g.dtype)
        sums[u] = total
    return uniq, sums, order, seg_start


def dense_table_grad(keys: np.ndarray, grads: np.ndarray, total_rows: int) -> np.ndarray:
    """The same gradient as a dense ``[total_rows, k]`` array (float64 accumulate) for cross-checks."""
    g = np.asarray(grads, dtype=np.float64)
    out = np.zeros((total_rows, g.shape[1]), dtype=np.float64)
    np.add.at(out, np.asarray(keys, dtype=np.int64).reshape(-1), g)
    return out


def csr_expand(offsets: np.ndarray) -> np.ndarray:
    """CSR offsets [B+1] -> sample index of every value (``np.repeat``)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    return np.repeat(np.arange(offsets.shape[0] - 1, dtype=np.int64), np.diff(offsets))


def shard_route(rows: np.ndarray, world: int):
    """Cyclic row sharding of SURVEY section 8(e): row r of a table lives on rank
    ``r mod W`` at local row ``r div W``.  Returns ``(owner, local_row)``."""
    rows = np.asarray(rows, dtype=np.int64)
    return rows % world, rows // world
