"""ctypes binding of librecman_b200.so (``include/recman_b200.h``).

No pybind / ATen dispatch and NO fallback: if the shared object is missing the
import of this module raises, and every call checks the C return code and
raises ``RecmanB200Error`` with the library's ``rm_last_error()`` message.
The library is built in-tree by ``__graft_entry__.build()`` /
``make -C recman_b200/csrc``.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librecman_b200.so")


class RecmanB200Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C recman_b200/csrc` (there is no CPU or PyTorch fallback)"
    )

lib = ctypes.CDLL(LIB_PATH)

P = c_void_p  # every device pointer travels as an integer address

_SIGS = {
    "rm_version": (ctypes.c_int, []),
    "rm_last_error": (c_char_p, []),
    "rm_device_check": (ctypes.c_int, [ctypes.c_int]),
    "rm_launch_count": (c_int64, []),
    "rm_gather_fwd": (ctypes.c_int, [P, P, P, c_int64, c_int32, c_int32, P, c_int64, P, P]),
    "rm_gather_pooled_fwd": (ctypes.c_int, [P, c_int64, c_int64, P, P, c_int64, c_int32, P, c_int64, P, P]),
    "rm_fm_fwd": (ctypes.c_int, [P, c_int64, P, c_int64, c_int32, c_int32, P, P, P]),
    "rm_fm_bwd": (ctypes.c_int, [P, c_int64, P, P, c_int64, c_int32, c_int32, P, c_int64, P, c_int32, P]),
    "rm_gather_fm_fwd": (
        ctypes.c_int,
        [P, P, P, P, P, P, P, c_int32, c_int64, c_int32, c_int32, P, c_int64, P, P, P, P, P],
    ),
    "rm_segment_reduce_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "rm_segment_plan_workspace_bytes": (c_size_t, [c_int64]),
    "rm_segment_plan": (ctypes.c_int, [P, P, c_int64, c_int32, c_int64, P, c_size_t, P, P, P, P, P, P]),
    "rm_segment_reduce": (ctypes.c_int, [P, c_int64, c_int32, c_int32, c_int64, P, P, P, P, P, c_size_t, P]),
    "rm_emb_fm_bwd": (
        ctypes.c_int,
        [P, P, c_int64, P, P, P, c_int32, c_int32, c_int64, P, P, P, P, P, P, P, c_size_t, P]),
    "rm_emb_fm_bwd_update": (
        ctypes.c_int,
        [P, P, c_int64, P, P, P, c_int32, c_int32, c_int64, P, P, P, P, P, P, P, c_int32, c_float, c_float, P, c_size_t, P]),
    "rm_cross_fwd": (ctypes.c_int, [P, c_int64, P, P, P, P, c_int64, c_int32, c_int32, P, P, P]),
    "rm_cross_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "rm_cross_bwd": (
        ctypes.c_int,
        [P, c_int64, P, P, P, P, P, c_int64, c_int32, c_int32, P, c_int64, c_int32, P, P, P, P, P, c_size_t, P],
    ),
    "rm_cin_layer_fwd": (
        ctypes.c_int,
        [P, c_int64, P, c_int64, P, P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, P, P, P, c_size_t,
         P],
    ),
    "rm_cin_layer_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "rm_cin_layer_bwd": (
        ctypes.c_int,
        [P, c_int64, P, c_int64, P, P, P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, P, P, P, P,
         c_int64, P, c_size_t, P],
    ),
    "rm_cin_layer_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "rm_cin_pool_fwd": (ctypes.c_int, [P, c_int64, c_int32, c_int32, c_int32, P, P]),
    "rm_cin_pool_bwd": (ctypes.c_int, [P, c_int64, P, c_int64, c_int32, c_int32, c_int32, P, P]),
    "rm_unpack_rows": (ctypes.c_int, [P, c_int64, c_int32, P, c_int32, c_int32, P, c_int64, P, P, P]),
    "rm_pack_grad_rows": (ctypes.c_int, [P, P, c_int64, P, P, P, c_int64, c_int32, P, c_int32, c_int32, P, P]),
    "rm_p2p_alloc": (ctypes.c_int, [c_size_t, P, P]),
    "rm_p2p_open": (ctypes.c_int, [P, P]),
    "rm_p2p_close": (ctypes.c_int, [P]),
    "rm_p2p_free": (ctypes.c_int, [P]),
    "rm_gather_fm_fwd_p2p": (
        ctypes.c_int,
        [P, P, P, c_int32, P, P, P, P, P, c_int32, c_int64, c_int32, c_int32, P, c_int64, P, P, P, P, P],
    ),
    "rm_shard_plan_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "rm_shard_plan": (
        ctypes.c_int,
        [P, c_int64, c_int32, c_int32, c_int32, P, P, c_int64, c_int64, P, c_size_t, P, P, P, P, P, P, P],
    ),
    "rm_segment_reduce_p2p": (ctypes.c_int, [P, P, c_int32, c_int32, c_int64, c_int32, c_int32, c_int64, P, P, P, P, P, P, P, c_size_t, P]),
    "rm_segment_reduce_p2p_update": (
        ctypes.c_int,
        [P, P, c_int32, c_int32, c_int64, c_int32, c_int32, c_int64, P, P, P, P, P, P, P, c_int32, c_float, c_float, P, c_size_t, P]),
    "rm_sparse_opt_step": (ctypes.c_int, [P, c_int32, P, P, P, c_int64, c_int32, c_float, c_float, P]),
    "rm_sparse_opt_step_strided": (ctypes.c_int, [P, c_int32, c_int64, P, P, P, c_int64, c_int32, c_float, c_float, P]),
    "rm_dense_opt_step": (ctypes.c_int, [P, P, c_int64, c_int32, c_float, c_float, P]),
    "rm_linear_bwd_input": (ctypes.c_int, [P, c_int64, c_int32, P, c_int32, P, c_int64, P]),
    "rm_linear_bwd_input_fm": (ctypes.c_int, [P, c_int64, c_int32, P, c_int32, c_int32, P, c_int64, P, P, P, P]),
    "rm_linear_bwd_weight_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "rm_linear_bwd_weight": (ctypes.c_int, [P, c_int64, P, c_int64, c_int32, c_int32, P, P, c_size_t, P]),
    "rm_dense_opt_step_multi": (ctypes.c_int, [P, P, P, c_int32, c_int32, c_float, c_float, P]),
    "rm_tower_supported": (ctypes.c_int, [c_int32, c_int32, c_int32, c_int32]),
    "rm_tower_fwd_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "rm_tower_fwd": (
        ctypes.c_int,
        [P, P, P, P, P, P, c_int32, c_int32, P, P, c_int32, c_int64, c_int32, c_int32, P, c_int64, P, P, P, P, P, P,
         c_size_t, P]),
    "rm_tower_units_per_field": (c_int32, [c_int64, c_int32]),
    "rm_tower_plan_workspace_bytes": (c_size_t, [c_int64]),
    "rm_tower_plan": (ctypes.c_int, [P, P, c_int64, c_int32, c_int64, c_int32, P, c_size_t, P, P, P, P, P, P]),
    "rm_tower_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "rm_tower_bwd_update": (
        ctypes.c_int,
        [P, P, P, P, P, P, P, P, P, P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_float, c_int32,
         P, P, P, P, P, c_size_t, P]),
    "rm_umma_probe": (ctypes.c_int, [P, P, c_int32, c_int32, P, P, P]),
    "rm_tower_fwd_p2p": (
        ctypes.c_int,
        [P, P, c_int32, P, P, P, P, P, c_int32, c_int32, P, P, c_int32, c_int64, c_int32, c_int32, P, P, P, P, P, P,
         c_size_t, P]),
    "rm_tower_shard_plan_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "rm_tower_shard_plan": (
        ctypes.c_int,
        [P, c_int64, c_int32, c_int32, c_int32, P, P, c_int64, c_int64, c_int64, c_int32, P, c_size_t, P, P, P, P, P, P,
         P]),
    "rm_deepfm_head_supported": (ctypes.c_int, [c_int32, c_int32]),
    "rm_deepfm_head_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "rm_deepfm_head": (
        ctypes.c_int,
        [P, P, P, P, P, P, P, P, P, P, c_int32, c_int64, c_int32, c_int32, c_int32, c_int32, c_float, P, P, P, P, P, P, P,
         P, P, P, P, P, P, c_size_t, P]),
}

EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    msg = lib.rm_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RecmanB200Error(f"{what} failed (rc={rc}): {last_error()}")


_profile = None  # name -> list of (start_event, end_event) while bench.py profiles a timed region


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero code."""
    if _profile is not None:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        _profile.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    check(rc, name)


def enable_profile():
    """Bracket every C-ABI call with CUDA events on the current stream (no synchronisation)."""
    global _profile
    _profile = {}


def disable_profile():
    """-> {entry point: (total ms, calls)}; synchronises once."""
    global _profile
    import torch

    torch.cuda.synchronize()
    out = {}
    for name, evs in (_profile or {}).items():
        out[name] = (sum(a.elapsed_time(b) for a, b in evs), len(evs))
    _profile = None
    return out


# enum mirrors
OPT_ADAM, OPT_ADAGRAD, OPT_GD = 0, 1, 2
ACT_IDENTITY, ACT_RELU, ACT_LEAKY_RELU = 0, 1, 2
CIN_FP32_SIMT, CIN_3XTF32, CIN_TF32 = 0, 1, 2

OPT_KINDS = {"adam": OPT_ADAM, "adagrad": OPT_ADAGRAD, "gd": OPT_GD, "momentum": OPT_GD, "sgd": OPT_GD}
ACT_KINDS = {"identity": ACT_IDENTITY, "linear": ACT_IDENTITY, "relu": ACT_RELU, "leaky_relu": ACT_LEAKY_RELU}
