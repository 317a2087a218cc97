"""The C-ABI library loads and exports every symbol include/recman_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "recman_b200.h")
LIB = os.path.join(ROOT, "recman_b200", "librecman_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(LIB):
        import __graft_entry__

        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ["rm_gather_fwd", "rm_segment_plan", "rm_segment_reduce", "rm_fm_fwd", "rm_fm_bwd", "rm_cross_fwd",
                 "rm_cross_bwd", "rm_cin_layer_fwd", "rm_cin_layer_bwd", "rm_gather_fm_fwd", "rm_emb_fm_bwd"]:
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    for name in declared_symbols():
        assert hasattr(built_lib, name), f"{name} declared in the header but not exported"


def test_ctypes_binding_covers_the_header(built_lib):
    from recman_b200 import _C

    assert sorted(_C.EXPORTS) == declared_symbols()
    assert _C.lib.rm_version() == 4


def test_argument_validation_without_a_gpu(built_lib):
    """Bad arguments are rejected before any CUDA call, with a message."""
    from recman_b200 import _C

    rc = _C.lib.rm_gather_fwd(None, None, None, 4, 2, 8, None, 16, None, None)
    assert rc == -1 and "null pointer" in _C.last_error()
    assert _C.lib.rm_gather_fwd(None, None, None, 0, 2, 8, None, 16, None, None) == 0  # empty batch is a no-op
    rc = _C.lib.rm_cross_fwd(1, 4096, 1, 1, 1, None, 4, 4096, 2, 1, None, None)
    assert rc == -2 and "2048" in _C.last_error()
    with pytest.raises(_C.RecmanB200Error):
        _C.call("rm_sparse_opt_step", 1, 4, 1, 1, 1, 10, 99, 0.1, 0.0, None)


def test_no_cpu_fallback_in_product_code():
    """The product package never imports the oracle."""
    pkg = os.path.join(ROOT, "recman_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
