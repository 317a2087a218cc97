"""recman_b200: B200-native (sm_100a) kernels + host layers for recman's CTR hot path.

Importing the package loads ``librecman_b200.so`` through ctypes and fails loudly if it has
not been built (``make -C recman_b200/csrc``); there is no CPU or eager-PyTorch fallback.
"""
from . import _C  # noqa: F401  (raises ImportError when the shared object is missing)

__version__ = "0.1.0"
