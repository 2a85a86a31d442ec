"""auto_differentiation_b200 — B200-native (sm_100a) implementation of the autodiff hot path of
Jianqoq/Auto-Differentiation behind the reference's own Python surface.

    from auto_differentiation_b200.auto_grad import tgrad as ad      # Matrix / Function / ops / gradient()
    from auto_differentiation_b200.layers import layers              # TimeAffine, LSTM, MultiHeadAttention, ...
    from auto_differentiation_b200.utils.optimizer import Adam
    from auto_differentiation_b200.utils.Train import Trainer

Compute goes through the C ABI of libb200ad.so (include/b200ad.h); there is no CPU fallback.
"""
from . import _lib
from .device import (DeviceArray, asarray, as_index, empty, zeros, full, set_gemm_precision,  # noqa: F401
                     get_gemm_precision)

__all__ = ["DeviceArray", "asarray", "as_index", "empty", "zeros", "full", "set_gemm_precision",
           "get_gemm_precision", "install_reference_aliases"]


def install_reference_aliases():
    """Register the tape under the module names the reference's scripts import (`ad`, `tgrad`), so that
    e.g. Test/MultiHeadAttentionTest.py's `import ad` resolves to the B200 path."""
    import sys
    from .auto_grad import tgrad
    sys.modules.setdefault("ad", tgrad)
    sys.modules.setdefault("tgrad", tgrad)
    return tgrad
