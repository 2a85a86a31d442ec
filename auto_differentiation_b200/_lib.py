"""ctypes binding of libb200ad.so (the C ABI declared in include/b200ad.h).

The product path has NO fallback: if the shared library is missing, or a compute entry is called
without a CUDA device, the call raises.  Building is explicit (``python -m
auto_differentiation_b200.build`` or ``__graft_entry__.build()``).
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200ad.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3

# enum b200ad_unary_op
U_EXP, U_TANH, U_SIGMOID, U_RELU, U_LN, U_LOGB, U_SQRT, U_SQUARE, U_POW, U_ABS, U_NEG = range(11)
U_SIN, U_COS, U_TAN, U_SEC, U_SINH, U_COSH, U_ARCSIN, U_ARCTAN, U_ARCOT, U_ARCSEC = range(11, 21)
U_SCALE, U_ADDC, U_DIVC, U_RSUBC, U_RDIVC = range(21, 26)
U_COT, U_CSC, U_ARCCOS = range(26, 29)
# enum b200ad_binary_op / b200ad_bcast_mode
B_ADD, B_SUB, B_MUL, B_DIV = range(4)
FULL, ROW, COL, SCALAR = range(4)
# enum b200ad_gemm_precision
GEMM_FP32, GEMM_TF32X3, GEMM_BF16, GEMM_AUTO, GEMM_TF32, GEMM_F16X3 = range(6)
ACT_NONE, ACT_TANH, ACT_SIGMOID, ACT_RELU = range(4)
OPT_CHUNK = 4096
ARG_WS_BYTES = 1184 * 8 + 1184 * 4


class GemmDesc(Structure):
    _fields_ = [("precision", c_int32), ("trans_a", c_int32), ("trans_b", c_int32), ("act", c_int32),
                ("m", c_int64), ("n", c_int64), ("k", c_int64),
                ("lda", c_int64), ("ldb", c_int64), ("ldc", c_int64),
                ("batch0", c_int64), ("batch1", c_int64),
                ("a_stride0", c_int64), ("a_stride1", c_int64), ("b_stride0", c_int64),
                ("b_stride1", c_int64), ("c_stride0", c_int64), ("c_stride1", c_int64),
                ("alpha", c_float), ("beta", c_float)]


class AttnStrides(Structure):
    _fields_ = [("batch", c_int64), ("head", c_int64), ("row", c_int64)]


class AttnDesc(Structure):
    _fields_ = [("batch", c_int64), ("heads", c_int64), ("seq_q", c_int64), ("seq_kv", c_int64),
                ("head_dim", c_int64), ("scale", c_float), ("reserved", c_int32),
                ("q", AttnStrides), ("k", AttnStrides), ("v", AttnStrides), ("o", AttnStrides)]


class TensorEntry(Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("m", c_void_p), ("v", c_void_p),
                ("numel", c_int64)]


P = c_void_p  # device pointers travel as integers
I64 = c_int64

# name -> (restype, argtypes); every symbol include/b200ad.h declares.
SIGNATURES = {
    "b200ad_abi_version": (c_int, []),
    "b200ad_last_error": (c_char_p, []),
    "b200ad_launch_count": (c_int64, []),
    "b200ad_device_info": (c_int, [c_char_p, c_int, POINTER(c_int), POINTER(c_int)]),
    "b200ad_unary_fwd": (c_int, [c_int, P, P, I64, c_float, P]),
    "b200ad_unary_bwd": (c_int, [c_int, P, P, P, P, I64, c_float, P]),
    "b200ad_unary_bwd_uses": (c_int, [c_int]),
    "b200ad_binary_fwd": (c_int, [c_int, P, P, P, I64, c_int, c_int, I64, P]),
    "b200ad_binary_bwd": (c_int, [c_int, P, P, P, P, P, I64, c_int, c_int, I64, c_int, P]),
    "b200ad_binary_bwd_rowfused": (c_int, [c_int, P, P, P, P, P, I64, I64, c_int, P, P]),
    "b200ad_reduce_ws_floats": (c_int64, [I64, I64, I64]),
    "b200ad_colsum_ws_floats": (c_int64, [I64, I64]),
    "b200ad_reduce": (c_int, [P, P, I64, I64, I64, c_float, P, P]),
    "b200ad_repeat_axis": (c_int, [P, P, I64, I64, I64, c_float, P]),
    "b200ad_softmax_fwd": (c_int, [P, P, I64, I64, c_float, P]),
    "b200ad_softmax_bwd": (c_int, [P, P, P, I64, I64, c_float, c_int, P]),
    "b200ad_xent_fwd": (c_int, [P, P, P, P, P, I64, I64, I64, c_float, c_int, c_int, P]),
    "b200ad_xent_bwd": (c_int, [P, P, P, P, I64, I64, I64, c_float, c_int, P]),
    "b200ad_argext_flat": (c_int, [P, I64, c_int, P, P, P]),
    "b200ad_argmax_rows": (c_int, [P, I64, I64, P, P]),
    "b200ad_embedding_gather": (c_int, [P, P, P, I64, I64, I64, P]),
    "b200ad_embedding_scatter_add": (c_int, [P, P, P, I64, I64, I64, P]),
    "b200ad_strided_copy": (c_int, [P, P, c_int, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), P]),
    "b200ad_fill": (c_int, [P, I64, c_float, P]),
    "b200ad_axpby": (c_int, [P, P, I64, c_float, c_float, P]),
    "b200ad_dropout_fwd": (c_int, [P, P, P, P, I64, c_float, c_float, P]),
    "b200ad_uniform": (c_int, [P, I64, ctypes.c_uint64, ctypes.c_uint64, P]),
    "b200ad_gemm": (c_int, [POINTER(GemmDesc), P, P, P, P, P, I64, P]),
    "b200ad_gemm_supported": (c_int, [POINTER(GemmDesc), P, P, P]),
    "b200ad_gemm_workspace_bytes": (c_int64, [POINTER(GemmDesc)]),
    "b200ad_gemm_set_recurrent_dispatch": (c_int, [c_int]),
    "b200ad_gemm_reload_env": (c_int, []),
    "b200ad_gemm_set_f16x3_auto": (c_int, [c_int]),
    "b200ad_gemm_shared_a_workspace_bytes": (c_int64, [POINTER(GemmDesc), c_int32]),
    "b200ad_gemm_shared_a": (c_int, [POINTER(GemmDesc), P, c_int32, POINTER(c_void_p), POINTER(c_void_p),
                                     POINTER(c_void_p), P, I64, P]),
    "b200ad_attn_supported": (c_int, [POINTER(AttnDesc)]),
    "b200ad_attn_set_variant": (c_int, [c_int]),
    "b200ad_attn_fwd_ws_floats": (c_int64, [POINTER(AttnDesc)]),
    "b200ad_attn_fwd": (c_int, [POINTER(AttnDesc), P, P, P, P, P, P, P]),
    "b200ad_attn_bwd_ws_floats": (c_int64, [POINTER(AttnDesc)]),
    "b200ad_attn_bwd": (c_int, [POINTER(AttnDesc), P, P, P, P, P, P, P, P, P, P, P]),
    "b200ad_attn_bwd_maxima": (c_int, [POINTER(AttnDesc), P, P, P, P, P, P, P, P, P, P, P, P]),
    "b200ad_lstm_cell_fwd": (c_int, [P, I64, P, I64, P, I64, P, I64, P, I64, I64, I64, P]),
    "b200ad_lstm_cell_bwd": (c_int, [P, I64, P, I64, P, P, I64, P, I64, P, I64, P, I64, P, I64, P,
                                     I64, I64, P]),
    "b200ad_rnn_cell_bwd": (c_int, [P, I64, P, I64, P, I64, P, I64, I64, I64, P]),
    "b200ad_rnn_seq_supported": (c_int, [I64, I64]),
    "b200ad_rnn_tanh_seq_fwd": (c_int, [P, P, P, I64, I64, I64, P]),
    "b200ad_rnn_tanh_seq_bwd": (c_int, [P, I64, I64, P, P, P, P, I64, I64, I64, P]),
    "b200ad_lstm_seq_supported": (c_int, [I64, I64]),
    "b200ad_lstm_seq_workspace_bytes": (I64, [I64, I64]),
    "b200ad_lstm_seq_fwd": (c_int, [P, P, P, P, P, I64, I64, I64, P, I64, P]),
    "b200ad_lstm_seq_bwd": (c_int, [P, I64, I64, P, P, P, P, P, P, P, I64, I64, I64, P, I64, P]),
    "b200ad_bias_act_fwd": (c_int, [P, I64, P, P, I64, I64, I64, c_int, P]),
    "b200ad_act_bwd": (c_int, [P, I64, P, I64, P, I64, I64, I64, c_int, P]),
    "b200ad_adam_multi": (c_int, [P, c_int, I64, P, P, c_float, c_float, c_float, c_float, c_float, P, P]),
    "b200ad_adam_step_scalar": (c_int, [P, ctypes.c_double, ctypes.c_double, ctypes.c_double, P, P]),
    "b200ad_sgd_multi": (c_int, [P, c_int, I64, P, P, c_float, c_float, P]),
    "b200ad_sumsq_multi": (c_int, [P, c_int, I64, P, P, P, P, P]),
    "b200ad_clip_scale_multi": (c_int, [P, c_int, I64, P, P, P, c_float, P]),
}


class B200ADError(RuntimeError):
    """A libb200ad.so entry returned a negative status."""

    def __init__(self, code, message):
        super().__init__("libb200ad error %d: %s" % (code, message))
        self.code = code
        self.message = message


_lib = None


def load():
    """Load the shared library (once) and attach the prototypes. Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libb200ad.so not found at %s — build it with `python -m auto_differentiation_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.b200ad_abi_version() != 1:
        raise ImportError("libb200ad.so ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().b200ad_last_error().decode("utf-8", "replace")


def check(rc):
    if rc < 0:
        raise B200ADError(rc, last_error())
    return rc


def launch_count():
    return int(load().b200ad_launch_count())
