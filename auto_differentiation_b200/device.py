"""Device array + functional kernel wrappers: the host-side seam between the reference-shaped Python
API (tape ops, layers, optimizers) and the C ABI of libb200ad.so.

`DeviceArray` plays the role CuPy's ndarray plays for the reference (`import cupy as np`,
layers/layers.py:6): an fp32 (or int64 index) array living in HBM.  PyTorch is used ONLY as the
allocator / stream / view-metadata provider (torch.empty, Tensor.view/permute, .data_ptr()); every
arithmetic result is produced by a kernel of libb200ad.so.  There is no CPU fallback.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib as L

_F32 = torch.float32
_I64 = torch.int64


def _prod(xs):
    p = 1
    for x in xs:
        p *= int(x)
    return p


_DEVICE = "cuda"  # where buffers live; the package itself only ever computes on a CUDA device


_cuda_ok = [False]


def require_cuda():
    if _cuda_ok[0]:
        return
    if not torch.cuda.is_available():
        raise RuntimeError("auto_differentiation_b200: no CUDA device — the B200 path has no CPU fallback")
    _cuda_ok[0] = True        # availability cannot change within a process: skip the (slow) query on every allocation


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the launch stream of every kernel).  Called once per
    launch: goes through torch's C entry points directly (the Python wrappers cost ~1.5 us each, which shows on the
    launch-latency-bound paths)."""
    if _raw_stream is not None and _raw_device is not None and _cuda_ok[0]:
        return _raw_stream(_raw_device())
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def to_device(t):
    return t.to(_DEVICE)


def raw_bytes(n):
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=_DEVICE)


# GEMM workspaces: launches of one stream are ordered, so consecutive products of a stream share ONE grow-only buffer instead
# of a torch.empty per call (~4 us of host time each on the launch-bound paths).  Outgrown buffers stay referenced (a CUDA
# graph may have baked their address in; sizes grow geometrically, so the total stays below twice the largest).  While a
# stream is being captured the allocation goes through torch (the graph's private pool), as before.
_ws_cache = {}
_ws_keep = []


def _workspace(nbytes, stream):
    if _DEVICE != "cpu" and torch.cuda.is_current_stream_capturing():
        return raw_bytes(nbytes)
    t = _ws_cache.get(stream)
    if t is None or t.numel() < nbytes:
        if t is not None:
            _ws_keep.append(t)
        size = 1 << max(int(nbytes) - 1, 1).bit_length()
        t = _ws_cache[stream] = torch.empty(size, dtype=torch.uint8, device=_DEVICE)
    return t


def _lib():
    return L.load()


# --------------------------------------------------------------------------------------------------
# allocation
# --------------------------------------------------------------------------------------------------
def empty(shape, dtype=_F32):
    require_cuda()
    if isinstance(shape, int):
        shape = (shape,)
    return DeviceArray(torch.empty(tuple(int(s) for s in shape), dtype=dtype, device=_DEVICE))


def zeros(shape, dtype=_F32):
    out = empty(shape, dtype)
    if dtype == _F32:
        fill_(out, 0.0)
    else:
        out.t.zero_()  # index buffers only (allocator-level memset, no arithmetic)
    return out


def zeros_like(a):
    return zeros(a.shape)


def full(shape, value):
    out = empty(shape)
    fill_(out, float(value))
    return out


def asarray(x, dtype=None):
    """Host (NumPy / list / scalar) or device input -> DeviceArray. Floats become fp32, ints int64
    only when asked (dtype='index'); this mirrors `cp.asarray`."""
    if isinstance(x, DeviceArray):
        return x
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = to_device(x)
        if dtype == "index":
            return DeviceArray(t.to(_I64))
        return DeviceArray(t if t.dtype == _F32 else t.to(_F32))
    a = np.asarray(x)
    if dtype == "index":
        a = np.ascontiguousarray(a, dtype=np.int64)
    else:
        a = np.ascontiguousarray(a, dtype=np.float32)
    return DeviceArray(to_device(torch.from_numpy(a)))


def as_index(x):
    if isinstance(x, DeviceArray):
        if x.t.dtype == _I64:
            return x
        raise TypeError("index arrays must be int64 DeviceArrays or host integer arrays")
    return asarray(x, dtype="index")


def is_scalar(x):
    return isinstance(x, (int, float, np.integer, np.floating)) or (isinstance(x, np.ndarray) and x.ndim == 0)


# --------------------------------------------------------------------------------------------------
# the array type
# --------------------------------------------------------------------------------------------------
class DeviceArray:
    __slots__ = ("t", "__weakref__")
    __array_priority__ = 1000.0

    def __init__(self, t):
        self.t = t

    # ---- metadata -------------------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return np.dtype(np.float32) if self.t.dtype == _F32 else np.dtype(np.int64)

    @property
    def T(self):
        n = self.t.dim()
        if n < 2:
            return self
        return DeviceArray(self.t.permute(*range(n - 1, -1, -1)))

    @property
    def ptr(self):
        return self.t.data_ptr()

    def __len__(self):
        if self.t.dim() == 0:
            raise TypeError("len() of unsized object")
        return self.t.shape[0]

    # ---- views / layout ---------------------------------------------------------------------------
    def transpose(self, *axes):
        if len(axes) == 0 or (len(axes) == 1 and axes[0] is None):
            return self.T
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        return DeviceArray(self.t.permute(*[int(a) for a in axes]))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = tuple(int(s) for s in shape)
        try:
            return DeviceArray(self.t.view(shape))
        except RuntimeError:
            return DeviceArray(self.contiguous().t.view(shape))

    def contiguous(self):
        if self.t.is_contiguous():
            return self
        out = empty(self.shape, self.t.dtype)
        if self.t.dtype != _F32:
            out.t.copy_(self.t)  # index arrays: plain copy, no arithmetic
            return out
        strided_copy_(self, out)
        return out

    def copy(self):
        out = empty(self.shape, self.t.dtype)
        if self.t.dtype != _F32:
            out.t.copy_(self.t)
        elif self.t.is_contiguous():
            L.check(_lib().b200ad_axpby(self.ptr, out.ptr, self.size, 1.0, 0.0, stream_ptr()))
        else:
            strided_copy_(self, out)
        return out

    def astype(self, dtype, copy=True):
        """Only the dtype the array already has is accepted (fp32 data / int64 indices; float64 maps to fp32 like every
        upload does): a silent no-op conversion would hand back the wrong type."""
        want = np.dtype(dtype) if not isinstance(dtype, torch.dtype) else None
        if isinstance(dtype, torch.dtype):
            same = dtype == self.t.dtype
        elif self.t.dtype == _F32:
            same = want.kind == "f"
        else:
            same = want == np.dtype(np.int64)
        if not same:
            raise TypeError("DeviceArray.astype(%s): device arrays are fp32 (data) or int64 (indices) only; this one is %s"
                            % (dtype, self.t.dtype))
        return self.copy() if copy else self

    def __getitem__(self, idx):
        if _is_basic_index(idx):
            return DeviceArray(self.t[idx])
        if isinstance(idx, (DeviceArray, np.ndarray, list)) and self.t.dim() == 2:
            return embedding_gather(self, as_index(idx))
        raise TypeError("DeviceArray supports basic indexing and W[int_array] row gathers only")

    def __setitem__(self, idx, value):
        whole = idx is Ellipsis or (isinstance(idx, slice) and idx == slice(None))
        target = self if whole else self[idx]
        if is_scalar(value):
            if target.t.is_contiguous():
                fill_(target, float(value))
            else:
                strided_copy_(DeviceArray(full((), value).t.expand(target.shape)), target)
            return
        src = asarray(value)
        if src.shape != target.shape:
            src = broadcast_to(src, target.shape)
        copy_into_(src, target)

    # ---- host access ------------------------------------------------------------------------------
    def get(self):
        return self.contiguous().t.cpu().numpy()

    numpy = get

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a if dtype is None else a.astype(dtype)

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.get().reshape(-1)[0].item()

    def __float__(self):
        return float(self.item())

    def __int__(self):
        return int(self.item())

    def __repr__(self):
        return "DeviceArray(%s, shape=%s)" % (np.array2string(self.get(), separator=", "), self.shape)

    # ---- arithmetic (each one is a libb200ad kernel) ------------------------------------------------
    def __add__(self, o):
        return binary(L.B_ADD, self, o)

    def __radd__(self, o):
        return binary(L.B_ADD, o, self)

    def __sub__(self, o):
        return binary(L.B_SUB, self, o)

    def __rsub__(self, o):
        return binary(L.B_SUB, o, self)

    def __mul__(self, o):
        return binary(L.B_MUL, self, o)

    def __rmul__(self, o):
        return binary(L.B_MUL, o, self)

    def __truediv__(self, o):
        return binary(L.B_DIV, self, o)

    def __rtruediv__(self, o):
        return binary(L.B_DIV, o, self)

    def __neg__(self):
        return unary(L.U_NEG, self)

    def __pow__(self, p):
        if not is_scalar(p):
            raise TypeError("DeviceArray ** exponent: scalar exponents only")
        return unary(L.U_SQUARE, self) if float(p) == 2.0 else unary(L.U_POW, self, float(p))

    def __matmul__(self, o):
        return matmul(self, asarray(o))

    def __rmatmul__(self, o):
        return matmul(asarray(o), self)

    def _inplace(self, op, o):
        res = binary(op, self, o)
        if res.shape != self.shape:
            raise ValueError("non-broadcastable output operand with shape %s" % (self.shape,))
        copy_into_(res, self)
        return self

    def __iadd__(self, o):
        return self._inplace(L.B_ADD, o)

    def __isub__(self, o):
        return self._inplace(L.B_SUB, o)

    def __imul__(self, o):
        return self._inplace(L.B_MUL, o)

    def __itruediv__(self, o):
        return self._inplace(L.B_DIV, o)

    def sum(self, axis=None, keepdims=False):
        return sum_axis(self, axis, keepdims)

    def argmax(self, axis=None):
        if axis is None:
            return argext_flat(self, False)
        ax = axis % self.ndim
        if ax != self.ndim - 1:
            raise NotImplementedError("argmax over a non-last axis")
        rows = _prod(self.shape[:-1])
        return argmax_rows(self.reshape(rows, self.shape[-1])).reshape(self.shape[:-1])


def _is_basic_index(idx):
    items = idx if isinstance(idx, tuple) else (idx,)
    for it in items:
        if not (isinstance(it, (int, slice, np.integer)) or it is Ellipsis or it is None):
            return False
    return True


# --------------------------------------------------------------------------------------------------
# raw helpers
# --------------------------------------------------------------------------------------------------
def fill_(a, value):
    L.check(_lib().b200ad_fill(a.ptr, a.size, float(value), stream_ptr()))
    return a


def strided_copy_(src, dst):
    """dst = src for views of the same shape (either side may be strided)."""
    rank = src.ndim
    if rank > 6:
        raise NotImplementedError("strided copy of rank > 6 views")
    shp = (ctypes.c_int64 * 6)(*([int(s) for s in src.shape] + [1] * (6 - rank)))
    std = (ctypes.c_int64 * 6)(*([int(s) for s in src.t.stride()] + [0] * (6 - rank)))
    ostd = None
    if not dst.t.is_contiguous():
        ostd = (ctypes.c_int64 * 6)(*([int(s) for s in dst.t.stride()] + [0] * (6 - rank)))
    L.check(_lib().b200ad_strided_copy(src.ptr, dst.ptr, rank, shp, std, ostd, stream_ptr()))
    return dst


def copy_into_(src, dst):
    """dst[...] = src  (same shape)."""
    if src.shape != dst.shape:
        raise ValueError("could not broadcast input array from shape %s into shape %s" % (src.shape, dst.shape))
    if not dst.t.is_contiguous():
        return strided_copy_(src, dst)
    if src.t.is_contiguous():
        if src.ptr != dst.ptr:
            L.check(_lib().b200ad_axpby(src.ptr, dst.ptr, dst.size, 1.0, 0.0, stream_ptr()))
    else:
        strided_copy_(src, dst)
    return dst


def axpby_(x, y, a, b):
    """y = a*x + b*y in place (both contiguous, same size)."""
    L.check(_lib().b200ad_axpby(x.ptr, y.ptr, y.size, float(a), float(b), stream_ptr()))
    return y


def dropout_fwd(x, u, ratio, scale):
    """(y, mask): mask = (u > ratio) * scale, y = x * mask in one pass (reference TimeDropout.forward
    layers.py:685-692); `u` holds the uniform draws, same size as x."""
    x = x.contiguous()
    u = u.contiguous()
    if u.size != x.size:
        raise ValueError("dropout: %d uniform draws for %d elements" % (u.size, x.size))
    y, mask = empty(x.shape), empty(x.shape)
    L.check(_lib().b200ad_dropout_fwd(x.ptr, u.ptr, mask.ptr, y.ptr, x.size, float(ratio), float(scale), stream_ptr()))
    return y, mask


_rng_state = {"seed": 0x1234ABCD5678EF01, "offset": 0}


def seed_uniform(seed):
    """Re-seed the device dropout stream (Philox counter restarts at 0)."""
    _rng_state["seed"], _rng_state["offset"] = int(seed) & 0xFFFFFFFFFFFFFFFF, 0


def uniform(shape):
    """U[0,1) draws on the device from the library's own counter-based generator (b200ad_uniform, Philox4x32-10)."""
    out = empty(tuple(int(s) for s in shape))
    n = out.size
    L.check(_lib().b200ad_uniform(out.ptr, n, _rng_state["seed"], _rng_state["offset"], stream_ptr()))
    _rng_state["offset"] += (n + 3) // 4
    return out


def broadcast_to(a, shape):
    shape = tuple(int(s) for s in shape)
    if a.shape == shape:
        return a
    out = empty(shape)
    strided_copy_(DeviceArray(a.t.expand(shape)), out)
    return out


# --------------------------------------------------------------------------------------------------
# elementwise
# --------------------------------------------------------------------------------------------------
def unary(op, x, param=0.0):
    x = x.contiguous()
    y = empty(x.shape)
    L.check(_lib().b200ad_unary_fwd(op, x.ptr, y.ptr, x.size, float(param), stream_ptr()))
    return y


_BWD_USES = {}


def unary_bwd(op, g, x=None, y=None, param=0.0):
    uses = _BWD_USES.get(op)
    if uses is None:
        uses = _BWD_USES[op] = _lib().b200ad_unary_bwd_uses(op)
    g = g.contiguous()
    src = None
    if uses == 0:
        src = x.contiguous()
    elif uses == 1:
        src = y.contiguous()
    if src is not None and src.shape != g.shape:
        src = broadcast_to(src, g.shape)
    out = empty(g.shape)
    L.check(_lib().b200ad_unary_bwd(op, g.ptr, src.ptr if uses == 0 else None, src.ptr if uses == 1 else None,
                                    out.ptr, g.size, float(param), stream_ptr()))
    return out


def _classify(shape, out_shape):
    """How an operand of `shape` indexes into an output of `out_shape` (after NumPy left-padding).
    Returns (mode, period) or None when the pattern needs materialising."""
    n = len(out_shape)
    pad = (1,) * (n - len(shape)) + tuple(shape)
    if pad == tuple(out_shape):
        return (L.FULL, 1)
    if _prod(pad) == 1:
        return (L.SCALAR, 1)
    # ROW: leading ones then an exact match of the trailing extents
    k = 0
    while k < n and pad[k] == 1:
        k += 1
    if pad[k:] == tuple(out_shape[k:]):
        return (L.ROW, _prod(out_shape[k:]))
    # COL: exact leading extents then trailing ones
    k = n
    while k > 0 and pad[k - 1] == 1:
        k -= 1
    if pad[:k] == tuple(out_shape[:k]):
        return (L.COL, _prod(out_shape[k:]))
    return None


def _plan_binary(a, b):
    out_shape = tuple(np.broadcast_shapes(a.shape, b.shape))
    ca, cb = _classify(a.shape, out_shape), _classify(b.shape, out_shape)
    if ca is None:
        a = broadcast_to(a, out_shape)
        ca = (L.FULL, 1)
    if cb is None:
        b = broadcast_to(b, out_shape)
        cb = (L.FULL, 1)
    pa = ca[1] if ca[0] in (L.ROW, L.COL) else None
    pb = cb[1] if cb[0] in (L.ROW, L.COL) else None
    if pa is not None and pb is not None and pa != pb:
        b = broadcast_to(b, out_shape)
        cb, pb = (L.FULL, 1), None
    period = pa if pa is not None else (pb if pb is not None else 1)
    return a, b, out_shape, ca[0], cb[0], period


_CONST_OPS = {  # (op, scalar side) -> unary op
    (L.B_ADD, "b"): L.U_ADDC, (L.B_ADD, "a"): L.U_ADDC,
    (L.B_MUL, "b"): L.U_SCALE, (L.B_MUL, "a"): L.U_SCALE,
    (L.B_DIV, "b"): L.U_DIVC, (L.B_DIV, "a"): L.U_RDIVC,
    (L.B_SUB, "a"): L.U_RSUBC,
}


def binary(op, a, b):
    """NumPy-broadcasting a (op) b. Host scalars are folded into a one-input kernel."""
    if is_scalar(b):
        if op == L.B_SUB:
            return unary(L.U_ADDC, asarray(a), -float(b))
        return unary(_CONST_OPS[(op, "b")], asarray(a), float(b))
    if is_scalar(a):
        return unary(_CONST_OPS[(op, "a")], asarray(b), float(a))
    a, b = asarray(a).contiguous(), asarray(b).contiguous()
    a, b, out_shape, ma, mb, period = _plan_binary(a, b)
    out = empty(out_shape)
    L.check(_lib().b200ad_binary_fwd(op, a.ptr, b.ptr, out.ptr, out.size, ma, mb, period, stream_ptr()))
    return out


def binary_bwd(op, g, a, b, need_a=True, need_b=True, ref_compat=True):
    """Operand gradients of a (op) b, already reduced to the operand shapes (check_shape,
    tgrad.py:2163-2187, in its general correct form)."""
    g = g.contiguous()
    a, b = a.contiguous(), b.contiguous()
    if op in (L.B_ADD, L.B_SUB):
        # no operand values needed: (g, g) / REF_COMPAT (-g, +g) tgrad.py:754
        sa, sb = 1.0, 1.0
        if op == L.B_SUB:
            sa, sb = (-1.0, 1.0) if ref_compat else (1.0, -1.0)
        ga = gb = None
        if need_a:
            ga = reduce_to_shape(g, a.shape, scale=sa)
        if need_b:
            gb = reduce_to_shape(g, b.shape, scale=sb)
        return ga, gb
    a2, b2, out_shape, ma, mb, period = _plan_binary(a, b)
    if tuple(out_shape) != g.shape:
        raise ValueError("upstream gradient shape %s does not match the op output %s" % (g.shape, out_shape))
    # fused single pass for the dominant pattern: a full, b a row vector
    if need_b and ma == L.FULL and mb == L.ROW and period >= 32:
        rows = g.size // period
        ga = empty(out_shape) if need_a else None
        gb = empty(b.shape)
        nws = int(_lib().b200ad_colsum_ws_floats(rows, period))
        ws = empty(max(nws, 1))
        L.check(_lib().b200ad_binary_bwd_rowfused(op, g.ptr, a2.ptr, b2.ptr, ga.ptr if need_a else None, gb.ptr,
                                                  rows, period, 1 if ref_compat else 0, ws.ptr, stream_ptr()))
        return ga, gb
    ga_full = empty(out_shape) if need_a else None
    gb_full = empty(out_shape) if need_b else None
    L.check(_lib().b200ad_binary_bwd(op, g.ptr, a2.ptr, b2.ptr, ga_full.ptr if need_a else None,
                                     gb_full.ptr if need_b else None, g.size, ma, mb, period,
                                     1 if ref_compat else 0, stream_ptr()))
    ga = reduce_to_shape(ga_full, a.shape) if need_a else None
    gb = reduce_to_shape(gb_full, b.shape) if need_b else None
    return ga, gb


# --------------------------------------------------------------------------------------------------
# reductions
# --------------------------------------------------------------------------------------------------
def reduce_mid(x, outer, red, inner, scale=1.0):
    """sum over the middle axis of a contiguous (outer, red, inner) view -> flat (outer*inner,)"""
    out = empty((outer * inner,))
    nws = int(_lib().b200ad_reduce_ws_floats(outer, red, inner))
    ws = empty(max(nws, 1))
    L.check(_lib().b200ad_reduce(x.ptr, out.ptr, outer, red, inner, float(scale), ws.ptr, stream_ptr()))
    return out


def reduce_to_shape(g, shape, scale=1.0):
    """Un-broadcast `g` to `shape` by summing (the general, correct check_shape)."""
    shape = tuple(int(s) for s in shape)
    gs = g.shape
    if gs == shape:
        return g if scale == 1.0 else unary(L.U_SCALE, g, scale)
    if len(gs) < len(shape):
        raise RuntimeError("grad shape %s smaller than %s" % (gs, shape))  # tgrad.py:2179
    pad = (1,) * (len(gs) - len(shape)) + shape
    for gd, td in zip(gs, pad):
        if td != gd and td != 1:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (gs, shape))
    g = g.contiguous()
    cur = list(gs)
    runs, i = [], 0
    while i < len(cur):
        if pad[i] == 1 and cur[i] != 1:
            j = i
            while j + 1 < len(cur) and pad[j + 1] == 1:
                j += 1
            runs.append((i, j))
            i = j + 1
        else:
            i += 1
    first = True
    for (i, j) in reversed(runs):
        outer, red, inner = _prod(cur[:i]), _prod(cur[i:j + 1]), _prod(cur[j + 1:])
        g = reduce_mid(g, outer, red, inner, scale if first else 1.0)
        first = False
        for k in range(i, j + 1):
            cur[k] = 1
    if first and scale != 1.0:
        g = unary(L.U_SCALE, g, scale)
    return g.reshape(shape)


def sum_axis(x, axis=None, keepdims=False):
    x = x.contiguous()
    shp = x.shape
    if axis is None:
        out = reduce_mid(x, 1, x.size, 1)
        return out.reshape((1,) * len(shp) if keepdims else ())
    ax = int(axis)
    if ax < 0:
        ax += len(shp)
    if not 0 <= ax < len(shp):
        raise ValueError("axis %d is out of bounds for array of dimension %d" % (axis, len(shp)))
    outer, red, inner = _prod(shp[:ax]), shp[ax], _prod(shp[ax + 1:])
    out = reduce_mid(x, outer, red, inner)
    new_shape = shp[:ax] + ((1,) if keepdims else ()) + shp[ax + 1:]
    return out.reshape(new_shape)


def repeat_axis(x, times, axis, scale=1.0):
    """np.repeat(x, times, axis) for an axis of extent 1 generalised: every slice along `axis` is
    repeated `times` times consecutively (repeat.result tgrad.py:795-797)."""
    x = x.contiguous()
    shp = x.shape
    ax = int(axis)
    if ax < 0:
        ax += len(shp)
    outer, inner = _prod(shp[:ax + 1]), _prod(shp[ax + 1:])
    out = empty(shp[:ax] + (shp[ax] * int(times),) + shp[ax + 1:])
    L.check(_lib().b200ad_repeat_axis(x.ptr, out.ptr, outer, int(times), inner, float(scale), stream_ptr()))
    return out


# --------------------------------------------------------------------------------------------------
# softmax / cross-entropy / argmax
# --------------------------------------------------------------------------------------------------
def softmax_lastaxis(x, in_scale=1.0):
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.size // cols if cols else 0
    y = empty(x.shape)
    L.check(_lib().b200ad_softmax_fwd(x.ptr, y.ptr, rows, cols, float(in_scale), stream_ptr()))
    return y


def softmax_bwd_lastaxis(g, y, out_scale=1.0, diag_only=False, out=None):
    g, y = g.contiguous(), y.contiguous()
    cols = y.shape[-1]
    rows = y.size // cols if cols else 0
    gx = out if out is not None else empty(y.shape)
    L.check(_lib().b200ad_softmax_bwd(g.ptr, y.ptr, gx.ptr, rows, cols, float(out_scale), 1 if diag_only else 0,
                                      stream_ptr()))
    return gx


def xent_fwd(logits, labels, ignore_label=-1, eps=0.0, denom_mode=0, input_is_probs=False, want_probs=True):
    """Returns (loss_and_count [2] device array, probs or None). labels: int64 DeviceArray (rows,)."""
    logits = logits.contiguous()
    rows, cols = logits.shape
    probs = empty((rows, cols)) if want_probs else None
    row_loss = empty((rows,))
    lc = empty((2,))
    L.check(_lib().b200ad_xent_fwd(logits.ptr, labels.ptr, probs.ptr if want_probs else None, row_loss.ptr, lc.ptr,
                                   rows, cols, int(ignore_label), float(eps), int(denom_mode),
                                   1 if input_is_probs else 0, stream_ptr()))
    return lc, probs


def xent_bwd(probs, labels, lc, dout=1.0, ignore_label=-1, out=None):
    rows, cols = probs.shape
    dx = out if out is not None else empty((rows, cols))
    L.check(_lib().b200ad_xent_bwd(probs.ptr, labels.ptr, lc.ptr, dx.ptr, rows, cols, int(ignore_label), float(dout),
                                   0, stream_ptr()))
    return dx


def argext_flat(x, is_min=False):
    x = x.contiguous()
    out = empty((), _I64)
    ws = raw_bytes(L.ARG_WS_BYTES)
    L.check(_lib().b200ad_argext_flat(x.ptr, x.size, 1 if is_min else 0, out.ptr, ws.data_ptr(), stream_ptr()))
    return out


def argmax_rows(x):
    x = x.contiguous()
    rows, cols = x.shape
    out = empty((rows,), _I64)
    L.check(_lib().b200ad_argmax_rows(x.ptr, rows, cols, out.ptr, stream_ptr()))
    return out


def embedding_gather(w, idx):
    w = w.contiguous()
    idx = idx.contiguous()
    vocab, dim = w.shape
    out = empty(idx.shape + (dim,))
    L.check(_lib().b200ad_embedding_gather(w.ptr, idx.ptr, out.ptr, idx.size, vocab, dim, stream_ptr()))
    return out


def embedding_scatter_add_(dout, idx, dw):
    dout, idx = dout.contiguous(), idx.contiguous()
    vocab, dim = dw.shape
    L.check(_lib().b200ad_embedding_scatter_add(dout.ptr, idx.ptr, dw.ptr, idx.size, vocab, dim, stream_ptr()))
    return dw


# --------------------------------------------------------------------------------------------------
# GEMM on (possibly strided) views
# --------------------------------------------------------------------------------------------------
_GEMM_PRECISION = {"auto": L.GEMM_AUTO, "fp32": L.GEMM_FP32, "tf32x3": L.GEMM_TF32X3, "bf16": L.GEMM_BF16,
                   "tf32": L.GEMM_TF32, "f16x3": L.GEMM_F16X3}
_default_precision = [L.GEMM_AUTO]


def set_gemm_precision(name):
    """'auto' (fp32-faithful: tcgen05 3xTF32 where the shape allows, else fp32 SIMT), 'fp32',
    'tf32x3', 'f16x3' (fp32-faithful like tf32x3: fp16 hi/lo pairs with row scales at the kind::f16 MMA rate),
    'bf16' (tensor-core bf16 inputs / fp32 accumulate — looser tolerance)."""
    _default_precision[0] = _GEMM_PRECISION[name]


def set_f16x3_auto(mode):
    """When 'auto' resolves to 'f16x3' (b200ad_gemm_set_f16x3_auto): 0 never, 1 by the library's device-time model
    (default), 2 whenever the path supports the problem.  Returns the previous mode."""
    _matmul_plans.clear()                  # cached plans carry the workspace size of the path they were planned for
    _shared_a_plans.clear()
    return L.check(_lib().b200ad_gemm_set_f16x3_auto(int(mode)))


def get_gemm_precision():
    return _default_precision[0]


class GemmTimer:
    """Optional per-launch CUDA-event timing of the GEMM launches (bench.py's roofline leg): events are
    recorded on the launching stream right before/after each b200ad_gemm call."""

    def __init__(self):
        self.records = []   # (key, flops, start_event, end_event)

    def summary(self):
        """{key: (launches, total_ms, flops_per_launch)} — call after a device synchronize."""
        out = {}
        for key, flops, e0, e1 in self.records:
            n, ms, _ = out.get(key, (0, 0.0, flops))
            out[key] = (n + 1, ms + e0.elapsed_time(e1), flops)
        return out


_gemm_timer = [None]


def set_gemm_timer(timer):
    _gemm_timer[0] = timer


def _mat_layout(t, rows, cols):
    """(trans, ld) for the last two axes of torch view `t`, or None if it must be copied."""
    sr, sc = t.stride(-2), t.stride(-1)
    if sc == 1 or cols == 1:
        ld = sr if rows > 1 else max(cols, 1)
        if ld >= max(cols, 1):
            return 0, ld
    if sr == 1 or rows == 1:
        ld = sc if cols > 1 else max(rows, 1)
        if ld >= max(rows, 1):
            return 1, ld
    return None


def _batch_plan(bshape, strides_list):
    dims = [(int(n), [int(s[i]) for s in strides_list]) for i, n in enumerate(bshape) if n != 1]
    merged = []
    for n, st in dims:
        if merged:
            pn, pst = merged[-1]
            if all(ps == s * n for ps, s in zip(pst, st)):
                merged[-1] = (pn * n, st)
                continue
        merged.append((n, st))
    return merged


_matmul_plans = {}


def matmul(a, b, out=None, alpha=1.0, beta=0.0, bias=None, act=L.ACT_NONE, precision=None, trans_a=False,
           trans_b=False):
    """np.matmul on device views: (..., M, K) @ (..., K, N) with broadcast batch axes.
    `trans_a/trans_b` apply a logical transpose of the last two axes first (free: folded into the
    kernel's operand layout).  `out` may be any view whose last axis is contiguous (written in place,
    `beta` accumulates)."""
    # ---- fast path: the same operand geometry was planned before (layers call the same products every step; the
    # planning below costs ~10 us of Python per call, which bounds the launch-latency-bound configs)
    ta, tb = a.t, b.t
    key = (ta.shape, ta.stride(), tb.shape, tb.stride(), None if out is None else (out.t.shape, out.t.stride()),
           trans_a, trans_b, act, precision, alpha, beta, bias is None, _default_precision[0])
    hit = _matmul_plans.get(key)
    if hit is not None:
        d, ws_bytes, out_shape, tkey, flops = hit
        if out is None:
            out = DeviceArray(torch.empty(out_shape, dtype=_F32, device=_DEVICE))
        sp = stream_ptr()
        ws = _workspace(ws_bytes, sp) if ws_bytes > 0 else None
        timer = _gemm_timer[0]
        if timer is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        L.check(_lib().b200ad_gemm(ctypes.byref(d), ta.data_ptr(), tb.data_ptr(), out.t.data_ptr(),
                                   bias.ptr if bias is not None else None, ws.data_ptr() if ws is not None else None,
                                   ws_bytes, sp))
        if timer is not None:
            e1.record()
            timer.records.append((tkey, flops, e0, e1))
        return out
    cacheable = True
    if ta.dim() == 1:
        ta = ta.unsqueeze(0)
    if tb.dim() == 1:
        tb = tb.unsqueeze(1)
    if trans_a:
        ta = ta.transpose(-1, -2)
    if trans_b:
        tb = tb.transpose(-1, -2)
    M, K = ta.shape[-2], ta.shape[-1]
    K2, N = tb.shape[-2], tb.shape[-1]
    if K != K2:
        raise ValueError("matmul: Input operand 1 has a mismatch in its core dimension 0 (size %d is different "
                         "from %d)" % (K2, K))
    la = _mat_layout(ta, M, K)
    if la is None:
        ta = DeviceArray(ta).contiguous().t
        la = (0, max(K, 1))
        cacheable = False                  # an operand had to be compacted first: not a pure view plan
    lb = _mat_layout(tb, K, N)
    if lb is None:
        tb = DeviceArray(tb).contiguous().t
        lb = (0, max(N, 1))
        cacheable = False
    bshape = tuple(np.broadcast_shapes(tuple(ta.shape[:-2]), tuple(tb.shape[:-2])))
    nb = len(bshape)

    def bstrides(t):
        shp, st = tuple(t.shape[:-2]), tuple(t.stride()[:-2])
        off = nb - len(shp)
        return [0] * off + [0 if shp[i] == 1 else st[i] for i in range(len(shp))]

    if out is None:
        out = empty(bshape + (M, N))
    tc = out.t
    if tuple(tc.shape) != bshape + (M, N):
        raise ValueError("matmul: out has shape %s, expected %s" % (tuple(tc.shape), bshape + (M, N)))
    if N > 1 and tc.stride(-1) != 1:
        raise ValueError("matmul: out must have a contiguous last axis")
    ldc = tc.stride(-2) if M > 1 else max(N, 1)
    if ldc < N:
        raise ValueError("matmul: out rows overlap")
    plan = _batch_plan(bshape, [bstrides(ta), bstrides(tb), list(tc.stride()[:-2])])
    prec = _default_precision[0] if precision is None else _GEMM_PRECISION.get(precision, precision)
    lib = _lib()
    s = stream_ptr()
    d = L.GemmDesc()
    d.precision, d.trans_a, d.trans_b, d.act = prec, la[0], lb[0], act
    d.m, d.n, d.k = M, N, K
    d.lda, d.ldb, d.ldc = la[1], lb[1], ldc
    d.alpha, d.beta = float(alpha), float(beta)
    bias_ptr = bias.ptr if bias is not None else None
    inner = plan[-2:] if len(plan) >= 2 else plan
    outer = plan[:-2] if len(plan) > 2 else []
    dims = [(1, [0, 0, 0])] * (2 - len(inner)) + inner
    d.batch0, d.batch1 = dims[0][0], dims[1][0]
    d.a_stride0, d.b_stride0, d.c_stride0 = dims[0][1]
    d.a_stride1, d.b_stride1, d.c_stride1 = dims[1][1]
    ws_bytes = int(lib.b200ad_gemm_workspace_bytes(ctypes.byref(d)))
    ws = _workspace(ws_bytes, s) if ws_bytes > 0 else None
    wptr = ws.data_ptr() if ws is not None else None
    pa, pb, pc = ta.data_ptr(), tb.data_ptr(), tc.data_ptr()
    if cacheable and not outer and a.t.dim() >= 2 and b.t.dim() >= 2:
        if len(_matmul_plans) > 4096:
            _matmul_plans.clear()
        _matmul_plans[key] = (d, ws_bytes, tuple(tc.shape), (M, N, K, d.batch0 * d.batch1, la[0], lb[0], prec),
                              2.0 * M * N * K * d.batch0 * d.batch1)
    timer = _gemm_timer[0]
    if timer is not None and not outer:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.b200ad_gemm(ctypes.byref(d), pa, pb, pc, bias_ptr, wptr, ws_bytes, s))
        e1.record()
        timer.records.append(((M, N, K, d.batch0 * d.batch1, la[0], lb[0], prec),
                              2.0 * M * N * K * d.batch0 * d.batch1, e0, e1))
    elif not outer:
        L.check(lib.b200ad_gemm(ctypes.byref(d), pa, pb, pc, bias_ptr, wptr, ws_bytes, s))
    else:  # more than two non-mergeable batch axes: loop the outermost ones on the host
        import itertools
        for idx in itertools.product(*[range(n) for n, _ in outer]):
            oa = sum(i * st[0] for i, (_, st) in zip(idx, outer)) * 4
            ob = sum(i * st[1] for i, (_, st) in zip(idx, outer)) * 4
            oc = sum(i * st[2] for i, (_, st) in zip(idx, outer)) * 4
            L.check(lib.b200ad_gemm(ctypes.byref(d), pa + oa, pb + ob, pc + oc, bias_ptr, wptr, ws_bytes, s))
    return out


_shared_a_plans = {}


def matmul_shared_a(a, bs, precision=None):
    """[a @ b for b in bs] for 2-D operands that share their geometry, through ONE b200ad_gemm_shared_a call: on the
    fp16-split path the shared operand is split once for all products (the Q / K / V projections of self-attention,
    reference layers/layers.py:1347-1349).  Anything else is a loop of `matmul`."""
    ta, tbs = a.t, [b.t for b in bs]
    same = (len(tbs) > 1 and ta.dim() == 2 and all(t.dim() == 2 and t.shape == tbs[0].shape and t.stride() == tbs[0].stride()
                                                   for t in tbs) and ta.shape[1] == tbs[0].shape[0])
    if not same:
        return [matmul(a, b, precision=precision) for b in bs]
    key = (ta.shape, ta.stride(), tbs[0].shape, tbs[0].stride(), len(tbs), precision, _default_precision[0])
    hit = _shared_a_plans.get(key)
    if hit is None:
        M, K = ta.shape
        N = tbs[0].shape[1]
        la, lb = _mat_layout(ta, M, K), _mat_layout(tbs[0], K, N)
        if la is None or lb is None or M == 0 or N == 0 or K == 0:
            return [matmul(a, b, precision=precision) for b in bs]
        d = L.GemmDesc()
        d.precision = _default_precision[0] if precision is None else _GEMM_PRECISION.get(precision, precision)
        d.trans_a, d.trans_b, d.act = la[0], lb[0], L.ACT_NONE
        d.m, d.n, d.k = M, N, K
        d.lda, d.ldb, d.ldc = la[1], lb[1], N
        d.batch0 = d.batch1 = 1
        d.alpha, d.beta = 1.0, 0.0
        ws_bytes = int(_lib().b200ad_gemm_shared_a_workspace_bytes(ctypes.byref(d), len(tbs)))
        if len(_shared_a_plans) > 256:
            _shared_a_plans.clear()
        hit = _shared_a_plans[key] = (d, ws_bytes, (M, N), (M, N, K, 1, la[0], lb[0], d.precision), 2.0 * M * N * K)
    d, ws_bytes, out_shape, tkey, flops = hit
    outs = [DeviceArray(torch.empty(out_shape, dtype=_F32, device=_DEVICE)) for _ in tbs]
    n = len(tbs)
    arr = ctypes.c_void_p * n
    sp = stream_ptr()
    ws = _workspace(ws_bytes, sp) if ws_bytes > 0 else None
    timer = _gemm_timer[0]
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(_lib().b200ad_gemm_shared_a(ctypes.byref(d), ta.data_ptr(), n, arr(*[t.data_ptr() for t in tbs]),
                                        arr(*[o.t.data_ptr() for o in outs]), None,
                                        ws.data_ptr() if ws is not None else None, ws_bytes, sp))
    if timer is not None:
        e1.record()
        # one record per product so that launch counts and per-launch flops stay comparable with `matmul` records
        timer.records.append((tkey, flops * n, e0, e1))
        timer.shared_a_products = getattr(timer, "shared_a_products", 0) + n
    return outs


# --------------------------------------------------------------------------------------------------
# fused attention core on head-split views (B, h, S, d_k)
# --------------------------------------------------------------------------------------------------
_fused_attention = [True]


def set_fused_attention(flag):
    """False routes MultiHeadAttention through the unfused GEMM + softmax kernels (materialised S x S)."""
    _fused_attention[0] = bool(flag)


def _attn_strides(s, t):
    s.batch, s.head, s.row = (int(t.stride(0)), int(t.stride(1)), int(t.stride(2)))


_attn_descs = {}       # (shapes, strides, dtypes, scale) -> AttnDesc | None: layers call the same geometry every step


def _attn_desc(q, k, v, o, scale):
    """q/k/v/o: DeviceArray views (B, h, S, d) with a contiguous last axis, or None if that does not hold."""
    ts = [x.t for x in (q, k, v, o)]
    key = (tuple((t.shape, t.stride(), t.dtype) for t in ts), float(scale))
    if key in _attn_descs:
        return _attn_descs[key]
    if len(_attn_descs) > 512:
        _attn_descs.clear()
    d = _attn_descs[key] = _attn_desc_build(ts, scale)
    return d


def _attn_desc_build(ts, scale):
    if any(t.dim() != 4 or t.dtype != _F32 or (t.shape[3] > 1 and t.stride(3) != 1) for t in ts):
        return None
    B, H, Sq, dk = ts[0].shape
    Skv = ts[1].shape[2]
    if tuple(ts[1].shape) != (B, H, Skv, dk) or tuple(ts[2].shape) != (B, H, Skv, dk) or tuple(ts[3].shape) != (B, H, Sq, dk):
        return None
    d = L.AttnDesc()
    d.batch, d.heads, d.seq_q, d.seq_kv, d.head_dim = B, H, Sq, Skv, dk
    d.scale = float(scale)
    for s, t in zip((d.q, d.k, d.v, d.o), ts):
        _attn_strides(s, t)
    return d


def _same_strides(a, b):
    return all(n == 1 or sa == sb for n, sa, sb in zip(a.shape, a.t.stride(), b.t.stride()))


def attn_supported(q, k, v):
    """True when the fused tcgen05 attention pass can take these head-split views."""
    if not _fused_attention[0]:
        return False
    d = _attn_desc(q, k, v, q, 1.0)
    return d is not None and bool(_lib().b200ad_attn_supported(ctypes.byref(d)))


def _timed(key, flops, call):
    timer = _gemm_timer[0]
    if timer is None:
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    timer.records.append((key, flops, e0, e1))


class _Lse(DeviceArray):
    """The log-sum-exp array a forward attention pass returns, carrying that pass's workspace: the kernel leaves the slab
    maxima of Q / K / V in it, which saves the backward of the SAME q / k / v three tensor reads.  Views and copies are plain
    DeviceArrays (the backward then computes the maxima itself)."""
    __slots__ = ("fwd_ws",)


def attn_fwd(q, k, v, scale):
    """O = softmax(scale * q k^T) v per (batch, head).  Returns (O as a (B, h, S, d) view of a (B, S, h, d)
    buffer — i.e. already in merged-head layout —, lse (B, h, S) in the log2 domain)."""
    B, H, Sq, dk = q.shape
    o = empty((B, Sq, H, dk)).transpose(0, 2, 1, 3)
    lse = _Lse(empty((B, H, Sq)).t)
    d = _attn_desc(q, k, v, o, scale)
    s = stream_ptr()
    ws = empty((max(1, int(_lib().b200ad_attn_fwd_ws_floats(ctypes.byref(d)))),))
    _timed(("attn_fwd", B, H, Sq, k.shape[2], dk), 4.0 * B * H * Sq * k.shape[2] * dk,
           lambda: L.check(_lib().b200ad_attn_fwd(ctypes.byref(d), q.ptr, k.ptr, v.ptr, o.ptr, lse.ptr, ws.ptr, s)))
    lse.fwd_ws = ws
    return o, lse


def attn_bwd(q, k, v, o, dout, lse, scale):
    """Gradients (dq, dk, dv) of attn_fwd, each a (B, h, S, d) view of a fresh (B, S, h, d) buffer.
    `dout` must share `o`'s strides."""
    B, H, Sq, dk = q.shape
    Skv = k.shape[2]
    if not _same_strides(dout, o):
        tmp = empty((B, Sq, H, dk)).transpose(0, 2, 1, 3)
        strided_copy_(dout, tmp)
        dout = tmp
    d = _attn_desc(q, k, v, o, scale)

    def like(x):
        # same strides as x (the ABI writes dq/dk/dv with q/k/v's strides): allocate in x's memory order
        order = sorted(range(4), key=lambda i: -x.t.stride(i))
        buf = torch.empty([x.t.shape[i] for i in order], dtype=_F32, device=x.t.device)
        inv = [order.index(i) for i in range(4)]
        out = buf.permute(*inv)
        return DeviceArray(out)

    dq, dk_, dv = like(q), like(k), like(v)
    if not (_same_strides(dq, q) and _same_strides(dk_, k) and _same_strides(dv, v)):
        # q/k/v are slices of wider buffers; the ABI writes dq/dk/dv with q/k/v's strides -> compact the operands once
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        d = _attn_desc(q, k, v, o, scale)
        dq, dk_, dv = empty(q.shape), empty(k.shape), empty(v.shape)
    ws = empty((int(_lib().b200ad_attn_bwd_ws_floats(ctypes.byref(d))),))
    fwd_ws = getattr(lse, "fwd_ws", None)          # the forward's slab maxima, if `lse` is the object attn_fwd returned
    s = stream_ptr()
    # algorithmic flops: backward = 2 x forward = 8 Sq Skv dk per (batch, head); the S / dP recomputation
    # the passes do is issued work, not algorithmic work
    _timed(("attn_bwd", B, H, Sq, Skv, dk), 8.0 * B * H * Sq * Skv * dk,
           lambda: L.check(_lib().b200ad_attn_bwd_maxima(ctypes.byref(d), q.ptr, k.ptr, v.ptr, o.ptr, dout.ptr, lse.ptr,
                                                         fwd_ws.ptr if fwd_ws is not None else None,
                                                         dq.ptr, dk_.ptr, dv.ptr, ws.ptr, s)))
    return dq, dk_, dv


# --------------------------------------------------------------------------------------------------
# strided row-wise helpers used by the layers
# --------------------------------------------------------------------------------------------------
def _rows_ld(a):
    """(ptr, ld, rows, cols) of a 2-D view whose last axis is contiguous."""
    t = a.t
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("expected a 2-D view with a contiguous last axis")
    rows, cols = t.shape
    ld = t.stride(0) if rows > 1 else max(cols, 1)
    return t.data_ptr(), ld, rows, cols


def bias_act(x, bias, act, out=None):
    px, ldx, rows, cols = _rows_ld(x)
    if out is None:
        out = empty((rows, cols))
    po, ldo, _, _ = _rows_ld(out)
    L.check(_lib().b200ad_bias_act_fwd(px, ldx, bias.ptr if bias is not None else None, po, ldo, rows, cols, act,
                                       stream_ptr()))
    return out


def act_bwd(g, y, act, out=None):
    pg, ldg, rows, cols = _rows_ld(g)
    py, ldy, _, _ = _rows_ld(y)
    if out is None:
        out = empty((rows, cols))
    po, ldo, _, _ = _rows_ld(out)
    L.check(_lib().b200ad_act_bwd(pg, ldg, py, ldy, po, ldo, rows, cols, act, stream_ptr()))
    return out


# --------------------------------------------------------------------------------------------------
# persistent recurrent time loops (csrc/rnn_seq.cu)
# --------------------------------------------------------------------------------------------------
_rnn_seq_ok = {}
_persistent_recurrence = [True]


def set_persistent_recurrence(flag):
    """False forces the per-step loop over b200ad_gemm (the reference's own loop structure; used by parity tests)."""
    _persistent_recurrence[0] = bool(flag)


def rnn_seq_supported(batch, hidden):
    if not _persistent_recurrence[0]:
        return False
    key = (int(batch), int(hidden))
    if key not in _rnn_seq_ok:
        _rnn_seq_ok[key] = bool(_lib().b200ad_rnn_seq_supported(key[0], key[1]))
    return _rnn_seq_ok[key]


_lstm_seq_ok = {}


def lstm_seq_supported(batch, hidden):
    """Whether the LSTM time loops run from one C call (csrc/lstm_seq.cu) for this (batch, hidden)."""
    if not _persistent_recurrence[0]:
        return False
    key = (int(batch), int(hidden))
    if key not in _lstm_seq_ok:
        _lstm_seq_ok[key] = bool(_lib().b200ad_lstm_seq_supported(key[0], key[1]))
    return _lstm_seq_ok[key]


def _lstm_seq_ws(N, H):
    nbytes = int(_lib().b200ad_lstm_seq_workspace_bytes(N, H))
    return (empty(((nbytes + 3) // 4,)) if nbytes else None), nbytes


def lstm_seq_fwd_(hs_ext, cs_ext, gates, pre, wh):
    """hs_ext / cs_ext (T+1, N, H) with [0] = h_0 / c_0; pre (T, N, 4H) = x Wx + b (scratch afterwards); gates (T, N, 4H):
    all T steps of the LSTM recurrence from one C call (two launches per step)."""
    T1, N, H = hs_ext.shape
    wh = wh.contiguous()
    ws, nbytes = _lstm_seq_ws(N, H)
    L.check(_lib().b200ad_lstm_seq_fwd(hs_ext.ptr, cs_ext.ptr, gates.ptr, pre.ptr, wh.ptr, T1 - 1, N, H,
                                       ws.ptr if ws is not None else None, nbytes, stream_ptr()))
    return hs_ext


def lstm_seq_bwd(dhs, hs_ext, cs_ext, gates, wh):
    """BPTT of the LSTM recurrence from one C call.  dhs: (N, T, H) view with a contiguous last axis.  Returns
    (dA (T, N, 4H), dh0 (N, H))."""
    T1, N, H = hs_ext.shape
    T = T1 - 1
    wh = wh.contiguous()
    if dhs.t.stride(2) != 1 or (dhs.t.stride(0) & 3) or (dhs.t.stride(1) & 3) or (dhs.ptr & 15):
        dhs = dhs.copy()
    dA, dh0, dc = empty((T, N, 4 * H)), empty((N, H)), empty((2, N, H))
    ws, nbytes = _lstm_seq_ws(N, H)
    L.check(_lib().b200ad_lstm_seq_bwd(dhs.ptr, dhs.t.stride(0), dhs.t.stride(1), hs_ext.ptr, cs_ext.ptr, gates.ptr, wh.ptr,
                                       dA.ptr, dh0.ptr, dc.ptr, T, N, H, ws.ptr if ws is not None else None, nbytes,
                                       stream_ptr()))
    return dA, dh0


def rnn_tanh_seq_fwd_(hs_ext, pre, wh):
    """hs_ext (T+1, N, H) contiguous with hs_ext[0] = h_0; pre (T, N, H): hs_ext[t+1] = tanh(pre[t] + hs_ext[t] @ wh)
    for all t, in ONE launch."""
    T1, N, H = hs_ext.shape
    wh, pre = wh.contiguous(), pre.contiguous()
    L.check(_lib().b200ad_rnn_tanh_seq_fwd(hs_ext.ptr, pre.ptr, wh.ptr, T1 - 1, N, H, stream_ptr()))
    return hs_ext


def rnn_tanh_seq_bwd(dhs, hs_ext, wh):
    """BPTT of the tanh recurrence in ONE launch.  dhs: (N, T, H) view with a contiguous last axis.  Returns
    (dA (T, N, H), dh0 (N, H))."""
    T1, N, H = hs_ext.shape
    T = T1 - 1
    wh = wh.contiguous()
    if dhs.t.stride(2) != 1 or (dhs.t.stride(0) & 3) or (dhs.t.stride(1) & 3) or (dhs.ptr & 15):
        dhs = dhs.copy()
    dA, dh0 = empty((T, N, H)), empty((N, H))
    L.check(_lib().b200ad_rnn_tanh_seq_bwd(dhs.ptr, dhs.t.stride(0), dhs.t.stride(1), hs_ext.ptr, wh.ptr, dA.ptr, dh0.ptr,
                                           T, N, H, stream_ptr()))
    return dA, dh0
