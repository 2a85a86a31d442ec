"""Define-by-run autodiff tape with the surface of the reference's ``auto_grad/tgrad.py`` (old module
name: ``ad``), computed on a B200 through libb200ad.so.

Drop-in contract: same class names and constructor signatures; every op is a
``Function`` subclass whose ``result()`` (reference tgrad.py:48-52) runs eagerly in the constructor
and whose ``get_grad(grad)`` (:54-58) returns the operand gradient(s); ``root.gradient(upstream)``
(:81-91) fills ``leaf.grad`` and accumulates across calls (:1885).  ``.val`` is a ``DeviceArray``
(fp32 in HBM) — pass it wherever an array is expected, ``np.asarray(node.val)`` copies to the host.

Deliberate differences from the reference (recorded in DESIGN.md):
  * backward is ONE topological sweep with interior accumulation instead of the per-path LIFO DFS
    of ``Prime.prime`` (:1873-1951): same sums, no k-fold re-walk of shared sub-graphs;
  * constant sub-graphs (Python scalars, e.g. ``ad.sqrt(d_k)``) are evaluated on the host and get no
    gradient work; the reference spends 38 % of an MHA step computing gradients for them;
  * ``sub`` keeps the reference's sign-inverted gradients (:753-756) and ``Divide`` its ``+1e-8``
    (:1004-1005) while ``REF_COMPAT`` is True (default) — parity is the contract;
  * ``check_shape`` (:2163-2187) is the general reduce-to-shape; it coincides with the reference
    wherever the reference's heuristic is right;
  * ``transpose``/``tensordot`` backward apply the INVERSE permutation (the reference re-applies the
    forward one, :597/:228, which is only right for involutions);
  * ``create_graph=True`` (symbolic gradient graph + graphviz view) is out of scope.
"""
import math

import numpy as np

from .. import _lib as L
from .. import device as D
from ..device import DeviceArray

REF_COMPAT = True  # replicate sub's sign inversion (tgrad.py:754); Divide eps is always kept


def _is_node(x):
    return hasattr(x, "vars")


def _is_host_scalar(v):
    return not isinstance(v, DeviceArray)


def _value_of(x):
    """Operand -> DeviceArray, or a host scalar for constants."""
    if _is_node(x):
        return x.val
    if isinstance(x, DeviceArray):
        return x
    if isinstance(x, np.ndarray) and x.ndim > 0:
        return D.asarray(x)
    if isinstance(x, (list, tuple)):
        return D.asarray(np.asarray(x))
    return x  # python / numpy scalar


def _needs_grad(x):
    return _is_node(x) and getattr(x, "_rg", False)


def _to_device_grad(g):
    if isinstance(g, DeviceArray):
        return g
    return D.asarray(np.asarray(g))


def check_shape(grad, inp):
    """Sum `grad` down to the shape of operand `inp` (reference :2163-2187)."""
    shape = getattr(inp, "shape", ())
    if shape is None:
        shape = ()
    if not isinstance(grad, DeviceArray):
        return grad
    return D.reduce_to_shape(grad, tuple(shape))


class Function:

    def __init__(self, x=None, y=None):
        self.x = x
        self.y = y
        self.vars = 1
        self.has_conv = 0
        self.repeats = None
        self._rg = _needs_grad(x) or _needs_grad(y)
        self._need = (True, True)
        self.val = self.result()
        self.grad_name = 0
        self.holder = None
        self.graph = None
        self.grad = None
        self.parent = None
        self.son = None
        self.index = None
        self.label = None
        self.x1 = None
        self.x2 = None
        self.abbreviation = None
        v = self.val
        self.size = v.size if hasattr(v, "size") else 1
        self.shape = tuple(v.shape) if hasattr(v, "shape") else ()

    @property
    def T(self):
        return self.val.T if hasattr(self.val, "T") else self.val

    def update(self):
        self.val = self.result()

    def update_label(self):
        pass

    def result(self):
        """require to implement by yourself"""
        pass

    def get_grad(self, grad):
        """require to implement by yourself"""
        pass

    def get_graph(self):
        raise NotImplementedError("symbolic gradient graphs (create_graph) are out of scope")

    def get_label(self):
        pass

    def visualize(self, *a, **k):
        raise NotImplementedError("graph visualisation is out of scope of the B200 hot path")

    def gradient(self, grad=None, create_graph=False):
        if create_graph:
            raise NotImplementedError("create_graph=True is out of scope of the B200 hot path")
        if grad is None:
            np.array(1.).reshape(self.shape)         # ValueError unless size 1, like the reference :90
            grad = D.full(self.shape, 1.0)           # built on the device (no host copy: CUDA-graph capturable)
        Prime(self, grad=grad)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return reshape(self, shape)

    def transpose(self, *axis):
        return transpose(self, axis if len(axis) > 0 else None)

    def get_values(self, x, y):
        return self.get_value(x), self.get_value(y)

    def get_value(self, x):
        if _is_node(x):
            if x.vars == 3:
                self.has_conv += 1
            elif x.has_conv:
                self.has_conv += x.has_conv
        return _value_of(x)

    def __str__(self):
        expression = np.array2string(np.asarray(self.val), prefix=f"{self.__class__.__name__}(", separator=', ')
        return f"{self.__class__.__name__}({expression}, shape={self.shape})"

    def __repr__(self):
        return str(self)

    def __neg__(self):
        return Multi(Matrix(-1, requires_grad=False), self)

    def __add__(self, other):
        return Add(self, other)

    def __radd__(self, other):
        return Add(other, self)

    def __mul__(self, other):
        return Multi(self, other)

    def __rmul__(self, other):
        return Multi(other, self)

    def __sub__(self, other):
        return sub(self, other)

    def __rsub__(self, other):
        return sub(other, self)

    def __truediv__(self, other):
        return Divide(self, other)

    def __rtruediv__(self, other):
        return Divide(other, self)

    def __pow__(self, p, modulo=None):
        return pow(self, p)

    def __matmul__(self, other):
        return Matmul(self, other)

    def __rmatmul__(self, other):
        return Matmul(other, self)

    def __iter__(self):
        for index in range(len(self)):
            yield Matrix(self.val[index], f"Mat{index}")

    def __abs__(self):
        return Abs(self)

    def __getitem__(self, item):
        return Slice(self.val[item], self, item)

    def __len__(self):
        return len(self.val)


class Matrix(Function):
    """Leaf tensor (reference :321-358). `requires_grad=False` is an extension used for internal
    constants (the `-1` of __neg__)."""

    def __init__(self, data, label=None, requires_grad=True):
        if isinstance(data, DeviceArray):
            x = data
        else:
            x = D.asarray(np.asarray(data))
        super().__init__(x)
        self.vars = 0
        self._rg = bool(requires_grad)
        self.label = label
        self.shape = x.shape
        self.size = x.size

    def result(self):
        return self.x

    def update(self):
        self.val = self.x

    def set_x(self, x):
        self.x = x if isinstance(x, DeviceArray) else D.asarray(np.asarray(x))
        self.val = self.x

    def get_label(self):
        return self.label

    def repeat(self, times, axis):
        return repeat(self, times, axis)

    def __len__(self):
        return len(self.x)


# --------------------------------------------------------------------------------------------------
# elementwise binary ops
# --------------------------------------------------------------------------------------------------
class _Binary(Function):
    _op = None

    def __init__(self, x, y):
        super().__init__(x, y)
        self.vars = 2

    def result(self):
        a, b = self.get_values(self.x, self.y)
        if _is_host_scalar(a) and _is_host_scalar(b):
            return {L.B_ADD: np.add, L.B_SUB: np.subtract, L.B_MUL: np.multiply, L.B_DIV: np.divide}[self._op](a, b)
        return D.binary(self._op, a, b)

    def get_grad(self, grad):
        need_a, need_b = self._need
        a, b = _value_of(self.x), _value_of(self.y)
        op = self._op
        if _is_host_scalar(a) or _is_host_scalar(b):
            # one side is a host constant: only the array side can carry a gradient
            ga = gb = None
            if not _is_host_scalar(a) and need_a:
                ga = self._grad_vs_const(grad, a, float(b), const_is_b=True)
            if not _is_host_scalar(b) and need_b:
                gb = self._grad_vs_const(grad, b, float(a), const_is_b=False)
            return ga, gb
        return D.binary_bwd(op, grad, a, b, need_a, need_b, REF_COMPAT)

    def _grad_vs_const(self, grad, arr, c, const_is_b):
        op = self._op
        if op == L.B_ADD:
            g = grad
        elif op == L.B_SUB:
            sign = (-1.0 if const_is_b else 1.0) if REF_COMPAT else (1.0 if const_is_b else -1.0)
            g = grad if sign == 1.0 else D.unary(L.U_SCALE, grad, sign)
        elif op == L.B_MUL:
            g = D.unary(L.U_SCALE, grad, c)
        elif const_is_b:   # arr / c : grad / (c + 1e-8)
            g = D.unary_bwd(L.U_DIVC, grad, param=c)
        else:              # c / arr : grad * (-c / (arr^2 + 1e-8))
            g = D.unary_bwd(L.U_RDIVC, grad, x=arr, param=c)
        return D.reduce_to_shape(g, arr.shape)


class Add(_Binary):
    _op = L.B_ADD


class sub(_Binary):
    _op = L.B_SUB


class Multi(_Binary):
    _op = L.B_MUL


class starmulti(_Binary):
    _op = L.B_MUL


class Divide(_Binary):
    _op = L.B_DIV

    def __init__(self, x, y):
        self._brocasted = False
        super().__init__(x, y)


# --------------------------------------------------------------------------------------------------
# elementwise unary ops
# --------------------------------------------------------------------------------------------------
_HOST_UNARY = {
    L.U_EXP: np.exp, L.U_TANH: np.tanh, L.U_SIGMOID: lambda v: 1. / (1. + np.exp(-v)),
    L.U_RELU: lambda v: np.maximum(v, 0), L.U_LN: np.log, L.U_SQRT: np.sqrt, L.U_SQUARE: np.square,
    L.U_ABS: np.abs, L.U_SIN: np.sin, L.U_COS: np.cos, L.U_TAN: np.tan, L.U_SEC: lambda v: 1. / np.cos(v),
    L.U_SINH: np.sinh, L.U_COSH: np.cosh, L.U_ARCSIN: np.arcsin, L.U_ARCTAN: np.arctan,
    L.U_ARCOT: lambda v: np.arctan(1 / v), L.U_ARCSEC: lambda v: np.arccos(1. / v),
    L.U_COT: lambda v: 1 / np.tan(v), L.U_CSC: lambda v: 1. / np.sin(v), L.U_ARCCOS: np.arccos,
}


class _Unary(Function):
    _op = None
    _param = 0.0

    def __init__(self, x, label_on=False):
        super().__init__(x)

    def _p(self):
        return self._param

    def result(self):
        val = self.get_value(self.x)
        if _is_host_scalar(val):
            return self._host(val)
        return D.unary(self._op, val, self._p())

    def _host(self, val):
        return _HOST_UNARY[self._op](val)

    def get_grad(self, grad):
        x = _value_of(self.x)
        if _is_host_scalar(x):
            return None
        g = D.unary_bwd(self._op, grad, x=x, y=self.val, param=self._p())
        return check_shape(g, self.x)


class exp(_Unary):
    _op = L.U_EXP


class tanh(_Unary):
    _op = L.U_TANH


class Sigmoid(_Unary):
    _op = L.U_SIGMOID


class relu(_Unary):
    """Extension: the reference tape has no relu (only layers.ReLU, layers.py:1280-1290)."""
    _op = L.U_RELU


class ln(_Unary):
    _op = L.U_LN

    def result(self):
        # the reference asserts all(x > 0) on the forward (:1296); checked only for host scalars here —
        # a device-wide check would force a host sync per op (log of a non-positive value yields
        # nan/-inf, which surfaces in the loss).
        val = self.get_value(self.x)
        if _is_host_scalar(val):
            assert np.all(val > 0), "input has element <= zero"
        return super().result()


class log(_Unary):
    """any base log (reference :1314-1341): log(base, x)"""
    _op = L.U_LOGB

    def __init__(self, base, x):
        self.base = base
        assert base != 0, "base can't be zero"
        super().__init__(x)

    def _p(self):
        return float(self.base)

    def _host(self, val):
        assert np.all(val > 0), "input has element <= zero"
        return np.log(val) / np.log(self.base)


class sqrt(_Unary):
    _op = L.U_SQRT


class square(_Unary):
    _op = L.U_SQUARE


class pow(Function):
    def __init__(self, x, power):
        self.power = power
        super().__init__(x, power)

    def result(self):
        val = self.get_value(self.x)
        if _is_host_scalar(val):
            return np.power(val, self.power)
        return D.unary(L.U_POW, val, float(self.power))

    def get_grad(self, grad):
        x = _value_of(self.x)
        if _is_host_scalar(x):
            return None
        return check_shape(D.unary_bwd(L.U_POW, grad, x=x, param=float(self.power)), self.x)


class Abs(_Unary):
    _op = L.U_ABS


class sin(_Unary):
    _op = L.U_SIN


class cos(_Unary):
    _op = L.U_COS


class tan(_Unary):
    _op = L.U_TAN


class sec(_Unary):
    _op = L.U_SEC


class sinh(_Unary):
    _op = L.U_SINH


class cosh(_Unary):
    _op = L.U_COSH


class arcsin(_Unary):
    _op = L.U_ARCSIN


class arctan(_Unary):
    _op = L.U_ARCTAN


class arcot(_Unary):
    _op = L.U_ARCOT


class arcsec(_Unary):
    _op = L.U_ARCSEC


class cot(_Unary):
    """1/tan(x) (reference :1632-1655).  The backward keeps the reference's expression `grad * csc(x)^2` (:1639),
    i.e. without the derivative's minus sign."""
    _op = L.U_COT


class csc(_Unary):
    """1/sin(x) (reference :1472-1495).  As shipped the backward calls the non-existent `np.csc` and raises; the
    expression it spells, -grad * csc(x) * cot(x) (:1476), is what runs here."""
    _op = L.U_CSC


class arcos(_Unary):
    """np.arccos (reference :1434-1456); backward as written there: grad * (-1 / sqrt(x + 1)) (:1438)."""
    _op = L.U_ARCCOS


class trace(Function):
    """np.trace of a square matrix (reference :616-637); backward grad * identity."""

    def __init__(self, x):
        super().__init__(x)
        self.x = x
        self.shape = None

    def result(self):
        val = self.get_value(self.x)
        assert val.ndim == 2, "trace: 2-D input expected"
        n = min(val.shape)
        diag = D.DeviceArray(val.t.diagonal()).contiguous()          # strided view of the diagonal -> (n,)
        return D.reduce_mid(diag, 1, n, 1, 1.0).reshape(())

    def get_grad(self, grad):
        val = self.get_value(self.x)
        assert val.shape[-2] == val.shape[-1], "input has to be square matrix"
        n = val.shape[0]
        self.shape = n
        out = D.zeros((n, n))
        g = D.broadcast_to(_to_device_grad(grad).reshape((1,)), (n,))
        D.copy_into_(g, D.DeviceArray(out.t.diagonal()))
        return check_shape(out, val)


# --------------------------------------------------------------------------------------------------
# reductions / softmax
# --------------------------------------------------------------------------------------------------
class sum(Function):
    def __init__(self, x, axis=None, keepdims=None):
        self.axis = axis
        self.keepdims = keepdims
        super().__init__(x)

    def result(self):
        if self.keepdims is None:
            # NumPy 2.x rejects keepdims=None (what the reference passes by default, :721)
            raise TypeError("sum: pass keepdims=True/False explicitly (keepdims=None is rejected by NumPy 2)")
        return D.sum_axis(self.get_value(self.x), self.axis, bool(self.keepdims))

    def get_grad(self, grad):
        xshape = list(self.x.val.shape)
        if self.axis is None:
            raise TypeError("sum(axis=None) has no backward in the reference (:700-701)")
        ax = self.axis % len(xshape)
        shape = list(xshape)
        shape[ax] = 1
        g = grad.reshape(shape)
        g = D.repeat_axis(g, xshape[ax], ax)   # reshape → repeat along the axis (:702-703)
        return check_shape(g, self.x)


class mean(Function):
    """Full mean of a leaf (reference :980-993)."""

    def __init__(self, x):
        super().__init__(x)
        self.x = x

    def result(self):
        val = self.get_value(self.x)
        return D.reduce_mid(val.contiguous(), 1, val.size, 1, 1.0 / val.size).reshape(())

    def get_grad(self, grad):
        xs = self.x.val
        g = D.unary(L.U_SCALE, grad.reshape((1,)), 1.0 / xs.size)
        return D.broadcast_to(g.reshape((1,) * len(xs.shape)), xs.shape)


class repeat(Function):
    def __init__(self, x, times: int, axis: int):
        self.axis = axis
        self.times = times
        super().__init__(x)

    def result(self):
        return D.repeat_axis(self.get_value(self.x), self.times, self.axis)

    def get_grad(self, grad):
        g = D.sum_axis(grad, self.axis, True)   # np.sum(grad, axis, keepdims=True) (:782)
        return check_shape(g, self.x)


def _softmax_axis(dim, ndim):
    return (dim - 1) % ndim  # the caller passes the rank, the op uses dim-1 (:944)


class softmax(Function):
    def __init__(self, x, dim):
        self.dim = dim - 1
        super().__init__(x)

    def _move(self, a):
        ax = self.dim % a.ndim
        if ax == a.ndim - 1:
            return a, None
        perm = [i for i in range(a.ndim) if i != ax] + [ax]
        return a.transpose(*perm).contiguous(), perm

    def result(self):
        val = self.get_value(self.x)
        moved, perm = self._move(val)
        y = D.softmax_lastaxis(moved)
        if perm is not None:
            y = y.transpose(*np.argsort(perm).tolist()).contiguous()
        return y

    _diag = False

    def get_grad(self, grad):
        y, perm = self._move(self.val)
        g, _ = self._move(grad)
        dx = D.softmax_bwd_lastaxis(g, y, 1.0, self._diag)
        if perm is not None:
            dx = dx.transpose(*np.argsort(perm).tolist()).contiguous()
        return check_shape(dx, self.x)


class SoftmaxWithLoss(softmax):
    """Same forward as softmax; backward is the diagonal-only g*y*(1-y) the reference ships (:919).
    Only its `.val` is used by TimeSoftmaxWithLoss (layers.py:529)."""
    _diag = True


# --------------------------------------------------------------------------------------------------
# contractions
# --------------------------------------------------------------------------------------------------
class Matmul(Function):
    def __init__(self, x, y):
        self._brocased = False
        super().__init__(x, y)
        self.vars = 2

    def result(self):
        a, b = self.get_values(self.x, self.y)
        a, b = D.asarray(a), D.asarray(b)
        if a.ndim > 2:
            self._brocased = True
        return D.matmul(a, b)

    def get_grad(self, grad):
        need_a, need_b = self._need
        a, b = D.asarray(_value_of(self.x)), D.asarray(_value_of(self.y))
        ga = gb = None
        if need_a:   # grad @ y.T   (:532-533; 4-D: transpose(0,1,3,2))
            ga = check_shape(D.matmul(grad, b, trans_b=True), self.x) if b.ndim >= 2 else None
        if need_b:   # x.T @ grad
            gb = check_shape(D.matmul(a, grad, trans_a=True), self.y) if a.ndim >= 2 else None
        return ga, gb


class tensordot(Function):
    """np.tensordot semantics (reference :204-318): permute + flatten both operands, ONE GEMM."""

    def __init__(self, x, y, axes=2):
        self.axes = axes
        super().__init__(x, y)
        self.vars = 2

    def result(self):
        a, b = self.get_values(self.x, self.y)
        a, b = D.asarray(a), D.asarray(b)
        axes = self.axes
        try:
            iter(axes)
        except Exception:
            axes_a, axes_b = list(range(-axes, 0)), list(range(0, axes))
        else:
            axes_a, axes_b = axes
        axes_a = [axes_a] if np.isscalar(axes_a) else list(axes_a)
        axes_b = [axes_b] if np.isscalar(axes_b) else list(axes_b)
        as_, bs = a.shape, b.shape
        equal = len(axes_a) == len(axes_b)
        if equal:
            for k in range(len(axes_a)):
                axes_a[k] %= a.ndim
                axes_b[k] %= b.ndim
                if as_[axes_a[k]] != bs[axes_b[k]]:
                    equal = False
                    break
        if not equal:
            raise ValueError("shape-mismatch for sum")  # reference :271
        notin_a = [k for k in range(a.ndim) if k not in axes_a]
        notin_b = [k for k in range(b.ndim) if k not in axes_b]
        self.newaxes_a = notin_a + axes_a
        self.newaxes_b = axes_b + notin_b
        n2 = D._prod(as_[ax] for ax in axes_a)
        olda = [as_[ax] for ax in notin_a]
        oldb = [bs[ax] for ax in notin_b]
        at = a.transpose(*self.newaxes_a)
        bt = b.transpose(*self.newaxes_b)
        self.newshape_a, self.newshape_b = at.shape, bt.shape   # permuted (un-flattened) shapes
        self.xt = at.reshape(D._prod(olda), n2)
        self.yt = bt.reshape(n2, D._prod(oldb))
        out = D.matmul(self.xt, self.yt)
        self.dot_shape = out.shape
        return out.reshape(olda + oldb)

    def get_grad(self, grad):
        need_a, need_b = self._need
        g = grad.reshape(self.dot_shape)
        ga = gb = None
        if need_a:
            at_grad = D.matmul(g, self.yt, trans_b=True).reshape(self.newshape_a)
            ga = at_grad.transpose(*np.argsort(self.newaxes_a).tolist())
        if need_b:
            bt_grad = D.matmul(self.xt, g, trans_a=True).reshape(self.newshape_b)
            gb = bt_grad.transpose(*np.argsort(self.newaxes_b).tolist())
        return ga, gb


# --------------------------------------------------------------------------------------------------
# layout / indexing / integer ops
# --------------------------------------------------------------------------------------------------
class reshape(Function):
    def __init__(self, data, shape):
        self._target = tuple(shape)
        super().__init__(data)
        self.x = data
        self._shape = data.shape
        self.size = data.size

    def result(self):
        return self.get_value(self.x).reshape(self._target)

    def get_grad(self, grad):
        assert grad.size == self.size
        return grad.reshape(self._shape)


class transpose(Function):
    def __init__(self, x, axis=None):
        self.axis = tuple(axis) if axis is not None else None
        super().__init__(x)

    def result(self):
        val = self.get_value(self.x)
        return val.T if self.axis is None else val.transpose(*self.axis)

    def get_grad(self, grad):
        g = grad.T if self.axis is None else grad.transpose(*np.argsort(self.axis).tolist())
        return check_shape(g, self.x)


class Slice(Function):
    def __init__(self, data, origin_data, index):
        super().__init__(data)
        self.x = origin_data
        self.sliced_data = data
        self._shape = data.shape
        self._index = index
        self._rg = _needs_grad(origin_data)

    def result(self):
        return self.get_value(self.x)

    def get_grad(self, grad=None):
        assert grad.shape == self._shape, f"grad shape {grad.shape} doesn't match {self._shape}"
        zeros = D.zeros(self.x.shape)
        zeros[self._index] = grad
        return zeros


class EqualSlice(Function):
    """np.split(x, num, axis) into `num` tape leaves (vars=3) whose gradients are re-stacked
    (reference :425-486)."""

    def __init__(self, origin_data, num, axis):
        self.num = num
        self.axis = axis
        super().__init__(origin_data)
        self.vars = 4
        self._shape = origin_data.shape
        self.x3 = self.get_real()
        self.grad = None
        self.cumulate = 0

    def result(self):
        val = self.get_value(self.x)
        ax = self.axis % val.ndim
        n = val.shape[ax]
        assert n % self.num == 0, "array split does not result in an equal division"
        self._ax, self._step = ax, n // self.num
        return val

    def get_real(self):
        ls = []
        val = self.val
        for index in range(self.num):
            sl = [slice(None)] * val.ndim
            sl[self._ax] = slice(index * self._step, (index + 1) * self._step)
            a = Matrix(val[tuple(sl)].contiguous())
            a.vars = 3
            a.parent = self
            a.label = f"converge{index}"
            a._rg = self._rg
            ls.append(a)
        return ls

    def get_grad(self, grad=None):
        out = D.zeros(self._shape)
        for index, child in enumerate(self.x3):
            if child.grad is not None:
                sl = [slice(None)] * len(self._shape)
                sl[self._ax] = slice(index * self._step, (index + 1) * self._step)
                out[tuple(sl)] = child.grad
            child.grad = None
        return check_shape(out, self.x)

    def __iter__(self):
        return iter(self.x3)

    def __len__(self):
        return self.num


class argmax(Function):
    """np.argmax over the FLATTENED array, first occurrence, int64 (reference :1803-1818)."""
    _is_min = False

    def __init__(self, x, label_on=False):
        super().__init__(x)
        self._rg = False

    def result(self):
        val = self.get_value(self.x)
        if _is_host_scalar(val):
            return np.int64(0)
        return D.argext_flat(val, self._is_min)

    def gradient(self, grad=None, grad_name="", create_graph=False):
        raise RuntimeError("Doesn't need gradient method")


class argmin(argmax):
    _is_min = True


# --------------------------------------------------------------------------------------------------
# reverse pass
# --------------------------------------------------------------------------------------------------
def _children(node):
    v = node.vars
    if v == 2:
        return (node.x, node.y)
    if v in (1, 4):
        return (node.x,)
    if v == 3:
        return (node.parent,)
    return ()


class Prime:
    """Reverse sweep (replaces reference Prime.prime :1873-1951): iterative post-order DFS to get a
    topological order, then ONE pass with interior accumulation."""

    def __init__(self, func, grad, create_graph=False, grad_name=""):
        if create_graph:
            raise NotImplementedError("create_graph=True is out of scope of the B200 hot path")
        self.grad = grad
        self.prime(func)

    def prime(self, func):
        if not _is_node(func):
            return
        root_grad = _to_device_grad(self.grad)
        # 1. topological order of the sub-graph that needs gradients
        order, seen = [], set()
        stack = [(func, False)]
        while stack:
            node, done = stack.pop()
            if done:
                order.append(node)
                continue
            if id(node) in seen:
                continue
            seen.add(id(node))
            stack.append((node, True))
            for ch in _children(node):
                if _is_node(ch) and ch._rg and id(ch) not in seen:
                    stack.append((ch, False))
        # 2. sweep parents-before-children, accumulating per node
        pending = {id(func): root_grad}
        for node in reversed(order):
            g = pending.pop(id(node), None)
            v = node.vars
            if v == 0 or v == 3:
                if g is not None:
                    node.grad = g if node.grad is None else node.grad + g   # accumulates across calls (:1885)
                if v == 3 and node.parent is not None:
                    node.parent.cumulate += 1
                continue
            if v == 4:
                gx = node.get_grad(None)
                node.cumulate = 0
                self._push(pending, node.x, gx)
                continue
            if g is None:
                continue
            if v == 2:
                node._need = (_needs_grad(node.x), _needs_grad(node.y))
                g1, g2 = node.get_grad(g)
                self._push(pending, node.x, g1)
                self._push(pending, node.y, g2)
            elif v == 1:
                if _needs_grad(node.x):
                    self._push(pending, node.x, node.get_grad(g))

    @staticmethod
    def _push(pending, child, g):
        if g is None or not _needs_grad(child):
            return
        key = id(child)
        prev = pending.get(key)
        pending[key] = g if prev is None else prev + g
