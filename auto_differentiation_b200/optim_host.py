"""Host side of the multi-tensor optimizers: builds the device table once and then issues ONE kernel
per update (reference: a Python loop of three NumPy expressions per parameter,
utils/optimizer.py:25-29 / layers/layers.py:69-73, :829-831, clip_grads :132-141).
"""
import ctypes

import numpy as np
import torch

from . import _lib as L
from . import device as D
from .device import DeviceArray


class TensorTable:
    """Device-resident {param, grad, m, v, numel} table + chunk index lists for the *_multi kernels."""

    def __init__(self, params, grads, ms=None, vs=None):
        n = len(params)
        entries = (L.TensorEntry * max(n, 1))()
        chunk_tensor, chunk_index = [], []
        for i in range(n):
            p, g = params[i], grads[i]
            for name, a in (("param", p), ("grad", g)):
                if not isinstance(a, DeviceArray):
                    raise TypeError("optimizer %ss must be DeviceArrays resident on the GPU (got %s); wrap host "
                                    "arrays with auto_differentiation_b200.asarray first" % (name, type(a).__name__))
                if not a.t.is_contiguous():
                    raise ValueError("optimizer %s %d is a non-contiguous view" % (name, i))
            if p.shape != g.shape:
                raise ValueError("param %d has shape %s but its grad has shape %s" % (i, p.shape, g.shape))
            entries[i].param = p.ptr
            entries[i].grad = g.ptr
            entries[i].m = ms[i].ptr if ms is not None else None
            entries[i].v = vs[i].ptr if vs is not None else None
            entries[i].numel = p.size
            nchunks = (p.size + L.OPT_CHUNK - 1) // L.OPT_CHUNK
            chunk_tensor.extend([i] * nchunks)
            chunk_index.extend(range(nchunks))
        self.n = n
        self.total_chunks = len(chunk_tensor)
        raw = np.frombuffer(bytes(entries), dtype=np.uint8).copy()
        self.table = D.to_device(torch.from_numpy(raw))
        self.chunk_tensor = D.to_device(torch.tensor(chunk_tensor if chunk_tensor else [0], dtype=torch.int32))
        self.chunk_index = D.to_device(torch.tensor(chunk_index if chunk_index else [0], dtype=torch.int32))
        self.key = self.make_key(params, grads)
        self.total_params = int(sum(p.size for p in params))

    @staticmethod
    def make_key(params, grads):
        return tuple((p.ptr, g.ptr, p.size) if isinstance(p, DeviceArray) and isinstance(g, DeviceArray) else None
                     for p, g in zip(params, grads))

    def args(self):
        return (self.table.data_ptr(), self.n, self.total_chunks, self.chunk_tensor.data_ptr(),
                self.chunk_index.data_ptr())


class AdamCore:
    """Adam exactly as the reference writes it (NOT torch-Adam): host scalar
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (1-b1)(g-m); v += (1-b2)(g^2-v); p -= lr_t*m/(sqrt(v)+1e-7)."""

    def __init__(self, lr=0.001, beta1=0.9, beta2=0.999, capturable=False):
        self.lr = lr
        self.beta1 = beta1
        self.beta2 = beta2
        self.iter = 0
        self.m = None
        self.v = None
        self.grad_scale = 1.0  # multiplies g on the fly (1/world_size after a sum all-reduce)
        self._table = None
        # capturable: the step counter and lr_t live on the device (b200ad_adam_step_scalar), so `update` can be
        # captured in a CUDA graph and replayed; `iter` on the host then counts eager calls only.
        self.capturable = capturable
        self._iter_dev = None
        self._lr_dev = None

    def step_count(self):
        """Number of updates applied so far.  With `capturable` the authoritative counter is the device scalar
        (graph replays advance it without touching the host attribute); reading it synchronises."""
        if self.capturable and self._iter_dev is not None:
            return int(self._iter_dev.t.item())
        return int(self.iter)

    def set_step_count(self, n):
        """Restore the step counter (checkpoint load): host attribute and, when it exists, the device scalar."""
        self.iter = int(n)
        if self._iter_dev is not None:
            self._iter_dev.t.fill_(int(n))

    def update(self, params, grads):
        if self.m is None:
            self.m, self.v = [], []
            for param in params:
                self.m.append(D.zeros_like(param))
                self.v.append(D.zeros_like(param))
        if self._table is None or self._table.key != TensorTable.make_key(params, grads):
            self._table = TensorTable(params, grads, self.m, self.v)
        self.iter += 1
        if self.capturable:
            if self._iter_dev is None:
                self._iter_dev = D.zeros((1,), torch.int64)
                self._iter_dev.t.fill_(self.iter - 1)
                self._lr_dev = D.zeros((1,))
            lib = L.load()
            L.check(lib.b200ad_adam_step_scalar(self._iter_dev.ptr, float(self.lr), float(self.beta1), float(self.beta2),
                                                self._lr_dev.ptr, D.stream_ptr()))
            L.check(lib.b200ad_adam_multi(*self._table.args(), 0.0, float(self.beta1), float(self.beta2), 1e-7,
                                          float(self.grad_scale), self._lr_dev.ptr, D.stream_ptr()))
            return
        lr_t = self.lr * np.sqrt(1.0 - self.beta2 ** self.iter) / (1.0 - self.beta1 ** self.iter)
        L.check(L.load().b200ad_adam_multi(*self._table.args(), float(lr_t), float(self.beta1), float(self.beta2),
                                           1e-7, float(self.grad_scale), None, D.stream_ptr()))


class SGDCore:
    def __init__(self, lr=0.01):
        self.lr = lr
        self.grad_scale = 1.0
        self._table = None

    def update(self, params, grads):
        if self._table is None or self._table.key != TensorTable.make_key(params, grads):
            self._table = TensorTable(params, grads)
        L.check(L.load().b200ad_sgd_multi(*self._table.args(), float(self.lr), float(self.grad_scale),
                                          D.stream_ptr()))


_clip_tables = {}


def clip_grads(grads, max_norm):
    """Global-norm clipping decided on the device (no host sync): rate = max/(norm+1e-6), scale if < 1."""
    key = TensorTable.make_key(grads, grads)
    table = _clip_tables.get(key)
    if table is None:
        if len(_clip_tables) > 16:
            _clip_tables.clear()
        table = _clip_tables[key] = TensorTable(grads, grads)
    ws = D.empty((max(table.total_chunks, 1),))
    norm_sq = D.empty((1,))
    lib = L.load()
    L.check(lib.b200ad_sumsq_multi(*table.args(), ws.ptr, norm_sq.ptr, D.stream_ptr()))
    L.check(lib.b200ad_clip_scale_multi(*table.args(), norm_sq.ptr, float(max_norm), D.stream_ptr()))
    return norm_sq


def _same_storage_transposed(pi, pj):
    return (isinstance(pi, DeviceArray) and isinstance(pj, DeviceArray) and pi.ndim == 2 and pj.ndim == 2
            and pi.ptr == pj.ptr and pi.T.shape == pj.shape and pi.t.stride() == pj.t.stride()[::-1])


def unique_params(params):
    """The parameters `remove_duplicate` would keep, WITHOUT touching any gradient buffer (checkpoint save / load
    must not fold tied gradients a second time)."""
    kept = []
    for p in params:
        if not any(p is q or _same_storage_transposed(q, p) for q in kept):
            kept.append(p)
    return kept


def remove_duplicate(params, grads):
    """Merge shared weights (reference layers.py:76-110 / utils/useful_func.py:24-47): identical
    objects, and transposed-view weight tying (`affine_W = embed_W.T`, layers.py:711), detected by
    storage identity instead of the reference's O(P^2) elementwise comparison."""
    params, grads = params[:], grads[:]
    while True:
        find_flg = False
        n = len(params)
        for i in range(0, n - 1):
            for j in range(i + 1, n):
                pi, pj = params[i], params[j]
                if pi is pj:
                    grads[i] += grads[j]
                    find_flg = True
                elif _same_storage_transposed(pi, pj):
                    grads[i] += grads[j].T
                    find_flg = True
                if find_flg:
                    params.pop(j)
                    grads.pop(j)
                    break
            if find_flg:
                break
        if not find_flg:
            break
    return params, grads
