"""Hand-written layers with the surface of the reference's ``layers/layers.py`` (forward/backward pairs
plus ``weights``/``grads`` lists mutated in place), computed on a B200 through libb200ad.so.

Where the reference loops in Python over time steps and launches ~20 CuPy temporaries per step, the
B200 path restructures the math (same results): all input projections, weight gradients and input
gradients of a recurrent layer are single large GEMMs over N*T rows; only the truly sequential part
(h_{t-1} @ Wh, and dh = dA_t @ Wh^T) stays in the loop, each step being one GEMM plus one fused
gate kernel working in place on time-major buffers.

Arrays: layers accept host (NumPy) arrays or DeviceArrays and return DeviceArrays (`np.asarray(out)`
copies to the host); weights handed to a constructor are moved to the device once and then updated in
place by the optimizers (identity is preserved for `remove_duplicate` and Adam's positional m/v).
"""
import math

import numpy as np

from .. import _lib as L
from .. import device as D
from ..device import DeviceArray, asarray
from ..optim_host import AdamCore, SGDCore, clip_grads, remove_duplicate  # noqa: F401
from ..auto_grad import tgrad as ad

np.random.seed(42)  # the reference reseeds the global RNG at import (layers.py:15)


def _dev(x):
    return x if isinstance(x, DeviceArray) else asarray(x)


def softmax(x, dims):
    """Stable softmax (reference layers.py:38-43)."""
    x = _dev(x)
    ax = dims % x.ndim
    if ax != x.ndim - 1:
        perm = [i for i in range(x.ndim) if i != ax] + [ax]
        y = D.softmax_lastaxis(x.transpose(*perm).contiguous())
        return y.transpose(*np.argsort(perm).tolist()).contiguous()
    return D.softmax_lastaxis(x)


class Adam(AdamCore):
    """Adam (reference layers.py:46-73 == utils/optimizer.py:4-29): one fused multi-tensor kernel."""


class SGD(SGDCore):
    """SGD (reference layers.py:821-831)."""


class Repeat:
    """np.repeat of a 2-D offset along rows ('y') or columns ('x'), kept in `.repeated` (reference layers.py:113-121)."""

    def __init__(self, offset_var, repeat_num, axis: str):
        axis = axis.lower()
        if axis == "y":
            self.repeated = D.repeat_axis(_dev(offset_var), repeat_num, 0)
        elif axis == "x":
            self.repeated = D.repeat_axis(_dev(offset_var), repeat_num, 1)
        else:
            raise ValueError("axis should be x or y")


class Sum:
    def forward(self, x):
        return D.sum_axis(_dev(x), 0, True)

    def backward(self, dy, collums):
        return D.repeat_axis(_dev(dy), collums, 0)


class MatMul:
    """Weight-only matmul layer (reference: RNN_Implementation.ipynb:293-315)."""

    def __init__(self, W):
        W = _dev(W)
        self.weights = [W]
        self.X = None
        self.gradients = [D.zeros_like(W)]
        self.grads = self.gradients

    def forward(self, forward_input):
        W, = self.weights
        self.X = _dev(forward_input)
        return D.matmul(self.X, W)

    def backward(self, d_backward_input):
        W = self.weights[0]
        d = _dev(d_backward_input)
        dX = D.matmul(d, W, trans_b=True)
        D.matmul(self.X, d, trans_a=True, out=self.gradients[0])
        return dX


# --------------------------------------------------------------------------------------------------
# affine / feed-forward / relu
# --------------------------------------------------------------------------------------------------
class TimeAffine:
    """(N,T,D) @ W + b (reference layers.py:244-314): one GEMM with the bias fused in the epilogue;
    backward = column-sum + two GEMMs written straight into `grads`."""

    def __init__(self, w, b):
        w, b = _dev(w), _dev(b)
        self.weights = [w, b]
        self.x = None
        self.forward_shape = None
        self.grads = [D.zeros(w.shape), D.zeros(b.shape)]
        self.accumulate = False

    def forward(self, x_sequence):
        x = _dev(x_sequence)
        N, T, Dm = x.shape
        W, b = self.weights
        rx = x.reshape(N * T, Dm)
        out = D.matmul(rx, W, bias=b.reshape(-1))
        self.x = rx
        out = out.reshape(N, T, W.shape[1])
        self.forward_shape = out.shape
        self._xshape = (N, T, Dm)
        return out

    def backward(self, d_sequence):
        N, T, Dm = self._xshape
        W, b = self.weights
        dout = _dev(d_sequence).reshape(N * T, W.shape[1]).contiguous()
        beta = 1.0 if self.accumulate else 0.0
        db = D.reduce_mid(dout, 1, N * T, W.shape[1])                      # np.sum(dout, axis=0)   :306
        D.axpby_(db, self.grads[1].reshape(-1), 1.0, beta)
        D.matmul(self.x, dout, trans_a=True, out=self.grads[0], beta=beta)  # rx.T @ dout           :307
        dx = D.matmul(dout, W, trans_b=True)                                # dout @ W.T            :308
        return dx.reshape(N, T, Dm)


class ReLU:
    def __init__(self):
        self.cache = None

    def forward(self, x):
        x = _dev(x)
        out = D.unary(L.U_RELU, x)
        self.cache = out   # (out > 0) == (x > 0)
        return out

    def backward(self, dout):
        return D.unary_bwd(L.U_RELU, _dev(dout), y=self.cache)


class FeedForward:
    """relu(x @ w + b) (reference layers.py:1242-1266).  The reference's backward stores dW/db into
    `self.weights` (a bug, :1263-1264); this follows the math and fills `self.grads`."""

    def __init__(self, w, b, dtype=np.float32):
        w, b = _dev(w), _dev(b)
        self.weights = [w, b]
        self.grads = [D.zeros(w.shape), D.zeros(b.shape)]
        self.cache = None
        self.accumulate = False

    def forward(self, x):
        w, b = self.weights
        x = _dev(x)
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        out = D.matmul(x2, w, bias=b.reshape(-1), act=L.ACT_RELU)
        self.cache = (x2, out, x.shape)
        return out.reshape(x.shape[:-1] + (w.shape[1],))

    def backward(self, dx):
        w, b = self.weights
        x2, out, xshape = self.cache
        g = _dev(dx).reshape(out.shape).contiguous()
        dz = D.act_bwd(g, out, L.ACT_RELU)
        beta = 1.0 if self.accumulate else 0.0
        D.matmul(x2, dz, trans_a=True, out=self.grads[0], beta=beta)
        db = D.reduce_mid(dz, 1, dz.shape[0], dz.shape[1])
        D.axpby_(db, self.grads[1].reshape(-1), 1.0, beta)
        return D.matmul(dz, w, trans_b=True).reshape(xshape)


class SkipConnection:
    def __init__(self):
        self.weights, self.grads = [], []

    def forward(self, x, y):
        return _dev(x) + _dev(y)

    def backward(self, dout):
        dout = _dev(dout)
        return dout, dout.copy()


# --------------------------------------------------------------------------------------------------
# recurrent cells
# --------------------------------------------------------------------------------------------------
def _cell_call(fn, *args):
    L.check(fn(*args, D.stream_ptr()))


def _last_contig(a):
    """`a` itself when its last axis is contiguous (time slices are then addressable in place)."""
    return a if (a.shape[-1] == 1 or a.t.stride(-1) == 1) else a.contiguous()


def _row_view(a):
    """(ptr, ld) of a 2-D view with contiguous rows."""
    p, ld, _, _ = D._rows_ld(a)
    return p, ld


class RNN:
    """tanh cell h' = tanh(x@Wx + h@Wh + b) (reference layers.py:144-182).  The shipped backward is
    broken (passes the summed db into both MatMul.backward, never fills grads); this
    implements the intended math, which equals what the reference TAPE yields for the same expression."""

    def __init__(self, w_input, w_prev, b):
        w_input, w_prev, b = _dev(w_input), _dev(w_prev), _dev(b)
        self.weights = [w_input, w_prev, b]
        self.grads = [D.zeros(w_input.shape), D.zeros(w_prev.shape), D.zeros(b.shape)]
        self.cache = None

    def forward(self, word_vector, h_prev):
        w_input, w_prev, b = self.weights
        x, h_prev = _dev(word_vector), _dev(h_prev)
        pre = D.matmul(x, w_input, bias=b.reshape(-1))
        D.matmul(h_prev, w_prev, out=pre, beta=1.0)
        h_next = D.bias_act(pre, None, L.ACT_TANH)
        self.cache = (x, h_prev, h_next)
        return h_next

    def backward(self, dh_next):
        w_input, w_prev, b = self.weights
        x, h_prev, h_next = self.cache
        dt = D.act_bwd(_dev(dh_next).contiguous(), h_next, L.ACT_TANH)   # dh*(1-h'^2)       :176
        D.copy_into_(D.reduce_mid(dt, 1, dt.shape[0], dt.shape[1]).reshape(b.shape), self.grads[2])  # db :177
        D.matmul(x, dt, trans_a=True, out=self.grads[0])
        D.matmul(h_prev, dt, trans_a=True, out=self.grads[1])
        dx = D.matmul(dt, w_input, trans_b=True)
        dh_prev = D.matmul(dt, w_prev, trans_b=True)
        return dh_prev, dx


class _TimeRecurrent:
    """Shared time-major machinery of TimeRNN / TimeLSTM."""
    _gate_mult = 1

    def _setup(self, wx, wh, b, stateful):
        wx, wh, b = _dev(wx), _dev(wh), _dev(b)
        self.weights = [wx, wh, b]
        self.grads = [D.zeros(wx.shape), D.zeros(wh.shape), D.zeros(b.shape)]
        self.stateful = stateful
        self.h, self.dh, self.c = None, None, None
        self.accumulate = False
        self.layers = []

    def _project_inputs(self, xs):
        """xs (N,T,D) -> time-major copy (T,N,D) and pre-activations (T,N,G) = x@Wx + b in ONE GEMM."""
        wx, wh, b = self.weights
        N, T, Dm = xs.shape
        G = wx.shape[1]
        xs_tm = xs.transpose(1, 0, 2).contiguous()
        pre = D.matmul(xs_tm.reshape(T * N, Dm), wx, bias=b.reshape(-1)).reshape(T, N, G)
        return xs_tm, pre

    def _weight_grads(self, xs_tm, hprev_tm, dA_tm):
        """dWx, dWh, db and dxs as four large GEMMs / one column-sum over all T*N rows."""
        wx, wh, b = self.weights
        T, N, Dm = xs_tm.shape
        G = dA_tm.shape[2]
        H = wh.shape[0]
        beta = 1.0 if self.accumulate else 0.0
        dA2 = dA_tm.reshape(T * N, G)
        D.matmul(xs_tm.reshape(T * N, Dm), dA2, trans_a=True, out=self.grads[0], beta=beta)
        D.matmul(hprev_tm.reshape(T * N, H), dA2, trans_a=True, out=self.grads[1], beta=beta)
        db = D.reduce_mid(dA2, 1, T * N, G)
        D.axpby_(db, self.grads[2].reshape(-1), 1.0, beta)
        dxs_tm = D.matmul(dA2, wx, trans_b=True).reshape(T, N, Dm)
        return dxs_tm.transpose(1, 0, 2)

    def set_state(self, h, c=None):
        self.h = None if h is None else _dev(h)
        self.c = None if c is None else _dev(c)

    def reset_state(self):
        self.h, self.c = None, None


class TimeRNN(_TimeRecurrent):
    """Unrolled tanh RNN (reference layers.py:185-225; forward semantics kept, backward as intended)."""

    def __init__(self, w_input, w_prev, b, stateful=False):
        self._setup(w_input, w_prev, b, stateful)

    def forward(self, x_sequence):
        wx, wh, b = self.weights
        xs = _dev(x_sequence)
        N, T, Dm = xs.shape
        H = wh.shape[0]
        if not self.stateful or self.h is None:
            self.h = D.zeros((N, H))
        hs_ext = D.empty((T + 1, N, H))       # hs_ext[t] = h_{t-1}; hs_ext[0] = initial state
        D.copy_into_(self.h, hs_ext[0])
        xs_tm = xs.transpose(1, 0, 2).contiguous()
        if D.rnn_seq_supported(N, H):
            # x@Wx + b for all steps in ONE GEMM, then the whole time loop in ONE persistent launch (csrc/rnn_seq.cu)
            pre = D.matmul(xs_tm.reshape(T * N, Dm), wx, bias=b.reshape(-1)).reshape(T, N, H)
            D.rnn_tanh_seq_fwd_(hs_ext, pre, wh)
        else:
            # the same GEMM written straight into hs_ext[1:]; each step then completes its slice in place:
            # hs_ext[t+1] = tanh(hs_ext[t+1] + h_{t-1} @ Wh) — one launch per step (fused beta / tanh epilogue)
            D.matmul(xs_tm.reshape(T * N, Dm), wx, bias=b.reshape(-1), out=hs_ext[1:].reshape(T * N, H))
            for t in range(T):
                D.matmul(hs_ext[t], wh, out=hs_ext[t + 1], beta=1.0, act=L.ACT_TANH)
        self.h = hs_ext[T]
        self._cache = (xs_tm, hs_ext)
        return hs_ext[1:].transpose(1, 0, 2)

    def backward(self, dhs):
        wx, wh, b = self.weights
        xs_tm, hs_ext = self._cache
        T, N, Dm = xs_tm.shape
        H = wh.shape[0]
        dhs = _last_contig(_dev(dhs))
        if D.rnn_seq_supported(N, H):
            dA, dh = D.rnn_tanh_seq_bwd(dhs, hs_ext, wh)     # all T + 1 BPTT phases in ONE persistent launch
        else:
            dA = D.empty((T, N, H))
            dh = None
            lib = L.load()
            for t in reversed(range(T)):
                pa, lda = _row_view(dhs[:, t, :])
                pb, ldb = _row_view(dh) if dh is not None else (None, 0)
                ph, ldh = _row_view(hs_ext[t + 1])
                pd, ldd = _row_view(dA[t])
                _cell_call(lib.b200ad_rnn_cell_bwd, pa, lda, pb, ldb, ph, ldh, pd, ldd, N, H)
                dh = D.matmul(dA[t], wh, trans_b=True)
        self.dh = dh
        return self._weight_grads(xs_tm, hs_ext[:T], dA)


class LSTM:
    """One LSTM step (reference layers.py:616-676): gates [f,g,i,o]; returns (c', h')."""

    def __init__(self, wx, wh, b, dtype=np.float32):
        wx, wh, b = _dev(wx), _dev(wh), _dev(b)
        self.weights = [wx, wh, b]
        self.grads = [D.zeros(wx.shape), D.zeros(wh.shape), D.zeros(b.shape)]
        self._var = None

    def forward(self, x, h, c):
        wx, wh, b = self.weights
        x, h, c = _dev(x), _dev(h).contiguous(), _dev(c).contiguous()
        N, H = h.shape
        a = D.matmul(x, wx, bias=b.reshape(-1))          # x@wx + b
        D.matmul(h, wh, out=a, beta=1.0)                 # + h@wh              :632
        gates, ct, ht = D.empty((N, 4 * H)), D.empty((N, H)), D.empty((N, H))
        _cell_call(L.load().b200ad_lstm_cell_fwd, a.ptr, 4 * H, c.ptr, H, gates.ptr, 4 * H, ct.ptr, H, ht.ptr, H,
                   N, H)
        self._var = (ct, ht, h, c, x, gates)
        return ct, ht

    def backward(self, dh_next, dc_next):
        wx, wh, b = self.weights
        ct, ht, h, c, x, gates = self._var
        N, H = h.shape
        dh_next, dc_next = _dev(dh_next).contiguous(), _dev(dc_next).contiguous()
        db = D.empty((N, 4 * H))                         # np.hstack((stack0..3))   :663
        dc = D.empty((N, H))
        _cell_call(L.load().b200ad_lstm_cell_bwd, dh_next.ptr, H, None, 0, dc_next.ptr, gates.ptr, 4 * H, c.ptr, H,
                   ct.ptr, H, ht.ptr, H, db.ptr, 4 * H, dc.ptr, N, H)
        D.matmul(x, db, trans_a=True, out=self.grads[0])  # dWx = x.T @ db            :666
        D.matmul(h, db, trans_a=True, out=self.grads[1])  # dWh = h.T @ db            :667
        dx = D.matmul(db, wx, trans_b=True)               # :668
        dh = D.matmul(db, wh, trans_b=True)               # :669
        D.copy_into_(D.reduce_mid(db, 1, N, 4 * H).reshape(b.shape), self.grads[2])  # :674
        return dx, dh, dc


class TimeLSTM(_TimeRecurrent):
    """Unrolled LSTM (reference layers.py:552-604)."""

    def __init__(self, wx, wh, b, stateful=False):
        self._setup(wx, wh, b, stateful)
        self.forward_shape = None

    def forward(self, xs):
        wx, wh, b = self.weights
        xs = _dev(xs)
        N, T, Dm = xs.shape
        H = wh.shape[0]
        if not self.stateful or self.h is None:
            self.h = D.zeros((N, H))
        if not self.stateful or self.c is None:
            self.c = D.zeros((N, H))
        xs_tm, pre = self._project_inputs(xs)          # pre: (T,N,4H)
        hs_ext = D.empty((T + 1, N, H))
        cs_ext = D.empty((T + 1, N, H))
        gates = D.empty((T, N, 4 * H))
        D.copy_into_(self.h, hs_ext[0])
        D.copy_into_(self.c, cs_ext[0])
        if D.lstm_seq_supported(N, H):
            # all T dependent steps from one C call (csrc/lstm_seq.cu): two launches per step, no host-language loop
            D.lstm_seq_fwd_(hs_ext, cs_ext, gates, pre, wh)
        else:
            lib = L.load()
            for t in range(T):
                D.matmul(hs_ext[t], wh, out=pre[t], beta=1.0)   # A_t = x_t@wx + b + h_{t-1}@wh     :632
                _cell_call(lib.b200ad_lstm_cell_fwd, pre[t].ptr, 4 * H, cs_ext[t].ptr, H, gates[t].ptr, 4 * H,
                           cs_ext[t + 1].ptr, H, hs_ext[t + 1].ptr, H, N, H)
        self.h, self.c = hs_ext[T], cs_ext[T]
        self._cache = (xs_tm, hs_ext, cs_ext, gates)
        out = hs_ext[1:].transpose(1, 0, 2)
        self.forward_shape = out.shape
        return out

    def backward(self, dhs):
        wx, wh, b = self.weights
        xs_tm, hs_ext, cs_ext, gates = self._cache
        T, N, Dm = xs_tm.shape
        H = wh.shape[0]
        dhs = _last_contig(_dev(dhs))
        if D.lstm_seq_supported(N, H):
            dA, self.dh = D.lstm_seq_bwd(dhs, hs_ext, cs_ext, gates, wh)
            return self._weight_grads(xs_tm, hs_ext[:T], dA)
        dA = D.empty((T, N, 4 * H))
        dc_bufs = (D.empty((N, H)), D.empty((N, H)))
        dh, dc = None, None
        lib = L.load()
        for t in reversed(range(T)):
            pa, lda = _row_view(dhs[:, t, :])
            pb, ldb = (dh.ptr, H) if dh is not None else (None, 0)
            dc_out = dc_bufs[t & 1]
            _cell_call(lib.b200ad_lstm_cell_bwd, pa, lda, pb, ldb, dc.ptr if dc is not None else None,
                       gates[t].ptr, 4 * H, cs_ext[t].ptr, H, cs_ext[t + 1].ptr, H, hs_ext[t + 1].ptr, H,
                       dA[t].ptr, 4 * H, dc_out.ptr, N, H)
            dc = dc_out
            dh = D.matmul(dA[t], wh, trans_b=True)          # dh = db @ wh.T                     :669
        self.dh = dh
        return self._weight_grads(xs_tm, hs_ext[:T], dA)


# --------------------------------------------------------------------------------------------------
# embedding / loss
# --------------------------------------------------------------------------------------------------
class Embedding:
    def __init__(self, W):
        W = _dev(W)
        self.params = [W]
        self.weights = self.params
        self.grads = [D.zeros(W.shape)]
        self.idx = None

    def forward(self, idx):
        W, = self.params
        self.idx = D.as_index(idx)
        return D.embedding_gather(W, self.idx)

    def backward(self, dout):
        dW, = self.grads
        D.embedding_scatter_add_(_dev(dout), self.idx, dW)   # dW[...] = 0; np.add.at(dW, idx, dout)
        return None


class TimeEmbedding:
    """(N,T) ids -> (N,T,D) (reference layers.py:461-507): one gather / one deterministic scatter-add
    over all N*T positions instead of a Python loop over T."""

    def __init__(self, w):
        w = _dev(w)
        self.weights = [w]
        self.grads = [D.zeros(w.shape)]
        self.layers = []
        self.idx = None

    def forward(self, xs):
        self.idx = D.as_index(xs)
        return D.embedding_gather(self.weights[0], self.idx)

    def backward(self, d_sequence):
        D.embedding_scatter_add_(_dev(d_sequence), self.idx, self.grads[0])
        return None


class TimeDropout:
    """Inverted dropout over a whole (N,T,H) block (reference layers.py:678-697): train mode draws
    `u`, keeps where u > ratio and scales by 1/(1-ratio); eval mode is the identity.  The draws come
    from the device generator; `reference_rng=True` takes them from the host `np.random.rand` stream
    in the reference's order, so that a run seeded like the reference reproduces its masks exactly."""

    def __init__(self, dropout_ratio=0.5, reference_rng=False):
        self.weights, self.grads = [], []
        self.dropout_ratio = dropout_ratio
        self.mask = None
        self.train_flg = True
        self.reference_rng = reference_rng

    def forward(self, xs):
        xs = _dev(xs)
        if not self.train_flg:
            self.mask = None
            return xs
        scale = 1 / (1.0 - self.dropout_ratio)
        if self.reference_rng:
            keep = np.random.rand(*xs.shape) > self.dropout_ratio      # fp64 compare as in the reference :687
            out, self.mask = D.dropout_fwd(xs, asarray(keep.astype(np.float32)), 0.5, scale)
        else:
            out, self.mask = D.dropout_fwd(xs, D.uniform(xs.shape), self.dropout_ratio, scale)
        return out

    def backward(self, dout):
        if self.mask is None:                 # eval-mode forward: identity (the reference would fail on `dout * None`)
            return dout
        return D.binary(L.B_MUL, _dev(dout).contiguous(), self.mask)


class Rnnlm:
    """Two stacked stateful LSTMs between dropout layers with the output projection tied to the
    embedding, `affine_W = embed_W.T` (reference layers.py:699-753).  The tied matrices are two views
    of one buffer; `remove_duplicate` folds their gradients before the update.  Needs
    wordvec_size == hidden_size, like the reference (second LSTM and the tied affine both consume H)."""

    def __init__(self, vocab_size=10000, wordvec_size=100, hidden_size=100, weights=None, reference_rng=False):
        V, Dm, H = vocab_size, wordvec_size, hidden_size
        if weights is None:
            rn = np.random.randn
            embed_W = (rn(V, Dm) / 100).astype("f")
            lstm_Wx = (rn(Dm, 4 * H) / np.sqrt(Dm)).astype("f")
            lstm_Wh = (rn(H, 4 * H) / np.sqrt(H)).astype("f")
            lstm_Wx2 = (rn(Dm, 4 * H) / np.sqrt(Dm)).astype("f")
            lstm_Wh2 = (rn(H, 4 * H) / np.sqrt(H)).astype("f")
            lstm_b = np.zeros(4 * H).astype("f")
            lstm_b2 = np.zeros(4 * H).astype("f")
            affine_b = np.zeros(V).astype("f")
        else:
            # parameter order of `params` (:731-734); the tied (Dm,V) view is rebuilt from the embedding
            embed_W, lstm_Wx, lstm_Wh, lstm_b, lstm_Wx2, lstm_Wh2, lstm_b2 = weights[:7]
            affine_b = weights[-1]
        embed_W = _dev(embed_W)
        affine_W = embed_W.T
        self.layers = [
            TimeEmbedding(embed_W),
            TimeDropout(reference_rng=reference_rng),
            TimeLSTM(lstm_Wx, lstm_Wh, lstm_b, stateful=True),
            TimeDropout(reference_rng=reference_rng),
            TimeLSTM(lstm_Wx2, lstm_Wh2, lstm_b2, stateful=True),
            TimeDropout(reference_rng=reference_rng),
            TimeAffine(affine_W, affine_b),
        ]
        self.loss_layer = TimeSoftmaxWithLoss()
        self.lstm_layer = [self.layers[2], self.layers[4]]
        self.drop_layers = [self.layers[1], self.layers[3], self.layers[5]]
        self.params, self.grads = [], []
        for layer in self.layers:
            self.params += layer.weights
            self.grads += layer.grads
        self.weights = self.params

    def train_mode(self, flag=True):
        for layer in self.drop_layers:
            layer.train_flg = flag

    def predict(self, xs):
        for layer in self.layers:
            xs = layer.forward(xs)
        return xs

    def forward(self, xs, ts):
        score = self.predict(xs)
        return self.loss_layer.forward(score, ts)

    def backward(self, dout=1):
        dout = self.loss_layer.backward(dout)
        for layer in reversed(self.layers):
            dout = layer.backward(dout)
        return dout

    def reset_state(self):
        for layer in self.lstm_layer:
            layer.reset_state()


class RnnlmTrainer:
    """Truncated-BPTT trainer (reference layers.py:756-818): `batch_size` read heads spaced
    len(x)//batch_size apart walk the corpus `time_size` tokens per iteration; each iteration is
    forward, backward, remove_duplicate, optional clip_grads, optimizer.update; perplexity =
    exp(mean loss since the last evaluation)."""

    def __init__(self, model, optimizer):
        self.model = model
        self.optimizer = optimizer
        self.time_idx = None
        self.ppl_list = None
        self.eval_interval = None
        self.current_epoch = 0

    def get_batch(self, x, t, batch_size, time_size):
        x, t = np.asarray(x), np.asarray(t)
        data_size = len(x)
        jump = data_size // batch_size
        # position of sample i at step k is (i*jump + time_idx + k) % data_size        :773-776
        pos = (np.arange(batch_size)[:, None] * jump + self.time_idx + np.arange(time_size)[None, :]) % data_size
        self.time_idx += time_size
        return x[pos].astype("i"), t[pos].astype("i")

    def fit(self, xs, ts, max_epoch=10, batch_size=20, time_size=35, max_grad=None, eval_interval=20):
        data_size = len(xs)
        max_iters = data_size // (batch_size * time_size)
        self.time_idx = 0
        self.ppl_list = []
        self.eval_interval = eval_interval
        model, optimizer = self.model, self.optimizer
        total_loss, loss_count = 0, 0
        for _ in range(max_epoch):
            for iters in range(max_iters):
                batch_x, batch_t = self.get_batch(xs, ts, batch_size, time_size)
                loss = model.forward(batch_x, batch_t)
                model.backward()
                params, grads = remove_duplicate(model.params, model.grads)
                if max_grad is not None:
                    clip_grads(grads, max_grad)
                optimizer.update(params, grads)
                total_loss = loss + total_loss            # stays on the device until an evaluation point
                loss_count += 1
                if (eval_interval is not None) and not iters % eval_interval:
                    self.ppl_list.append(float(np.exp(float(total_loss) / loss_count)))
                    total_loss, loss_count = 0, 0
            self.current_epoch += 1
        return self.ppl_list

    def save(self, prefix="weights"):
        params, _ = remove_duplicate(self.model.params, self.model.grads)
        for index, p in enumerate(params):
            np.save("%s%d.npy" % (prefix, index), np.asarray(p))


class TimeSoftmaxWithLoss:
    """Masked mean cross-entropy over (N,T,V) logits (reference layers.py:510-549).  `forward`
    returns the loss as a 0-d DeviceArray (float(loss) syncs; arithmetic on it stays on the device)."""

    def __init__(self):
        self.weights, self.grads = [], []
        self.cache = None
        self.ignore_label = -1

    def forward(self, xs, ts):
        xs = _dev(xs)
        N, T, V = xs.shape
        if getattr(ts, "ndim", 0) == 3:            # one-hot labels -> argmax(axis=2)   :519-520
            ts = D.argmax_rows(_dev(ts).reshape(N * T, V))
        else:
            ts = D.as_index(ts).reshape(N * T)
        lc, ys = D.xent_fwd(xs.reshape(N * T, V), ts, self.ignore_label, 0.0, 0)
        self.cache = (ts, ys, lc, (N, T, V))
        return lc[0]

    def backward(self, dout=1):
        ts, ys, lc, (N, T, V) = self.cache
        dx = D.xent_bwd(ys, ts, lc, float(dout), self.ignore_label, out=ys)   # in place on the cached probs :541
        return dx.reshape(N, T, V)


# --------------------------------------------------------------------------------------------------
# multi-head attention
# --------------------------------------------------------------------------------------------------
class MultiHeadAttention:
    """Reference forward layers.py:1331-1375 (backward there is `pass`); the backward implemented here
    yields the gradients of the tape twin Test/MultiHeadAttentionTest.py:49-91.

    softmax(QK^T/sqrt(d_k)) is evaluated in the stable form — equal to the reference's naive
    exp/sum (:1356) whenever that is finite.  Head split/merge are strided views consumed directly
    by the batched GEMMs (no permute copies)."""

    def __init__(self, wordvec_size, hidden_size, num_heads, weights=None):
        self.num_heads = num_heads
        self.d_k = wordvec_size // num_heads
        assert isinstance(self.d_k, int), f"{self.d_k} is not int"
        if weights is None:
            weights = [np.random.randn(wordvec_size, hidden_size) for _ in range(4)]
        self.weights = [_dev(w) for w in weights]
        self.grads = [D.zeros(w.shape) for w in self.weights]
        self.cache = None
        self.accumulate = False

    def forward(self, query, key, value):
        shared = query is key and key is value      # self-attention: one upload, one operand split for Q / K / V
        query = _dev(query)
        key, value = (query, query) if shared else (_dev(key), _dev(value))
        B, S, Dm = query.shape
        h = self.num_heads
        wq, wk, wv, wo = self.weights
        Hd = wq.shape[1]
        dk = Hd // h
        q2, k2, v2 = (a.reshape(B * S, Dm) for a in (query, key, value))
        if shared:
            Q, K, V = D.matmul_shared_a(q2, [wq, wk, wv])                        # :1347-1349
        else:
            Q, K, V = D.matmul(q2, wq), D.matmul(k2, wk), D.matmul(v2, wv)      # :1347-1349
        Qh, Kh, Vh = (a.reshape(B, S, h, dk).transpose(0, 2, 1, 3) for a in (Q, K, V))
        scale = 1.0 / math.sqrt(dk)
        if D.attn_supported(Qh, Kh, Vh):
            # fused tcgen05 attention pass (:1355-1357): the S x S scores stay on the SM; only the row log-sum-exp
            # is kept for the backward
            Oh, lse = D.attn_fwd(Qh, Kh, Vh, scale)
            O2 = Oh.transpose(0, 2, 1, 3).reshape(B * S, h * dk)
            P = None
        else:
            scores = D.matmul(Qh, Kh, trans_b=True)                              # :1355 (scale folded into softmax)
            P = D.softmax_lastaxis(scores, in_scale=scale)                       # :1356
            O = D.empty((B, S, h, dk))
            D.matmul(P, Vh, out=O.transpose(0, 2, 1, 3))                         # :1357 written in merged layout
            O2 = O.reshape(B * S, h * dk)
            lse = None
        out = D.matmul(O2, wo).reshape(B, S, wo.shape[1])                        # :1360
        self.cache = (q2, k2, v2, Qh, Kh, Vh, P, O2, (B, S, Dm, h, dk), lse)
        return out

    def backward(self, dout=None, notify=None):
        """`notify(i)` (optional) is called right after the last kernel writing gradient bucket i was
        enqueued (1 = wo, 2 = wv, 3 = wk, 4 = wq) so a data-parallel reducer can start that all-reduce
        while the rest of the backward still runs.  With `self.input_grads = False` the three GEMMs
        producing d(query/key/value) are skipped (the reference tape always computes them)."""
        if dout is None:
            return None   # the reference's backward is a stub (layers.py:1366-1367)
        wq, wk, wv, wo = self.weights
        q2, k2, v2, Qh, Kh, Vh, P, O2, (B, S, Dm, h, dk), lse = self.cache
        beta = 1.0 if self.accumulate else 0.0
        d2 = _dev(dout).reshape(B * S, wo.shape[1]).contiguous()
        D.matmul(O2, d2, trans_a=True, out=self.grads[3], beta=beta)             # dwo
        if notify:
            notify(1)
        dO = D.matmul(d2, wo, trans_b=True).reshape(B, S, h, dk)
        dOh = dO.transpose(0, 2, 1, 3)
        if P is None:
            Oh = O2.reshape(B, S, h, dk).transpose(0, 2, 1, 3)
            dQh, dKh, dVh = D.attn_bwd(Qh, Kh, Vh, Oh, dOh, lse, 1.0 / math.sqrt(dk))
            dQ, dK, dV = (g.transpose(0, 2, 1, 3) for g in (dQh, dKh, dVh))      # (B,S,h,dk), contiguous
            dV2 = dV.reshape(B * S, h * dk)
            D.matmul(v2, dV2, trans_a=True, out=self.grads[2], beta=beta)        # dwv
            if notify:
                notify(2)
        else:
            dP = D.matmul(dOh, Vh, trans_b=True)                                 # (B,h,S,S)
            dV = D.empty((B, S, h, dk))
            D.matmul(P, dOh, trans_a=True, out=dV.transpose(0, 2, 1, 3))
            dV2 = dV.reshape(B * S, h * dk)
            D.matmul(v2, dV2, trans_a=True, out=self.grads[2], beta=beta)        # dwv
            if notify:
                notify(2)
            dS = D.softmax_bwd_lastaxis(dP, P, out_scale=1.0 / math.sqrt(dk), out=dP)
            dQ, dK = D.empty((B, S, h, dk)), D.empty((B, S, h, dk))
            D.matmul(dS, Kh, out=dQ.transpose(0, 2, 1, 3))
            D.matmul(dS, Qh, trans_a=True, out=dK.transpose(0, 2, 1, 3))
        dQ2, dK2 = dQ.reshape(B * S, h * dk), dK.reshape(B * S, h * dk)
        D.matmul(k2, dK2, trans_a=True, out=self.grads[1], beta=beta)            # dwk
        if notify:
            notify(3)
        D.matmul(q2, dQ2, trans_a=True, out=self.grads[0], beta=beta)            # dwq
        if notify:
            notify(4)
        if not getattr(self, "input_grads", True):
            return None
        dquery = D.matmul(dQ2, wq, trans_b=True).reshape(B, S, Dm)
        dkey = D.matmul(dK2, wk, trans_b=True).reshape(B, S, Dm)
        dvalue = D.matmul(dV2, wv, trans_b=True).reshape(B, S, Dm)
        return dquery, dkey, dvalue

    def split_heads(self, x):
        x = _dev(x)
        batch_size, sequence_length, _ = x.shape
        return x.reshape(batch_size, sequence_length, self.num_heads, self.d_k).transpose(0, 2, 1, 3)

    def combine_heads(self, x):
        x = _dev(x)
        batch_size, _, sequence_length, _ = x.shape
        return x.transpose(0, 2, 1, 3).reshape(batch_size, sequence_length, -1)


# --------------------------------------------------------------------------------------------------
# Bahdanau-style dot attention over encoder states
# --------------------------------------------------------------------------------------------------
class WeightSum:
    """c[n] = sum_t a[n,t] * hs[n,t,:]  (reference layers.py:964-989) as one batched (1 x T) @ (T x H) product."""

    def __init__(self):
        self.weights, self.grads = [], []
        self.cache = None

    def forward(self, hs, a):
        hs, a = _dev(hs), _dev(a)
        N, T, H = hs.shape
        c = D.matmul(a.reshape(N, 1, T), hs).reshape(N, H)
        self.cache = (hs, a)
        return c

    def backward(self, dout):
        hs, a = self.cache
        N, T, H = hs.shape
        d = _dev(dout).reshape(N, 1, H)
        da = D.matmul(d, hs, trans_b=True).reshape(N, T)            # sum(dout*hs, axis=2)         :984
        dhs = D.matmul(a.reshape(N, T, 1), d)                       # repeat(dout) * ar            :985
        return dhs, da


class AttentionWeight:
    """weight = softmax_t(hs[n,t,:] . h[n,:])  (reference layers.py:992-1030)."""

    def __init__(self):
        self.weights, self.grads = [], []
        self.cache = None

    def forward(self, hs, h):
        hs, h = _dev(hs), _dev(h)
        N, T, H = hs.shape
        c = D.matmul(hs, h.reshape(N, H, 1)).reshape(N, T)
        weight = D.softmax_lastaxis(c)
        self.cache = (hs, h, weight)
        return weight

    def backward(self, dout):
        hs, h, weight = self.cache
        N, T, H = hs.shape
        ds = D.softmax_bwd_lastaxis(_dev(dout).reshape(N, T), weight)   # common_term                  :1018-1019
        dh = D.matmul(ds.reshape(N, 1, T), hs).reshape(N, H)            # sum(common*hs, axis=1)       :1023
        dhs = D.matmul(ds.reshape(N, T, 1), h.reshape(N, 1, H))         # common_term_expanded * h     :1026
        return dhs, dh


class Attention:
    """AttentionWeight + WeightSum for one decoder step (reference layers.py:1033-1052)."""

    def __init__(self):
        self.weights, self.grads = [], []
        self.layers = [AttentionWeight(), WeightSum()]
        self.hs_weight = None

    def forward(self, hs, h):
        attention_weight, weight_sum = self.layers
        a = attention_weight.forward(hs, h)
        c = weight_sum.forward(hs, a)
        self.hs_weight = a
        return c

    def backward(self, dout):
        attention_weight, weight_sum = self.layers
        hs0, da = weight_sum.backward(dout)
        hs1, dh = attention_weight.backward(da)
        return hs0 + hs1, dh


class TimeAttention:
    """All decoder steps at once (reference layers.py:1055-1083 loops over time in Python): scores =
    hs_dec @ hs_enc^T, softmax over the encoder axis, context = a @ hs_enc — three batched GEMMs forward, four
    backward (the two encoder-gradient terms accumulate in the GEMM epilogue)."""

    def __init__(self):
        self.weights, self.grads = [], []
        self.layers = []
        self.hs_weight = None
        self.cache = None

    def forward(self, hs_enc, hs_dec):
        hs_enc, hs_dec = _dev(hs_enc), _dev(hs_dec)
        scores = D.matmul(hs_dec, hs_enc, trans_b=True)          # (N, Td, Te)
        a = D.softmax_lastaxis(scores)
        out = D.matmul(a, hs_enc)                                # (N, Td, H)
        self.hs_weight = a
        self.cache = (hs_enc, hs_dec, a)
        return out

    def backward(self, dout):
        hs_enc, hs_dec, a = self.cache
        dout = _dev(dout)
        da = D.matmul(dout, hs_enc, trans_b=True)                # (N, Td, Te)
        dhs_enc = D.matmul(a, dout, trans_a=True)                # a^T @ dout            (WeightSum term)
        ds = D.softmax_bwd_lastaxis(da, a, out=da)
        dhs_dec = D.matmul(ds, hs_enc)                           # (N, Td, H)
        D.matmul(ds, hs_dec, trans_a=True, out=dhs_enc, beta=1.0)   # + ds^T @ hs_dec     (AttentionWeight term)
        return dhs_enc, dhs_dec


# --------------------------------------------------------------------------------------------------
# AttentionSeq2seq: encoder / attention decoder / loss composition and its fit loop
# --------------------------------------------------------------------------------------------------
def _concat_last(a, b):
    """np.concatenate((a, b), axis=2) for (N, T, X) device arrays (two strided copies into one buffer)."""
    N, T, Xa = a.shape
    Xb = b.shape[2]
    out = D.empty((N, T, Xa + Xb))
    D.copy_into_(a, out[:, :, :Xa])
    D.copy_into_(b, out[:, :, Xa:])
    return out


class Encoder:
    """TimeEmbedding -> TimeLSTM (reference layers.py:834-863); returns the last hidden state.  `weights`
    (optional, [embed, wx, wh, b]) replaces the reference's np.random.randn initialisation (:838-841)."""

    def __init__(self, vocab_size, wordvec_size=100, hidden_size=100, weights=None):
        if isinstance(vocab_size, list):
            vocab_size = vocab_size[0]
        if weights is None:
            weights = [np.random.randn(vocab_size, wordvec_size) / 100,
                       np.random.randn(wordvec_size, 4 * hidden_size) / np.sqrt(wordvec_size),
                       np.random.randn(hidden_size, 4 * hidden_size) / np.sqrt(hidden_size),
                       np.zeros(4 * hidden_size)]
        w1, w2, w3, b = weights
        self.layers = [TimeEmbedding(w1), TimeLSTM(w2, w3, b, stateful=False)]
        self.weights = self.layers[0].weights + self.layers[1].weights
        self.grads = self.layers[0].grads + self.layers[1].grads
        self.hs = None

    def forward(self, xs):
        wordembed, lstm = self.layers
        hs = lstm.forward(wordembed.forward(xs))
        self.hs = hs
        return hs[:, -1, :]

    def backward(self, dh):
        dhs = D.zeros(self.hs.shape)
        D.copy_into_(_dev(dh), dhs[:, -1, :])
        wordembed, lstm = self.layers
        wordembed.backward(lstm.backward(dhs))


class Decoder:
    """"Peeky" decoder (reference layers.py:866-930): the encoder's last hidden state h is concatenated to every LSTM
    input and to every affine input; TimeEmbedding -> concat(h, e_t) -> TimeLSTM (stateful, started from h) ->
    concat(h, s_t) -> TimeAffine."""

    def __init__(self, vocab_size=12, wordvec_size=100, hidden_size=100, weights=None):
        V, Dm, H = vocab_size, wordvec_size, hidden_size
        if weights is None:
            rn = np.random.randn
            weights = [rn(V, Dm) / 100, rn(Dm + H, 4 * H) / np.sqrt(Dm), rn(H, 4 * H) / np.sqrt(H), np.zeros(4 * H),
                       rn(H * 2, V) / np.sqrt(H), np.zeros(V)]
        w1, w2, w3, b, w4, b1 = weights
        self.layers = [TimeEmbedding(w1), TimeLSTM(w2, w3, b, stateful=True), TimeAffine(w4, b1)]
        self.loss_layer = [TimeSoftmaxWithLoss()]
        self.weights, self.grads = [], []
        self.hidden_size = hidden_size
        for layer in self.layers:
            self.weights += layer.weights
            self.grads += layer.grads

    def _peek(self, h, T):
        """np.repeat(h, T, axis=0).reshape(N, T, H) (:893): h broadcast along time, materialised once."""
        N, H = h.shape
        return D.broadcast_to(h.reshape((N, 1, H)), (N, T, H))

    def forward(self, xs, hidden_state):
        h = _dev(hidden_state).contiguous()
        embed, lstm, affine = self.layers
        lstm.set_state(h)                                                        # :889
        out = embed.forward(xs)
        N, T = out.shape[:2]
        hidden = self._peek(h, T)
        out = lstm.forward(_concat_last(hidden, out))                             # :894-895
        return affine.forward(_concat_last(hidden, out))                          # :896-897

    def backward(self, dout):
        embed, lstm, affine = self.layers
        H = self.hidden_size
        dout = affine.backward(dout)
        dhs0, dout = dout[:, :, :H], dout[:, :, H:]                                # :903
        dout = lstm.backward(dout.contiguous())
        dhs1, dout = dout[:, :, :H], dout[:, :, H:]                                # :905
        embed.backward(dout)
        dhs = D.binary(L.B_ADD, dhs0.contiguous(), dhs1.contiguous())
        return D.binary(L.B_ADD, lstm.dh, D.sum_axis(dhs, 1, False))              # lstm.dh + sum_t dhs   :908

    def generate(self, h, xs, size):
        h = _dev(h).contiguous()
        embed, lstm, affine = self.layers
        lstm.set_state(h)
        N, T = np.asarray(xs).shape
        hidden = h.reshape((N, T, self.hidden_size))                              # :918 (N = T = 1)
        sampled, ids = [], np.asarray(xs)
        for _ in range(size):
            out = lstm.forward(_concat_last(hidden, embed.forward(ids)))
            score = affine.forward(_concat_last(hidden, out))
            sample_id = int(D.argext_flat(score.reshape(-1)))                     # np.argmax(score, axis=2) with N = T = 1
            ids = np.array([[sample_id]])
            sampled.append(sample_id)
        return sampled


class Seq2Seq:
    """Encoder + peeky Decoder (reference layers.py:933-961)."""

    def __init__(self, word_id, vocab_size, wordvec_size=100, hidden_size=100, weights=None):
        ew = dw = None
        if weights is not None:
            ew, dw = list(weights[:4]), list(weights[4:])
        self.encoder = Encoder(vocab_size, wordvec_size, hidden_size, weights=ew)
        self.decoder = Decoder(vocab_size, wordvec_size, hidden_size, weights=dw)
        self.loss_layer = TimeSoftmaxWithLoss()
        self.word_id = word_id
        self.weights = self.encoder.weights + self.decoder.weights
        self.grads = self.encoder.grads + self.decoder.grads

    def forward(self, xs, ts):
        ts = np.asarray(ts)
        decoder_input, decoder_target = ts[:, :-1], ts[:, 1:]
        h = self.encoder.forward(np.asarray(xs))
        score = self.decoder.forward(decoder_input, h)
        return self.loss_layer.forward(score, decoder_target)

    def backward(self, dout=1):
        dout = self.loss_layer.backward(dout)
        dh = self.decoder.backward(dout)
        self.encoder.backward(dh)

    def generate(self, question, word_id, size):
        h = self.encoder.forward(np.asarray(question))
        return self.decoder.generate(h, np.array(word_id["_"]).reshape(1, 1), size)

    def reset_state(self):
        self.decoder.layers[1].reset_state()


class AttentionEncoder(Encoder):
    """Returns every hidden state (reference layers.py:1159-1168)."""

    def forward(self, xs):
        return self.layers[1].forward(self.layers[0].forward(xs))

    def backward(self, dhs):
        dout = self.layers[1].backward(dhs)
        return self.layers[0].backward(dout)


class AttentionDecoder:
    """TimeEmbedding -> TimeLSTM (stateful, started from the encoder's last state) -> TimeAttention ->
    concat(context, hidden) -> TimeAffine (reference layers.py:1086-1156)."""

    def __init__(self, vocab_size, wordvec_size, hidden_size, weights=None):
        if isinstance(vocab_size, list):
            vocab_size = vocab_size[0]
        V, Dm, H = vocab_size, wordvec_size, hidden_size
        if weights is None:
            rn = np.random.randn
            weights = [(rn(V, Dm) / 100).astype("f"), (rn(Dm, 4 * H) / np.sqrt(Dm)).astype("f"),
                       (rn(H, 4 * H) / np.sqrt(H)).astype("f"), np.zeros(4 * H).astype("f"),
                       (rn(2 * H, V) / np.sqrt(2 * H)).astype("f"), np.zeros(V).astype("f")]
        embed_W, lstm_Wx, lstm_Wh, lstm_b, affine_W, affine_b = weights
        self.embed = TimeEmbedding(embed_W)
        self.lstm = TimeLSTM(lstm_Wx, lstm_Wh, lstm_b, stateful=True)
        self.attention = TimeAttention()
        self.affine = TimeAffine(affine_W, affine_b)
        self.weights, self.grads = [], []
        for layer in (self.embed, self.lstm, self.attention, self.affine):
            self.weights += layer.weights
            self.grads += layer.grads

    def forward(self, xs, enc_hs):
        enc_hs = _dev(enc_hs)
        self.lstm.set_state(enc_hs[:, -1].contiguous())                           # :1112-1113
        dec_hs = self.lstm.forward(self.embed.forward(xs))
        c = self.attention.forward(enc_hs, dec_hs)
        return self.affine.forward(_concat_last(c, dec_hs))                       # :1118-1119

    def backward(self, dscore):
        dout = self.affine.backward(dscore)
        N, T, H2 = dout.shape
        H = H2 // 2
        dc, ddec_hs0 = dout[:, :, :H].contiguous(), dout[:, :, H:]
        denc_hs, ddec_hs1 = self.attention.backward(dc)
        ddec_hs = D.binary(L.B_ADD, ddec_hs0.contiguous(), ddec_hs1)
        dout = self.lstm.backward(ddec_hs)
        last = denc_hs[:, -1]                                                     # denc_hs[:, -1] += dh   :1133
        D.copy_into_(D.binary(L.B_ADD, last.contiguous(), self.lstm.dh), last)
        self.embed.backward(dout)
        return denc_hs

    def generate(self, enc_hs, start_id, sample_size):
        enc_hs = _dev(enc_hs)
        sampled, sample_id = [], start_id
        self.lstm.set_state(enc_hs[:, -1].contiguous())
        for _ in range(sample_size):
            x = np.array([sample_id]).reshape((1, 1))
            dec_hs = self.lstm.forward(self.embed.forward(x))
            c = self.attention.forward(enc_hs, dec_hs)
            score = self.affine.forward(_concat_last(c, dec_hs))
            sample_id = int(D.argext_flat(score.reshape(-1)))                     # np.argmax(score.flatten())  :1153
            sampled.append(sample_id)
        return sampled


class AttentionSeq2seq:
    """Encoder-decoder with dot attention (reference layers.py:1171-1240; Seq2Seq.forward/backward :942-953).
    `fit(train_questions, train_answer, optimizer=..., max_epoch=..., batch_size=..., max_grad=...)` keeps the
    reference's loop: forward, backward, global-norm clip (its local clip_grad :1192-1199 — note: no epsilon, unlike
    layers.clip_grads), optimizer.update.  The clip is one multi-tensor sum-of-squares + one scale launch."""

    def __init__(self, word_id, vocab_size, wordvec_size, hidden_size, weights=None):
        ew = dw = None
        if weights is not None:
            ew, dw = list(weights[:4]), list(weights[4:])
        self.encoder = AttentionEncoder(vocab_size, wordvec_size, hidden_size, weights=ew)
        self.decoder = AttentionDecoder(vocab_size, wordvec_size, hidden_size, weights=dw)
        self.loss_layer = TimeSoftmaxWithLoss()
        self.weights = self.encoder.weights + self.decoder.weights
        self.grads = self.encoder.grads + self.decoder.grads
        self.word_id = word_id
        self.vocab_size = vocab_size
        self.wordvec_size = wordvec_size
        self.hidden_size = hidden_size

    def forward(self, xs, ts):
        ts = np.asarray(ts)
        decoder_input, decoder_target = ts[:, :-1], ts[:, 1:]
        enc_hs = self.encoder.forward(np.asarray(xs))
        score = self.decoder.forward(decoder_input, enc_hs)
        return self.loss_layer.forward(score, decoder_target)

    def backward(self, dout=1):
        dscore = self.loss_layer.backward(dout)
        denc_hs = self.decoder.backward(dscore)
        self.encoder.backward(denc_hs)

    def generate(self, question, word_id, size):
        enc_hs = self.encoder.forward(np.asarray(question))
        return self.decoder.generate(enc_hs, word_id["_"] if isinstance(word_id, dict) else word_id, size)

    def reset_state(self):
        self.decoder.lstm.reset_state()

    @staticmethod
    def clip_grad(grads, max_grad):
        """fit's local clip (:1192-1199): scale every gradient by max_grad / norm when norm > max_grad."""
        total = None
        for g in grads:
            s = D.sum_axis(D.unary(L.U_SQUARE, g.reshape(-1)))
            total = s if total is None else D.binary(L.B_ADD, total, s)
        norm = math.sqrt(float(total))
        if norm > max_grad:
            for g in grads:
                D.axpby_(g, g, 0.0, max_grad / norm)      # g *= max_grad / norm, in place
        return norm

    def fit(self, train_questions, train_answer, **fit_params):
        optimizer = fit_params["optimizer"]
        max_epoch = fit_params["max_epoch"]
        batch_size = fit_params["batch_size"]
        max_grad = fit_params.get("max_grad")
        max_iter = len(train_questions) // batch_size
        history = []
        for _ in range(max_epoch):
            for iters in range(max_iter):
                bq = train_questions[iters * batch_size:(iters + 1) * batch_size]
                ba = train_answer[iters * batch_size:(iters + 1) * batch_size]
                loss = self.forward(bq, ba)
                self.backward()
                if max_grad is not None:
                    self.clip_grad(self.grads, max_grad)
                optimizer.update(self.weights, self.grads)
                history.append(loss)
        return [float(v) for v in history]

    def get_params(self, deep=False):
        return {"word_id": self.word_id, "vocab_size": self.vocab_size, "wordvec_size": self.wordvec_size,
                "hidden_size": self.hidden_size}

    def set_params(self, **params):
        for key, value in params.items():
            setattr(self, key, value)
        return self
