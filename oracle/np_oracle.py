"""TEST INFRASTRUCTURE — CPU (NumPy) restatement of the reference's algorithm for the autodiff hot path.

This is the parity ORACLE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
(``auto_differentiation_b200``) never does.  It travels to the GPU box (``/root/reference`` does not).

Pinning: the reference ships no golden vectors and its one test has no assertions,
so the pin is the reference code itself: ``oracle/make_golden.py`` runs the UNMODIFIED reference
(imported in place by ``oracle/ref_loader.py``) on seeded inputs and commits inputs+outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below against those
fixtures (and against the live reference when /root/reference is present).

Every function cites the reference lines it restates (paths relative to /root/reference).
All arithmetic is NumPy in the dtype of the inputs (the reference's default is float64).
"""
import numpy as np

# ======================================================================================================
# check_shape — auto_grad/tgrad.py:2163-2187 (the heuristic, quirks included)
# ======================================================================================================
def check_shape(grad, shape):
    grad = np.asarray(grad)
    shape = tuple(shape)
    gshape = grad.shape
    offset = len(gshape) - len(shape)
    if offset == 0 and gshape == shape:
        return grad
    if offset < 0:
        raise RuntimeError(f"grad shape {gshape} smaller than {shape}")          # :2179
    if offset == 0 and gshape != shape:
        repeats = [max(s1, s2) // min(s1, s2) for s1, s2 in zip(shape, gshape)]   # :2181
        grad = np.sum(grad, axis=int(np.argmax(repeats)), keepdims=True)          # :2182
    for _ in range(offset):
        grad = np.sum(grad, axis=0)                                               # :2183-2184
    size = int(np.prod(shape)) if shape else 1
    if shape and np.size(grad) != size and not np.size(grad) % size:
        grad = np.sum(grad, axis=0, keepdims=True)                                # :2185-2186
    return grad


def reduce_to_shape(grad, shape):
    """The general (correct) un-broadcast; equals check_shape wherever that is right (SURVEY B-3)."""
    grad = np.asarray(grad)
    shape = tuple(shape)
    if grad.shape == shape:
        return grad
    if grad.ndim < len(shape):
        raise RuntimeError(f"grad shape {grad.shape} smaller than {shape}")
    pad = (1,) * (grad.ndim - len(shape)) + shape
    axes = tuple(i for i, (g, t) in enumerate(zip(grad.shape, pad)) if t == 1 and g != 1)
    return np.sum(grad, axis=axes, keepdims=True).reshape(shape)


# ======================================================================================================
# the tape — Function tgrad.py:14-201, Matrix :321-358, op classes, Prime.prime :1873-1951
# ======================================================================================================
REF_COMPAT = True


class Node:
    """One tape node: eager forward in the constructor (tgrad.py:22), operand back-pointers."""
    __slots__ = ("op", "ins", "val", "par", "grad", "arity", "saved")

    def __init__(self, op, ins, par):
        self.op, self.ins, self.par = op, ins, par
        self.grad = None
        self.saved = None
        self.arity = 0 if op == "leaf" else len(ins)
        self.val = par["data"] if op == "leaf" else FWD[op](self, *[value(i) for i in ins])

    @property
    def shape(self):
        return np.shape(self.val)


def value(x):
    return x.val if isinstance(x, Node) else x                 # get_value :125-137


def leaf(data):
    return Node("leaf", (), {"data": data if isinstance(data, np.ndarray) else np.array(data)})  # :323-325


def op(name, *ins, **par):
    return Node(name, ins, par)


def _shape_of(x):
    return x.shape if isinstance(x, Node) else np.shape(x)


def _cs(g, operand):
    return check_shape(g, _shape_of(operand))


def _stable_softmax(val, axis):                                  # :964-970 / :926-932 / layers.py:38-43
    x = val - val.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / np.sum(e, axis=axis, keepdims=True)


def _tensordot_fwd(n, a, b):                                     # :230-306 (NumPy's tensordot restated)
    axes = n.par.get("axes", 2)
    try:
        iter(axes)
    except Exception:
        axes_a, axes_b = list(range(-axes, 0)), list(range(0, axes))
    else:
        axes_a, axes_b = axes
    axes_a = [axes_a] if np.isscalar(axes_a) else list(axes_a)
    axes_b = [axes_b] if np.isscalar(axes_b) else list(axes_b)
    if len(axes_a) != len(axes_b):
        raise ValueError("shape-mismatch for sum")
    for k in range(len(axes_a)):
        if a.shape[axes_a[k]] != b.shape[axes_b[k]]:
            raise ValueError("shape-mismatch for sum")           # :271
        axes_a[k] %= a.ndim
        axes_b[k] %= b.ndim
    notin_a = [k for k in range(a.ndim) if k not in axes_a]
    notin_b = [k for k in range(b.ndim) if k not in axes_b]
    newaxes_a, newaxes_b = notin_a + axes_a, axes_b + notin_b
    n2 = int(np.prod([a.shape[ax] for ax in axes_a]))
    olda, oldb = [a.shape[ax] for ax in notin_a], [b.shape[ax] for ax in notin_b]
    at, bt = a.transpose(newaxes_a), b.transpose(newaxes_b)      # :294-295
    a2 = at.reshape(int(np.prod(olda)), n2)                      # :300
    b2 = bt.reshape(n2, int(np.prod(oldb)))                      # :301
    out = a2 @ b2                                                # :304
    n.saved = (a2, b2, at.shape, bt.shape, newaxes_a, newaxes_b, out.shape)
    return out.reshape(olda + oldb)


def _tensordot_bwd(n, g):                                        # :222-228
    a2, b2, ashape, bshape, newaxes_a, newaxes_b, dot_shape = n.saved
    g = g.reshape(dot_shape)
    ga = (g @ b2.T).reshape(ashape)
    gb = (a2.T @ g).reshape(bshape)
    # the reference re-applies the forward permutation (correct for involutions only); identical to the
    # inverse permutation for every permutation that occurs on the hot path
    return ga.transpose(np.argsort(newaxes_a)), gb.transpose(np.argsort(newaxes_b))


def _matmul_bwd(n, g):                                           # :530-533
    x, y = value(n.ins[0]), value(n.ins[1])
    if np.ndim(x) > 2:
        return g @ np.swapaxes(y, -1, -2), np.swapaxes(x, -1, -2) @ g   # 4-D: transpose(0,1,3,2) :532
    return _cs(g @ y.T, n.ins[0]), _cs(x.T @ g, n.ins[1])


def _sum_bwd(n, g):                                              # :699-704
    x = value(n.ins[0])
    shape = list(x.shape)
    shape[n.par["axis"]] = 1
    g = np.reshape(g, shape).repeat(x.shape[n.par["axis"]], axis=n.par["axis"])
    return (_cs(g, n.ins[0]),)


def _softmax_bwd(n, g):                                          # :947-951
    dx = n.val * g
    dx = dx - n.val * np.sum(dx, axis=n.par["dim"] - 1, keepdims=True)
    return (_cs(dx, n.ins[0]),)


def _slice_bwd(n, g):                                            # :370-378
    z = np.zeros(_shape_of(n.ins[0]))
    z[n.par["index"]] = g
    return (z,)


FWD = {
    "add": lambda n, a, b: np.add(a, b),                                         # :682-684
    "sub": lambda n, a, b: np.subtract(a, b),                                    # :769-771
    "mul": lambda n, a, b: np.multiply(a, b),                                    # :878-880
    "div": lambda n, a, b: a / b,                                                # :1026-1028
    "exp": lambda n, a: np.exp(a),                                               # :1073-1075
    "tanh": lambda n, a: np.tanh(a),                                             # :1613-1615
    "sigmoid": lambda n, a: 1. / (1. + np.exp(-a)),                              # :901-903
    "relu": lambda n, a: np.maximum(a, 0),                                       # layers.py:1284-1286
    "ln": lambda n, a: np.log(a),                                                # :1294-1297
    "log": lambda n, a: np.log(a) / np.log(n.par["base"]),                       # :1338-1341
    "sqrt": lambda n, a: np.sqrt(a),                                             # :1784-1786
    "square": lambda n, a: np.square(a),                                         # :1741-1743
    "pow": lambda n, a: np.power(a, n.par["power"]),                             # :1103-1105
    "abs": lambda n, a: np.abs(a),                                               # :1846-1848
    "sin": lambda n, a: np.sin(a), "cos": lambda n, a: np.cos(a), "tan": lambda n, a: np.tan(a),
    "sec": lambda n, a: 1. / np.cos(a), "sinh": lambda n, a: np.sinh(a), "cosh": lambda n, a: np.cosh(a),
    "arcsin": lambda n, a: np.arcsin(a), "arctan": lambda n, a: np.arctan(a),
    "arcot": lambda n, a: np.arctan(1 / a), "arcsec": lambda n, a: np.arccos(1. / a),
    "cot": lambda n, a: 1 / np.tan(a), "csc": lambda n, a: 1. / np.sin(a), "arcos": lambda n, a: np.arccos(a),   # :1654, :1493, :1455
    "trace": lambda n, a: np.trace(a),                                           # :635-637
    "sum": lambda n, a: np.sum(a, axis=n.par["axis"], keepdims=n.par["keepdims"]),   # :720-721
    "mean": lambda n, a: np.mean(a),                                             # :991-993
    "repeat": lambda n, a: np.repeat(a, n.par["times"], axis=n.par["axis"]),     # :795-797
    "softmax": lambda n, a: _stable_softmax(a, n.par["dim"] - 1),                # :964-970 (dim-1!)
    "softmax_with_loss": lambda n, a: _stable_softmax(a, n.par["dim"] - 1),      # :926-932
    "matmul": lambda n, a, b: np.matmul(a, b),                                   # :535-539
    "tensordot": _tensordot_fwd,
    "reshape": lambda n, a: np.reshape(a, n.par["shape"]),                       # :519-520
    "transpose": lambda n, a: a.T if n.par.get("axes") is None else a.transpose(n.par["axes"]),  # :611-613
    "slice": lambda n, a: a[n.par["index"]],                                     # :193-195, :392-393
}

# operand gradients; each returns a tuple aligned with node.ins
BWD = {
    "add": lambda n, g: (_cs(g.copy(), n.ins[0]), _cs(g.copy(), n.ins[1])),      # :672-674
    "sub": lambda n, g: ((_cs(-1 * g, n.ins[0]), _cs(g, n.ins[1])) if REF_COMPAT  # :753-756 (sign-inverted)
                         else (_cs(g, n.ins[0]), _cs(-1 * g, n.ins[1]))),
    "mul": lambda n, g: (_cs(g * value(n.ins[1]), n.ins[0]), _cs(g * value(n.ins[0]), n.ins[1])),   # :850-852
    "div": lambda n, g: (_cs(g / (value(n.ins[1]) + 1e-8), n.ins[0]),            # :1004
                         _cs(g * (-value(n.ins[0]) / (np.square(value(n.ins[1])) + 1e-8)), n.ins[1])),  # :1005
    "exp": lambda n, g: (_cs(g * np.exp(value(n.ins[0])), n.ins[0]),),           # :1058-1060
    "tanh": lambda n, g: (_cs(g * (1.0 - np.square(n.val)), n.ins[0]),),         # :1596-1598
    "sigmoid": lambda n, g: (g * n.val * (1. - n.val),),                         # :887-888
    "relu": lambda n, g: (g * (value(n.ins[0]) > 0),),                           # layers.py:1288-1290
    "ln": lambda n, g: (_cs(g * (1 / value(n.ins[0])), n.ins[0]),),              # :1280-1282
    "log": lambda n, g: (_cs(g * (1 / (value(n.ins[0]) * np.log(n.par["base"]))), n.ins[0]),),   # :1324-1326
    "sqrt": lambda n, g: (_cs(g * 0.5 * np.power(value(n.ins[0]), -0.5), n.ins[0]),),   # :1769-1772
    "square": lambda n, g: (_cs(g * 2 * value(n.ins[0]), n.ins[0]),),            # :1725-1728
    "pow": lambda n, g: (_cs(g * n.par["power"] * np.power(value(n.ins[0]), n.par["power"] - 1), n.ins[0]),),
    "abs": lambda n, g: (_cs(g * np.where(value(n.ins[0]) != n.val, -1., 1.), n.ins[0]),),   # :1850-1853
    "sin": lambda n, g: (_cs(g * np.cos(value(n.ins[0])), n.ins[0]),),           # :1120-1122
    "cos": lambda n, g: (_cs(-g * np.sin(value(n.ins[0])), n.ins[0]),),          # :1363-1365
    "tan": lambda n, g: (_cs(g * np.square(1. / np.cos(value(n.ins[0]))), n.ins[0]),),   # :1517-1519
    "sec": lambda n, g: (_cs(g * np.tan(value(n.ins[0])) * (1. / np.cos(value(n.ins[0]))), n.ins[0]),),   # :1151
    "sinh": lambda n, g: (_cs(g * np.cosh(value(n.ins[0])), n.ins[0]),),         # :1184-1186
    "cosh": lambda n, g: (_cs(g * np.sinh(value(n.ins[0])), n.ins[0]),),         # :1401-1403
    "arcsin": lambda n, g: (_cs(g * (1. / np.sqrt(1. - np.square(value(n.ins[0])))), n.ins[0]),),   # :1215
    "arctan": lambda n, g: (_cs(g * (1 / (1 + np.square(value(n.ins[0])))), n.ins[0]),),   # :1558
    "arcot": lambda n, g: (_cs(g * (-1 / (1 + np.square(value(n.ins[0])))), n.ins[0]),),   # :1682
    "arcsec": lambda n, g: (_cs(g * (1. / (np.abs(value(n.ins[0])) * np.sqrt(np.square(value(n.ins[0])) - 1.))),
                                n.ins[0]),),                                     # :1246
    "cot": lambda n, g: (_cs(g * np.square(1. / np.sin(value(n.ins[0]))), n.ins[0]),),       # :1639-1641 (sign as written)
    "csc": lambda n, g: (_cs(-g * (1. / np.sin(value(n.ins[0]))) * (1. / np.tan(value(n.ins[0]))), n.ins[0]),),   # :1476 (intent; shipped code raises)
    "arcos": lambda n, g: (_cs(g * (-1. / np.sqrt(value(n.ins[0]) + 1.)), n.ins[0]),),       # :1438-1440 (as written)
    "trace": lambda n, g: (_cs(g * np.identity(np.shape(value(n.ins[0]))[0]), n.ins[0]),),   # :623-628
    "sum": _sum_bwd,
    "mean": lambda n, g: (np.full(np.shape(value(n.ins[0])), g / np.size(value(n.ins[0]))),),   # :986-989
    "repeat": lambda n, g: (_cs(np.sum(g, axis=n.par["axis"], keepdims=True), n.ins[0]),),      # :781-783
    "softmax": _softmax_bwd,
    "softmax_with_loss": lambda n, g: (g * n.val * (1 - n.val),),                # :918-919 (diagonal only)
    "matmul": _matmul_bwd,
    "tensordot": _tensordot_bwd,
    "reshape": lambda n, g: (np.reshape(g, _shape_of(n.ins[0])),),               # :498-500
    "transpose": lambda n, g: (_cs(g.T if n.par.get("axes") is None
                                   else g.transpose(np.argsort(n.par["axes"])), n.ins[0]),),   # :596-598
    "slice": _slice_bwd,
}


def backward(root, grad=None):
    """Prime.prime tgrad.py:1873-1951: LIFO DFS, y-branch first, NO interior accumulation (a node
    reachable by k paths is differentiated k times); leaves accumulate `.grad` (:1885)."""
    if grad is None:
        grad = np.array(1.).reshape(np.shape(root.val))                           # :89-90
    stack = [(root, grad)]
    while stack:
        node, g = stack.pop()
        if not isinstance(node, Node):                                            # :1879-1880
            continue
        if node.op == "leaf":
            node.grad = g if node.grad is None else node.grad + g                 # :1885
            continue
        gs = BWD[node.op](node, g)
        for operand, og in zip(node.ins, gs):                                     # x pushed first, y popped first
            stack.append((operand, og))


def argmax_flat(x):
    return np.int64(np.argmax(x))                                                 # :1810-1812


def argmin_flat(x):
    return np.int64(np.argmin(x))                                                 # :1827-1829


# ======================================================================================================
# layers — layers/layers.py
# ======================================================================================================
def time_affine_forward(x, W, b):                                                 # :273-285
    N, T, D = x.shape
    out = np.dot(x.reshape(N * T, -1), W) + b
    return out.reshape(N, T, -1)


def time_affine_backward(x, W, dout):                                             # :298-314
    N, T, D = x.shape
    d2 = dout.reshape(N * T, -1)
    rx = x.reshape(N * T, -1)
    db = np.sum(d2, axis=0)
    dW = np.dot(rx.T, d2)
    dx = np.dot(d2, W.T).reshape(*x.shape)
    return dx, dW, db


def relu_forward(x):
    return np.maximum(x, 0)                                                       # :1284-1286


def relu_backward(x, dout):
    return dout * (x > 0)                                                         # :1288-1290


def feed_forward_forward(x, w, b):                                                # :1249-1256
    return np.maximum(x @ w + b, 0)


def feed_forward_backward(x, w, z_out, dout):                                     # :1258-1266 (math)
    dz = dout * (z_out > 0)
    return dz @ w.T, x.T @ dz, np.sum(dz, axis=0, keepdims=True)


def _sigmoid(x):
    return 1 / (1 + np.exp(-x))                                                   # :629-630


def lstm_forward(x, h, c, wx, wh, b):                                             # LSTM.forward :624-646
    first = x @ wx + h @ wh + b
    s1, s2, s3, s4 = np.split(first, 4, axis=1)
    f, g, i, o = _sigmoid(s1), np.tanh(s2), _sigmoid(s3), _sigmoid(s4)
    ct = f * c + g * i
    ht = o * np.tanh(ct)
    return ct, ht, (ct, ht, h, c, x, f, g, i, o)


def lstm_backward(cache, wx, wh, dh_next, dc_next):                               # LSTM.backward :648-676
    ct, ht, h, c, x, f, g, i, o = cache
    tang_c_2 = np.tanh(ct) ** 2
    dh_next_o = dh_next * o
    dh_next_o_i = dh_next_o * i
    dc_next_i = dc_next * i
    dh_next_o_f = dh_next_o * f
    tang_c_2_1 = 1 - tang_c_2
    stack0 = (dh_next_o_f * tang_c_2_1 + dc_next * f) * c * (1 - f)
    stack1 = (dh_next_o_i * (1 - tang_c_2) + dc_next_i) * (1 - g ** 2)
    stack2 = (dh_next_o_i * tang_c_2_1 + dc_next_i) * g * (1 - i)
    stack3 = dh_next * ht * (1 - o)
    db = np.hstack((stack0, stack1, stack2, stack3))
    dWx, dWh = x.T @ db, h.T @ db
    dx, dh = db @ wx.T, db @ wh.T
    dc = dh_next_o_f * (1 - tang_c_2) + dc_next * f
    return dx, dh, dc, dWx, dWh, np.sum(db, axis=0)


def time_lstm_forward(xs, wx, wh, b, h0=None, c0=None):                           # TimeLSTM.forward :561-579
    N, T, D = xs.shape
    H = wh.shape[0]
    h = np.zeros((N, H), dtype=xs.dtype) if h0 is None else h0
    c = np.zeros((N, H), dtype=xs.dtype) if c0 is None else c0
    hs = np.empty((N, T, H), dtype=xs.dtype)
    caches = []
    for t in range(T):
        c, h, cache = lstm_forward(xs[:, t, :], h, c, wx, wh, b)
        hs[:, t, :] = h
        caches.append(cache)
    return hs, caches, h, c


def time_lstm_backward(caches, wx, wh, dhs):                                      # TimeLSTM.backward :581-597
    N, T, H = dhs.shape
    dxs = np.empty((N, T, wx.shape[0]), dtype=dhs.dtype)
    dh, dc = 0, np.zeros((N, H), dtype=dhs.dtype)
    grads = [0, 0, 0]
    for t in reversed(range(T)):
        dx, dh, dc, dWx, dWh, db = lstm_backward(caches[t], wx, wh, dhs[:, t, :] + dh, dc)
        dxs[:, t, :] = dx
        grads[0] = grads[0] + dWx
        grads[1] = grads[1] + dWh
        grads[2] = grads[2] + db
    return dxs, grads, dh


def rnn_forward(x, h_prev, wx, wh, b):                                            # RNN.forward :156-173
    return np.tanh(x @ wx + h_prev @ wh + b)


def rnn_backward(x, h_prev, h_next, wx, wh, dh_next):
    """Intended math of RNN.backward :175-182 (the shipped body is broken); equals what the
    reference TAPE yields for tanh(x@Wx + h@Wh + b) — checked in tests/test_oracle_golden.py."""
    dt = dh_next * (1 - np.square(h_next))                                        # :176
    db = np.sum(dt, axis=0)                                                       # :177
    return dt @ wh.T, dt @ wx.T, x.T @ dt, h_prev.T @ dt, db                      # dh_prev, dx, dWx, dWh, db


def time_rnn_forward(xs, wx, wh, b, h0=None):                                     # TimeRNN.forward :193-208
    N, T, D = xs.shape
    H = wh.shape[0]
    h = np.zeros((N, H), dtype=xs.dtype) if h0 is None else h0
    hs = np.empty((N, T, H), dtype=xs.dtype)
    hprevs = []
    for t in range(T):
        hprevs.append(h)
        h = rnn_forward(xs[:, t, :], h, wx, wh, b)
        hs[:, t, :] = h
    return hs, hprevs


def time_rnn_backward(xs, hs, hprevs, wx, wh, dhs):                               # TimeRNN.backward :210-225 (intended)
    N, T, H = dhs.shape
    dxs = np.empty((N, T, wx.shape[0]), dtype=dhs.dtype)
    dh = 0
    gWx, gWh, gb = 0, 0, 0
    for t in reversed(range(T)):
        dh, dx, dWx, dWh, db = rnn_backward(xs[:, t, :], hprevs[t], hs[:, t, :], wx, wh, dhs[:, t, :] + dh)
        dxs[:, t, :] = dx
        gWx, gWh, gb = gWx + dWx, gWh + dWh, gb + db
    return dxs, (gWx, gWh, gb), dh


def embedding_forward(W, idx):
    return W[idx]                                                                 # :448-452


def embedding_backward(W_shape, idx, dout, dtype=np.float64):                     # :454-458
    dW = np.zeros(W_shape, dtype=dtype)
    np.add.at(dW, idx, dout)
    return dW


def time_embedding_backward(W_shape, xs, dout, dtype=np.float64):                 # :498-507 (sum over t of per-step add.at)
    grad = 0
    for t in range(xs.shape[1]):
        grad = grad + embedding_backward(W_shape, xs[:, t], dout[:, t, :], dtype)
    return grad


def time_softmax_with_loss_forward(xs, ts, ignore_label=-1):                      # :516-536
    N, T, V = xs.shape
    if ts.ndim == 3:
        ts = ts.argmax(axis=2)                                                    # :519-520
    mask = (ts != ignore_label)
    xs2, ts2, mask2 = xs.reshape(N * T, V), ts.reshape(N * T), mask.reshape(N * T)
    ys = _stable_softmax(xs2, 1)                                                  # ad.SoftmaxWithLoss(xs, 2).val :529
    ls = np.log(ys[np.arange(N * T), ts2])
    ls = ls * mask2
    loss = -np.sum(ls) / mask2.sum()
    return loss, (ts2, ys, mask2, (N, T, V))


def time_softmax_with_loss_backward(cache, dout=1):                               # :538-549
    ts, ys, mask, (N, T, V) = cache
    dx = ys.copy()
    dx[np.arange(N * T), ts] -= 1
    dx *= dout
    dx /= mask.sum()
    dx *= mask[:, np.newaxis]
    return dx.reshape((N, T, V))


def multi_cross_entropy_error(y, t):                                              # utils/activation_func.py:4-13
    if y.ndim == 1:
        t = t.reshape(1, t.size)
        y = y.reshape(1, y.size)
    if t.size == y.size:
        t = t.argmax(axis=1)
    batch_size = y.shape[0]
    return -np.sum(np.log(y[np.arange(batch_size), t] + 1e-7)) / batch_size


def mha_forward(query, key, value, wq, wk, wv, wo, num_heads):                    # MultiHeadAttention.forward :1342-1364
    B, S, _ = query.shape
    d_k = wq.shape[1] // num_heads
    Q, K, V = query @ wq, key @ wk, value @ wv
    Qh = Q.reshape(B, S, num_heads, d_k).transpose(0, 2, 1, 3)
    Kh = K.reshape(B, S, num_heads, d_k).transpose(0, 2, 1, 3)
    Vh = V.reshape(B, S, num_heads, d_k).transpose(0, 2, 1, 3)
    scores = (Qh @ Kh.transpose(0, 1, 3, 2)) / np.sqrt(d_k)
    weights = np.exp(scores) / np.sum(np.exp(scores), axis=-1, keepdims=True)     # naive softmax :1356
    out = (weights @ Vh).transpose(0, 2, 1, 3).reshape(B, S, -1)
    return out @ wo


def mha_tape(query, key, value, wq, wk, wv, wo, num_heads):
    """The tape twin Test/MultiHeadAttentionTest.py:57-84 — same op sequence, including building
    exp(scores) twice (:77-78).  Returns (output node, dict of leaves)."""
    B, S, _ = query.shape
    d_k = wq.shape[1] // num_heads
    q, k, v = leaf(query), leaf(key), leaf(value)
    Wq, Wk, Wv, Wo = leaf(wq), leaf(wk), leaf(wv), leaf(wo)
    _q = op("tensordot", q, Wq, axes=([-1], [-2]))                                # :70
    _k = op("tensordot", k, Wk, axes=([-1], [-2]))
    _v = op("tensordot", v, Wv, axes=([-1], [-2]))
    split = lambda n: op("transpose", op("reshape", n, shape=(B, S, num_heads, d_k)), axes=(0, 2, 1, 3))  # :73-75
    qh, kh, vh = split(_q), split(_k), split(_v)
    scores = op("div", op("matmul", qh, op("transpose", kh, axes=(0, 1, 3, 2))), op("sqrt", d_k))   # :76
    e = op("exp", scores)                                                         # :77
    weights = op("div", op("exp", scores), op("sum", e, axis=-1, keepdims=True))  # :78
    ao = op("matmul", weights, vh)                                                # :79
    concat = op("reshape", op("transpose", ao, axes=(0, 2, 1, 3)), shape=(B, S, num_heads * d_k))   # :80
    out = op("tensordot", concat, Wo, axes=([2], [0]))                            # :81
    return out, {"query": q, "key": k, "value": v, "wq": Wq, "wk": Wk, "wv": Wv, "wo": Wo}


# ======================================================================================================
# optimizers — utils/optimizer.py, layers/layers.py
# ======================================================================================================
class Adam:                                                                       # utils/optimizer.py:4-29
    def __init__(self, lr=0.001, beta1=0.9, beta2=0.999):
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2
        self.iter, self.m, self.v = 0, None, None

    def update(self, params, grads):
        if self.m is None:
            self.m = [np.zeros_like(p) for p in params]
            self.v = [np.zeros_like(p) for p in params]
        self.iter += 1
        lr_t = self.lr * np.sqrt(1.0 - self.beta2 ** self.iter) / (1.0 - self.beta1 ** self.iter)   # :22-23
        for i in range(len(params)):
            self.m[i] += (1 - self.beta1) * (grads[i] - self.m[i])               # :26
            self.v[i] += (1 - self.beta2) * (grads[i] ** 2 - self.v[i])          # :27
            params[i] -= lr_t * self.m[i] / (np.sqrt(self.v[i]) + 1e-7)          # :29


class SGD:                                                                        # layers.py:821-831
    def __init__(self, lr=0.01):
        self.lr = lr

    def update(self, params, grads):
        for i in range(len(params)):
            params[i] -= self.lr * grads[i]


def clip_grads(grads, max_norm):                                                  # layers.py:132-141
    total_norm = 0
    for grad in grads:
        total_norm += np.sum(grad ** 2)
    total_norm = np.sqrt(total_norm)
    rate = max_norm / (total_norm + 1e-6)
    if rate < 1:
        for grad in grads:
            grad *= rate
    return total_norm


# ======================================================================================================
# composite workloads used by bench.py's CPU arm and by the parity tests
# ======================================================================================================
def mlp_tape_step(x, labels, W1, b1, W2, b2):
    """Config 1: softmax(tanh(x@W1+b1)@W2+b2, 2) on the tape + multi_cross_entropy_error;
    the upstream gradient of the mean CE w.r.t. the probabilities is fed to gradient()."""
    X, w1, B1, w2, B2 = leaf(x), leaf(W1), leaf(b1), leaf(W2), leaf(b2)
    h = op("tanh", op("add", op("matmul", X, w1), B1))
    y = op("softmax", op("add", op("matmul", h, w2), B2), dim=2)
    loss = multi_cross_entropy_error(y.val, labels)
    n = x.shape[0]
    dy = np.zeros_like(y.val)
    dy[np.arange(n), labels] = -1.0 / (y.val[np.arange(n), labels] + 1e-7) / n
    backward(y, dy)
    return loss, y.val, [w1.grad, B1.grad, w2.grad, B2.grad], X.grad


def c5_model_step(x, labels, params, heads=8):
    """Config 5 model: MHA(q=k=v=x) -> mean over the sequence -> affine 512->256 + tanh ->
    affine 256->10 -> softmax-CE, forward + backward; returns (loss, grads in params order).
    MHA runs on the tape twin (mha_tape); the head reuses the layer formulas."""
    wq, wk, wv, wo, W1, b1, W2, b2 = params
    B, S, Dm = x.shape
    out, leaves = mha_tape(x, x, x, wq, wk, wv, wo, heads)
    a = out.val                                   # (B,S,H)
    pooled = a.sum(axis=1) / S                    # mean over the sequence
    z1 = pooled @ W1 + b1
    h1 = np.tanh(z1)
    logits = h1 @ W2 + b2
    loss, cache = time_softmax_with_loss_forward(logits.reshape(B, 1, -1), labels.reshape(B, 1))
    dlogits = time_softmax_with_loss_backward(cache).reshape(B, -1)
    dW2, db2 = h1.T @ dlogits, dlogits.sum(axis=0)
    dh1 = dlogits @ W2.T
    dz1 = dh1 * (1 - h1 * h1)
    dW1, db1 = pooled.T @ dz1, dz1.sum(axis=0)
    dpooled = dz1 @ W1.T
    da = np.repeat(dpooled[:, None, :], S, axis=1) / S
    backward(out, da)
    g = leaves
    return loss, [g["wq"].grad, g["wk"].grad, g["wv"].grad, g["wo"].grad, dW1, db1, dW2, db2], \
        g["query"].grad + g["key"].grad + g["value"].grad


# ======================================================================================================
# Bahdanau-style dot attention over encoder states — layers/layers.py:964-1083
# ======================================================================================================
def weight_sum_forward(hs, a):                                                    # WeightSum.forward :970-981
    N, T, H = hs.shape
    return np.sum(np.reshape(a, (N, T, 1)) * hs, axis=1)


def weight_sum_backward(hs, a, dout):                                             # WeightSum.backward :983-989
    N, T, H = hs.shape
    d = np.reshape(dout, (N, 1, H))
    da = np.reshape(np.sum(d * hs, axis=2, keepdims=True), (N, T))
    dhs = np.repeat(d, repeats=T, axis=1) * np.reshape(a, (N, T, 1))
    return dhs, da


def attention_weight_forward(hs, h):                                              # AttentionWeight.forward :998-1007
    N, T, H = hs.shape
    c = np.sum(np.repeat(np.reshape(h, (N, 1, H)), T, 1) * hs, axis=2)
    return _stable_softmax(c, c.ndim - 1)


def attention_weight_backward(hs, h, dout):                                       # AttentionWeight.backward :1009-1030
    N, T, H = hs.shape
    rep = np.repeat(np.reshape(h, (N, 1, H)), repeats=T, axis=1)
    sm = _stable_softmax(np.sum(rep * hs, axis=2), 1)
    common = sm * dout - sm * np.sum(sm * dout, axis=1, keepdims=True)
    common_e = np.repeat(common[..., np.newaxis], H, axis=2)
    dh = np.reshape(np.sum(common_e * hs, axis=1, keepdims=True), (N, H))
    return common_e * rep, dh


def attention_forward(hs, h):                                                     # Attention.forward :1040-1045
    a = attention_weight_forward(hs, h)
    return weight_sum_forward(hs, a), a


def attention_backward(hs, h, a, dout):                                           # Attention.backward :1047-1052
    hs0, da = weight_sum_backward(hs, a, dout)
    hs1, dh = attention_weight_backward(hs, h, da)
    return hs0 + hs1, dh


def time_attention_forward(hs_enc, hs_dec):                                       # TimeAttention.forward :1061-1071
    N, Td, H = hs_dec.shape
    out = np.empty(hs_dec.shape)
    weights = []
    for i in range(Td):
        out[:, i, :], a = attention_forward(hs_enc, hs_dec[:, i, :])
        weights.append(a)
    return out, weights


def time_attention_backward(hs_enc, hs_dec, weights, dout):                       # TimeAttention.backward :1073-1083
    dhs_dec = np.empty(dout.shape)
    dhs_enc = 0
    for i in range(dout.shape[1]):
        dhs, dh = attention_backward(hs_enc, hs_dec[:, i, :], weights[i], dout[:, i, :])
        dhs_enc = dhs_enc + dhs
        dhs_dec[:, i, :] = dh
    return dhs_enc, dhs_dec


# ======================================================================================================
# AttentionSeq2seq — layers/layers.py:1086-1208 (encoder, attention decoder, loss; clip_grad + optimizer in fit)
# ======================================================================================================
def attention_seq2seq_step(xs, ts, weights):
    """One forward + backward of AttentionSeq2seq (:942-953 through AttentionEncoder :1159-1168 and AttentionDecoder
    :1111-1136).  weights = [enc_embed, enc_wx, enc_wh, enc_b, dec_embed, dec_wx, dec_wh, dec_b, affine_W, affine_b]
    (the order of model.weights, :1175).  Returns (loss, grads in the same order)."""
    e_emb, e_wx, e_wh, e_b, d_emb, d_wx, d_wh, d_b, a_W, a_b = weights
    dec_in, dec_tgt = ts[:, :-1], ts[:, 1:]                                        # :943
    # encoder (:1160-1163)
    e_x = embedding_forward(e_emb, xs)
    enc_hs, e_caches, _, _ = time_lstm_forward(e_x, e_wx, e_wh, e_b)
    # decoder (:1111-1121): the decoder LSTM starts from the encoder's last hidden state, cell state zero
    d_x = embedding_forward(d_emb, dec_in)
    dec_hs, d_caches, _, _ = time_lstm_forward(d_x, d_wx, d_wh, d_b, h0=enc_hs[:, -1])
    c, att_w = time_attention_forward(enc_hs, dec_hs)
    cat = np.concatenate((c, dec_hs), axis=2)
    score = time_affine_forward(cat, a_W, a_b)
    loss, lcache = time_softmax_with_loss_forward(score, dec_tgt)                   # :946
    # backward (:950-952, :1123-1136)
    dscore = time_softmax_with_loss_backward(lcache, 1)
    dcat, g_aW, g_ab = time_affine_backward(cat, a_W, dscore)
    H = dec_hs.shape[2]
    dc, ddec0 = dcat[:, :, :H], dcat[:, :, H:]
    denc_hs, ddec1 = time_attention_backward(enc_hs, dec_hs, att_w, dc)
    d_dx, d_grads, dh0 = time_lstm_backward(d_caches, d_wx, d_wh, ddec0 + ddec1)
    denc_hs = denc_hs.copy()
    denc_hs[:, -1] += dh0                                                          # :1133
    g_demb = time_embedding_backward(d_emb.shape, dec_in, d_dx)
    e_dx, e_grads, _ = time_lstm_backward(e_caches, e_wx, e_wh, denc_hs)           # :1165-1167
    g_eemb = time_embedding_backward(e_emb.shape, xs, e_dx)
    return loss, [g_eemb, e_grads[0], e_grads[1], e_grads[2], g_demb, d_grads[0], d_grads[1], d_grads[2], g_aW, g_ab]


def seq2seq_clip_grad(grads, max_grad):                                           # fit's local clip_grad :1192-1199
    total = 0.0
    for g in grads:
        total += np.sum(g ** 2)
    norm = np.sqrt(total)
    if norm > max_grad:
        for g in grads:
            g *= max_grad / norm
    return norm


def attention_seq2seq_generate(question, weights, start_id, sample_size):         # :955-958 + AttentionDecoder.generate :1138-1156
    e_emb, e_wx, e_wh, e_b, d_emb, d_wx, d_wh, d_b, a_W, a_b = weights
    enc_hs, _, _, _ = time_lstm_forward(embedding_forward(e_emb, question), e_wx, e_wh, e_b)
    h, cst = enc_hs[:, -1], None
    sampled, sid = [], start_id
    for _ in range(sample_size):
        x = np.array([sid]).reshape(1, 1)
        dec_hs, _, h, cst = time_lstm_forward(embedding_forward(d_emb, x), d_wx, d_wh, d_b, h0=h, c0=cst)
        c, _ = time_attention_forward(enc_hs, dec_hs)
        score = time_affine_forward(np.concatenate((c, dec_hs), axis=2), a_W, a_b)
        sid = int(np.argmax(score.flatten()))
        sampled.append(sid)
    return sampled


# ------------------------------------------------------------------------------------------------------
# Rnnlm (layers.py:699-753) and RnnlmTrainer.get_batch (:765-777)
# ------------------------------------------------------------------------------------------------------
def dropout_masks(shapes, ratio=0.5):                                             # TimeDropout.forward :685-689
    """The masks one training forward draws from the global NumPy RNG, in layer order."""
    scale = 1 / (1.0 - ratio)
    return [(np.random.rand(*s) > ratio).astype(np.float32) * scale for s in shapes]


def rnnlm_step(xs, ts, weights, masks=None, state=None):
    """One forward + backward of Rnnlm.  weights = [embed_W, wx1, wh1, b1, wx2, wh2, b2, affine_b] (the list
    remove_duplicate leaves: the tied affine_W = embed_W.T (:711) is folded into the embedding).  `masks` = three
    dropout masks (N,T,H) or None for eval mode; `state` = (h1, c1, h2, c2) carried between calls (stateful=True
    :718-720) or None.  Returns (loss, merged grads in the order of `weights`, new state, score)."""
    emb, wx1, wh1, b1, wx2, wh2, b2, a_b = weights
    m0, m1, m2 = masks if masks is not None else (1.0, 1.0, 1.0)
    h1, c1, h2, c2 = state if state is not None else (None, None, None, None)
    x0 = embedding_forward(emb, xs) * m0
    hs1, caches1, h1, c1 = time_lstm_forward(x0, wx1, wh1, b1, h0=h1, c0=c1)
    x1 = hs1 * m1
    hs2, caches2, h2, c2 = time_lstm_forward(x1, wx2, wh2, b2, h0=h2, c0=c2)
    x2 = hs2 * m2
    score = time_affine_forward(x2, emb.T, a_b)
    loss, lcache = time_softmax_with_loss_forward(score, ts)
    dscore = time_softmax_with_loss_backward(lcache, 1)
    dx2, g_aW, g_ab = time_affine_backward(x2, emb.T, dscore)
    dx1, g2, _ = time_lstm_backward(caches2, wx2, wh2, dx2 * m2)
    dx0, g1, _ = time_lstm_backward(caches1, wx1, wh1, dx1 * m1)
    g_emb = time_embedding_backward(emb.shape, xs, dx0 * m0)
    g_emb = g_emb + g_aW.T                                                         # remove_duplicate :96-98
    return loss, [g_emb, g1[0], g1[1], g1[2], g2[0], g2[1], g2[2], g_ab], (h1, c1, h2, c2), score


def rnnlm_get_batch(x, t, batch_size, time_size, time_idx):                        # RnnlmTrainer.get_batch :765-777
    data_size = len(x)
    jump = data_size // batch_size
    bx = np.empty((batch_size, time_size), dtype=np.int32)
    bt = np.empty((batch_size, time_size), dtype=np.int32)
    for k in range(time_size):
        for i in range(batch_size):
            bx[i, k] = x[(i * jump + time_idx) % data_size]
            bt[i, k] = t[(i * jump + time_idx) % data_size]
        time_idx += 1
    return bx, bt, time_idx


# ------------------------------------------------------------------------------------------------------
# Seq2Seq with the peeky Decoder (layers.py:866-958)
# ------------------------------------------------------------------------------------------------------
def peeky_seq2seq_step(xs, ts, weights):
    """One forward + backward.  weights = [enc_embed, enc_wx, enc_wh, enc_b, dec_embed, dec_wx (D+H, 4H), dec_wh, dec_b,
    affine_W (2H, V), affine_b] (the order of model.weights :939).  Returns (loss, grads in the same order)."""
    e_emb, e_wx, e_wh, e_b, d_emb, d_wx, d_wh, d_b, a_W, a_b = weights
    dec_in, dec_tgt = ts[:, :-1], ts[:, 1:]                                        # :943
    enc_hs, e_caches, _, _ = time_lstm_forward(embedding_forward(e_emb, xs), e_wx, e_wh, e_b)
    h = enc_hs[:, -1, :]                                                           # Encoder.forward :855
    N, T = dec_in.shape
    H = h.shape[1]
    hidden = np.repeat(h, T, axis=0).reshape(N, T, H)                              # :893
    d_x = np.concatenate((hidden, embedding_forward(d_emb, dec_in)), axis=-1)      # :894
    dec_hs, d_caches, _, _ = time_lstm_forward(d_x, d_wx, d_wh, d_b, h0=h)         # set_state(h) :889
    cat = np.concatenate((hidden, dec_hs), axis=-1)                                # :896
    score = time_affine_forward(cat, a_W, a_b)
    loss, lcache = time_softmax_with_loss_forward(score, dec_tgt)
    dscore = time_softmax_with_loss_backward(lcache, 1)
    dcat, g_aW, g_ab = time_affine_backward(cat, a_W, dscore)
    dhs0, ddec = dcat[:, :, :H], dcat[:, :, H:]                                    # :903
    d_dx, d_grads, dh_lstm = time_lstm_backward(d_caches, d_wx, d_wh, ddec)
    dhs1, demb = d_dx[:, :, :H], d_dx[:, :, H:]                                    # :905
    g_demb = time_embedding_backward(d_emb.shape, dec_in, demb)
    dh = dh_lstm + np.sum(dhs0 + dhs1, axis=1)                                     # :907-908
    denc = np.zeros_like(enc_hs)
    denc[:, -1, :] = dh                                                            # Encoder.backward :858-859
    e_dx, e_grads, _ = time_lstm_backward(e_caches, e_wx, e_wh, denc)
    g_eemb = time_embedding_backward(e_emb.shape, xs, e_dx)
    return loss, [g_eemb, e_grads[0], e_grads[1], e_grads[2], g_demb, d_grads[0], d_grads[1], d_grads[2], g_aW, g_ab]


def peeky_seq2seq_generate(question, weights, start_id, size):                    # :955-958 + Decoder.generate :911-930
    e_emb, e_wx, e_wh, e_b, d_emb, d_wx, d_wh, d_b, a_W, a_b = weights
    enc_hs, _, _, _ = time_lstm_forward(embedding_forward(e_emb, question), e_wx, e_wh, e_b)
    h0 = enc_hs[:, -1, :]
    hidden = h0.reshape(1, 1, -1)
    h, cst, sid, sampled = h0, None, start_id, []
    for _ in range(size):
        x = np.concatenate((hidden, embedding_forward(d_emb, np.array([[sid]]))), axis=-1)
        out, _, h, cst = time_lstm_forward(x, d_wx, d_wh, d_b, h0=h, c0=cst)
        score = time_affine_forward(np.concatenate((hidden, out), axis=-1), a_W, a_b)
        sid = int(np.argmax(score, axis=2).reshape(-1)[0])
        sampled.append(sid)
    return sampled
