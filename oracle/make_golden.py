"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported in place by oracle/ref_loader.py) on seeded inputs.

    python oracle/make_golden.py            # only works where /root/reference exists (build container)

Each .npz holds the inputs and the reference's outputs for one case; tests/test_oracle_golden.py pins
oracle/np_oracle.py against them and the `-m gpu` tests pin the CUDA path.  Reference call-site rules
(SURVEY Appendix A): `node.gradient(g)` alone on its line, `keepdims=` explicit for ad.sum, scalars
wrapped in ad.Matrix, leaves' `.grad` reset per case.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

UNARY = {  # name -> (reference ctor, input domain)
    "exp": ("exp", "normal"), "tanh": ("tanh", "normal"), "sigmoid": ("Sigmoid", "normal"),
    "ln": ("ln", "positive"), "sqrt": ("sqrt", "positive"), "square": ("square", "normal"),
    "abs": ("Abs", "normal"), "sin": ("sin", "normal"), "cos": ("cos", "normal"), "tan": ("tan", "unit"),
    "sec": ("sec", "unit"), "sinh": ("sinh", "normal"), "cosh": ("cosh", "normal"), "arcsin": ("arcsin", "unit"),
    "arctan": ("arctan", "normal"), "arcot": ("arcot", "positive"), "arcsec": ("arcsec", "gt1"),
}
BINARY = {"add": "Add", "sub": "sub", "mul": "Multi", "div": "Divide"}
BCAST = [((6, 8), (6, 8)), ((6, 8), (1, 8)), ((6, 8), (8,)), ((6, 8), (6, 1)), ((6, 8), (1, 1)),
         ((2, 3, 4), (4,)), ((2, 3, 4), (1, 3, 4)), ((2, 3, 4), (2, 1, 4)), ((2, 3, 4), (2, 3, 1)),
         ((2, 3, 4), (3, 4)), ((6, 1), (1, 8)), ((2, 4, 8, 8), (2, 4, 8, 1)), ((2, 4, 8, 8), (8,)),
         ((16, 64), (1, 64))]


def domain(rng, kind, shape):
    if kind == "normal":
        return rng.standard_normal(shape)
    if kind == "positive":
        return rng.uniform(0.5, 1.5, shape)
    if kind == "unit":
        return rng.uniform(-0.9, 0.9, shape)
    if kind == "gt1":
        return rng.uniform(1.1, 3.0, shape)
    raise ValueError(kind)


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("  wrote %-16s %d arrays" % (name, len(arrays)))


def gen_unary(ad):
    rng = np.random.default_rng(101)
    out = {}
    for name, (ctor, kind) in UNARY.items():
        x = domain(rng, kind, (5, 7))
        g = rng.standard_normal((5, 7))
        X = ad.Matrix(x)
        node = getattr(ad, ctor)(X)
        node.gradient(g)
        out[name + "_x"], out[name + "_g"], out[name + "_y"], out[name + "_gx"] = x, g, node.val, X.grad
    # log(base, x) and pow(x, p)
    x = domain(rng, "positive", (5, 7)); g = rng.standard_normal((5, 7))
    X = ad.Matrix(x); node = ad.log(3.0, X)
    node.gradient(g)
    out["log_x"], out["log_g"], out["log_y"], out["log_gx"] = x, g, node.val, X.grad
    X = ad.Matrix(x); node = ad.pow(X, 3)
    node.gradient(g)
    out["pow_x"], out["pow_g"], out["pow_y"], out["pow_gx"] = x, g, node.val, X.grad
    # neg = Multi(Matrix(-1), x)
    X = ad.Matrix(x); node = -X
    node.gradient(g)
    out["neg_x"], out["neg_g"], out["neg_y"], out["neg_gx"] = x, g, node.val, X.grad
    save("unary", **out)


def gen_unary_extra(ad):
    """cot / arcos (backward quirks of :1639, :1438 included) and trace; csc is absent on purpose: its backward raises in
    the reference (`np.csc`), only its forward is recorded."""
    rng = np.random.default_rng(110)
    out = {}
    for name, ctor, kind in (("cot", "cot", "unit"), ("arcos", "arcos", "unit")):
        x = domain(rng, kind, (5, 7))
        g = rng.standard_normal((5, 7))
        X = ad.Matrix(x)
        node = getattr(ad, ctor)(X)
        node.gradient(g)
        out[name + "_x"], out[name + "_g"], out[name + "_y"], out[name + "_gx"] = x, g, node.val, X.grad
    x = domain(rng, "unit", (5, 7))
    out["csc_x"], out["csc_y"] = x, ad.csc(ad.Matrix(x)).val
    x = rng.standard_normal((6, 6)); g = np.float64(rng.standard_normal())
    X = ad.Matrix(x); node = ad.trace(X)
    node.gradient(g)
    out["trace_x"], out["trace_g"], out["trace_y"], out["trace_gx"] = x, g, np.float64(node.val), X.grad
    save("unary_extra", **out)


def gen_binary(ad):
    rng = np.random.default_rng(102)
    out = {}
    for name, ctor in BINARY.items():
        for k, (sa, sb) in enumerate(BCAST):
            for swap in (False, True):
                s1, s2 = (sb, sa) if swap else (sa, sb)
                a = rng.uniform(0.5, 1.5, s1)
                b = rng.uniform(0.5, 1.5, s2)
                A, B = ad.Matrix(a), ad.Matrix(b)
                node = getattr(ad, ctor)(A, B)
                g = rng.standard_normal(node.val.shape)
                node.gradient(g)
                key = "%s_%d_%d" % (name, k, int(swap))
                out[key + "_a"], out[key + "_b"], out[key + "_g"] = a, b, g
                out[key + "_y"], out[key + "_ga"], out[key + "_gb"] = node.val, A.grad, B.grad
    save("binary", **out)


def gen_structural(ad):
    rng = np.random.default_rng(103)
    out = {}
    # 2-D matmul
    a, b = rng.standard_normal((7, 5)), rng.standard_normal((5, 9))
    A, B = ad.Matrix(a), ad.Matrix(b)
    node = A @ B
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(mm2_a=a, mm2_b=b, mm2_g=g, mm2_y=node.val, mm2_ga=A.grad, mm2_gb=B.grad)
    # 4-D batched matmul
    a, b = rng.standard_normal((2, 3, 6, 4)), rng.standard_normal((2, 3, 4, 5))
    A, B = ad.Matrix(a), ad.Matrix(b)
    node = A @ B
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(mm4_a=a, mm4_b=b, mm4_g=g, mm4_y=node.val, mm4_ga=A.grad, mm4_gb=B.grad)
    # tensordot (the MHA projection pattern) and default axes=2
    a, b = rng.standard_normal((3, 5, 8)), rng.standard_normal((8, 6))
    A, B = ad.Matrix(a), ad.Matrix(b)
    node = ad.tensordot(A, B, axes=([-1], [-2]))
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(td_a=a, td_b=b, td_g=g, td_y=node.val, td_ga=A.grad, td_gb=B.grad)
    a, b = rng.standard_normal((3, 4, 5)), rng.standard_normal((4, 5, 6))
    A, B = ad.Matrix(a), ad.Matrix(b)
    node = ad.tensordot(A, B)
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(td2_a=a, td2_b=b, td2_g=g, td2_y=node.val, td2_ga=A.grad, td2_gb=B.grad)
    # sum over each axis, keepdims both ways
    x = rng.standard_normal((3, 4, 5))
    for ax in (0, 1, 2, -1):
        for kd in (True, False):
            X = ad.Matrix(x)
            node = ad.sum(X, axis=ax, keepdims=kd)
            g = rng.standard_normal(node.val.shape)
            node.gradient(g)
            key = "sum_%d_%d" % (ax % 3 if ax >= 0 else 9, int(kd))
            out[key + "_g"], out[key + "_y"], out[key + "_gx"] = g, node.val, X.grad
    out["sum_x"] = x
    # softmax (2-D, dim = rank) and (4-D last axis), SoftmaxWithLoss
    x = rng.standard_normal((6, 10)); g = rng.standard_normal((6, 10))
    X = ad.Matrix(x); node = ad.softmax(X, 2)
    node.gradient(g)
    out.update(sm_x=x, sm_g=g, sm_y=node.val, sm_gx=X.grad)
    x4 = rng.standard_normal((2, 3, 4, 12)); g4 = rng.standard_normal((2, 3, 4, 12))
    X = ad.Matrix(x4); node = ad.softmax(X, 4)
    node.gradient(g4)
    out.update(sm4_x=x4, sm4_g=g4, sm4_y=node.val, sm4_gx=X.grad)
    X = ad.Matrix(x); node = ad.SoftmaxWithLoss(X, 2)
    node.gradient(g)
    out.update(swl_y=node.val, swl_gx=X.grad)
    # reshape / transpose (involutions) / slice
    x = rng.standard_normal((2, 6, 4, 5))
    X = ad.Matrix(x); node = X.transpose(0, 2, 1, 3)
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(tr_x=x, tr_g=g, tr_y=node.val, tr_gx=X.grad)
    X = ad.Matrix(x); node = X.reshape(12, 20)
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(rs_g=g, rs_y=node.val, rs_gx=X.grad)
    x2 = rng.standard_normal((6, 8))
    X = ad.Matrix(x2); node = X[1:4]
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(sl_x=x2, sl_g=g, sl_y=node.val, sl_gx=X.grad)
    # mean / repeat
    X = ad.Matrix(x2); node = ad.mean(X)
    node.gradient(np.array(2.5))
    out.update(mean_y=node.val, mean_gx=X.grad)
    x1 = rng.standard_normal((6, 1))
    X = ad.Matrix(x1); node = ad.repeat(X, 5, 1)
    g = rng.standard_normal(node.val.shape)
    node.gradient(g)
    out.update(rep_x=x1, rep_g=g, rep_y=node.val, rep_gx=X.grad)
    # argmax / argmin incl. ties (first occurrence) on the flattened array
    t = rng.integers(0, 5, (7, 9)).astype(np.float64)
    out.update(arg_x=t, arg_max=ad.argmax(ad.Matrix(t)).val, arg_min=ad.argmin(ad.Matrix(t)).val)
    # leaf accumulation across two gradient() calls + a diamond graph (path multiplicity)
    x = rng.standard_normal((4, 4)); g = rng.standard_normal((4, 4))
    X = ad.Matrix(x)
    e = ad.exp(X)
    node = e * e + ad.tanh(e)
    node.gradient(g)
    first = X.grad.copy()
    node.gradient(g)
    out.update(dia_x=x, dia_g=g, dia_y=node.val, dia_gx1=first, dia_gx2=X.grad)
    save("structural", **out)


def gen_mlp(ad, U):
    rng = np.random.default_rng(1235)
    n, d, h, c = 16, 48, 32, 10
    x = rng.standard_normal((n, d))
    labels = rng.integers(0, c, n)
    W1 = rng.standard_normal((d, h)) / np.sqrt(d); b1 = np.zeros((1, h))
    W2 = rng.standard_normal((h, c)) / np.sqrt(h); b2 = np.zeros((1, c))
    X, w1, B1, w2, B2 = ad.Matrix(x), ad.Matrix(W1), ad.Matrix(b1), ad.Matrix(W2), ad.Matrix(b2)
    y = ad.softmax(ad.tanh(X @ w1 + B1) @ w2 + B2, 2)
    loss = U[2].multi_cross_entropy_error(y.val, labels)
    dy = np.zeros_like(y.val)
    dy[np.arange(n), labels] = -1.0 / (y.val[np.arange(n), labels] + 1e-7) / n
    y.gradient(dy)
    params = [W1.copy(), b1.copy(), W2.copy(), b2.copy()]
    grads = [w1.grad, B1.grad, w2.grad, B2.grad]
    opt = U[0].Adam(lr=1e-3)
    opt.update(params, grads)
    opt.update(params, grads)
    save("mlp", x=x, labels=labels, W1=W1, b1=b1, W2=W2, b2=b2, y=y.val, loss=np.float64(loss),
         gW1=w1.grad, gb1=B1.grad, gW2=w2.grad, gb2=B2.grad, gx=X.grad,
         aW1=params[0], ab1=params[1], aW2=params[2], ab2=params[3])


def gen_mha(ad, L):
    # the reference's own test shapes and seed (Test/MultiHeadAttentionTest.py:93-115), with the
    # constructor-arity bug fixed in memory only
    np.random.seed(42)
    B, S, wordvec, hidden, heads = 10, 30, 16, 20, 4
    w1 = np.random.randn(wordvec, hidden); w2 = np.random.randn(wordvec, hidden)
    w3 = np.random.randn(wordvec, hidden); w4 = np.random.randn(hidden, hidden)
    query = np.random.randn(B, S, wordvec); key = np.random.randn(B, S, wordvec)
    value = np.random.randn(B, S, wordvec)
    # un-normalised randn weights overflow nothing at this size but give huge scores; also record a
    # variance-scaled variant, which is what the benchmark configs use
    for tag, scale in (("raw", 1.0), ("scaled", 1.0 / np.sqrt(wordvec))):
        d_k = hidden // heads
        q, k, v = ad.Matrix(query), ad.Matrix(key), ad.Matrix(value)
        wq, wk, wv, wo = ad.Matrix(w1 * scale), ad.Matrix(w2 * scale), ad.Matrix(w3 * scale), ad.Matrix(w4 * scale)
        _q = ad.tensordot(q, wq, axes=([-1], [-2]))
        _k = ad.tensordot(k, wk, axes=([-1], [-2]))
        _v = ad.tensordot(v, wv, axes=([-1], [-2]))
        qh = _q.reshape(B, S, heads, d_k).transpose(0, 2, 1, 3)
        kh = _k.reshape(B, S, heads, d_k).transpose(0, 2, 1, 3)
        vh = _v.reshape(B, S, heads, d_k).transpose(0, 2, 1, 3)
        scores = (qh @ kh.transpose(0, 1, 3, 2)) / ad.sqrt(d_k)
        e = ad.exp(scores)
        weights = ad.exp(scores) / ad.sum(e, axis=-1, keepdims=True)
        ao = weights @ vh
        concat = ao.transpose(0, 2, 1, 3).reshape(B, S, -1)
        output = ad.tensordot(concat, wo, axes=([2], [0]))
        dout = output.val if tag == "raw" else np.random.default_rng(7).standard_normal(output.val.shape)
        output.gradient(dout)
        save("mha_" + tag, query=query, key=key, value=value, wq=w1 * scale, wk=w2 * scale, wv=w3 * scale,
             wo=w4 * scale, heads=np.int64(heads), out=output.val, dout=dout, gwq=wq.grad, gwk=wk.grad,
             gwv=wv.grad, gwo=wo.grad, gquery=q.grad, gkey=k.grad, gvalue=v.grad)
    # layers.MultiHeadAttention forward (wordvec == hidden required, layers.py:1336-1339)
    rng = np.random.default_rng(104)
    layer = L.MultiHeadAttention(16, 16, 4)
    layer.weights = [rng.standard_normal((16, 16)) / 4 for _ in range(4)]
    x = rng.standard_normal((3, 7, 16))
    out = layer.forward(x, x, x)
    save("mha_layer", x=x, wq=layer.weights[0], wk=layer.weights[1], wv=layer.weights[2], wo=layer.weights[3],
         out=out)


def gen_layers(ad, L, U):
    rng = np.random.default_rng(105)
    out = {}
    # TimeAffine
    N, T, Dm, V = 3, 5, 4, 7
    x = rng.standard_normal((N, T, Dm)); W = rng.standard_normal((Dm, V)); b = rng.standard_normal(V)
    lay = L.TimeAffine(W.copy(), b.copy())
    y = lay.forward(x)
    d = rng.standard_normal(y.shape)
    dx = lay.backward(d)
    out.update(ta_x=x, ta_W=W, ta_b=b, ta_y=y, ta_d=d, ta_dx=dx, ta_dW=lay.grads[0], ta_db=lay.grads[1])
    # LSTM cell + TimeLSTM (fp64 weights; the reference's internal buffers are float32 for hs/dxs)
    H = 6
    wx = rng.standard_normal((Dm, 4 * H)) / 2; wh = rng.standard_normal((H, 4 * H)) / 2; bb = rng.standard_normal(4 * H) / 2
    xc, hc, cc = rng.standard_normal((N, Dm)), rng.standard_normal((N, H)), rng.standard_normal((N, H))
    cell = L.LSTM(wx.copy(), wh.copy(), bb.copy(), dtype=np.float64)
    ct, ht = cell.forward(xc, hc, cc)
    dh, dc = rng.standard_normal((N, H)), rng.standard_normal((N, H))
    dxc, dhc, dcc = cell.backward(dh, dc)
    out.update(lc_x=xc, lc_h=hc, lc_c=cc, lc_wx=wx, lc_wh=wh, lc_b=bb, lc_ct=ct, lc_ht=ht, lc_dh=dh, lc_dc=dc,
               lc_dx=dxc, lc_dhp=dhc, lc_dcp=dcc, lc_dWx=cell.grads[0], lc_dWh=cell.grads[1], lc_db=cell.grads[2])
    xs = rng.standard_normal((N, T, Dm))
    tl = L.TimeLSTM(wx.copy(), wh.copy(), bb.copy())
    hs = tl.forward(xs)
    dhs = rng.standard_normal(hs.shape)
    dxs = tl.backward(dhs)
    out.update(tl_xs=xs, tl_hs=hs, tl_dhs=dhs, tl_dxs=dxs, tl_dWx=tl.grads[0], tl_dWh=tl.grads[1], tl_db=tl.grads[2],
               tl_dh=tl.dh)
    # tanh RNN on the reference TAPE (layers.RNN is unimportable/broken): two steps
    wxr = rng.standard_normal((Dm, H)) / 2; whr = rng.standard_normal((H, H)) / 2; br = rng.standard_normal((1, H)) / 2
    x0, x1 = rng.standard_normal((N, Dm)), rng.standard_normal((N, Dm))
    h0 = np.zeros((N, H))
    X0, X1, WX, WH, BR, H0 = (ad.Matrix(a) for a in (x0, x1, wxr, whr, br, h0))
    h1 = ad.tanh(X0 @ WX + H0 @ WH + BR)
    h2 = ad.tanh(X1 @ WX + h1 @ WH + BR)
    g2 = rng.standard_normal((N, H))
    h2.gradient(g2)
    out.update(rnn_x0=x0, rnn_x1=x1, rnn_wx=wxr, rnn_wh=whr, rnn_b=br, rnn_h1=h1.val, rnn_h2=h2.val, rnn_g2=g2,
               rnn_dWx=WX.grad, rnn_dWh=WH.grad, rnn_db=BR.grad, rnn_dx0=X0.grad, rnn_dx1=X1.grad)
    # TimeSoftmaxWithLoss: integer labels with ignore (-1), and one-hot labels
    logits = rng.standard_normal((N, T, V))
    ts = rng.integers(0, V, (N, T)); ts[0, 1] = -1; ts[2, 4] = -1
    sl = L.TimeSoftmaxWithLoss()
    loss = sl.forward(logits.copy(), ts)
    dl = sl.backward()
    onehot = np.eye(V)[rng.integers(0, V, (N, T))]
    sl2 = L.TimeSoftmaxWithLoss()
    loss2 = sl2.forward(logits.copy(), onehot)
    dl2 = sl2.backward(2.0)
    out.update(sl_logits=logits, sl_ts=ts, sl_loss=np.float64(loss), sl_dx=dl, sl_onehot=onehot,
               sl_loss2=np.float64(loss2), sl_dx2=dl2)
    # Embedding / TimeEmbedding (repeated ids -> accumulation)
    Wm = rng.standard_normal((9, 4))
    ids = rng.integers(0, 9, (N, T)); ids[1] = ids[0]
    te = L.TimeEmbedding(Wm.copy())
    emb = te.forward(ids)
    de = rng.standard_normal(emb.shape)
    te.backward(de)
    out.update(em_W=Wm, em_ids=ids, em_out=emb, em_d=de, em_dW=te.grads[0])
    # ReLU / FeedForward forward / softmax fn
    xr = rng.standard_normal((5, 6)); xr[0, 0] = 0.0
    r = L.ReLU(); yr = r.forward(xr); dr = rng.standard_normal(xr.shape)
    out.update(relu_x=xr, relu_y=yr, relu_d=dr, relu_dx=r.backward(dr))
    wf = rng.standard_normal((6, 8)); bf = rng.standard_normal((1, 8))
    ff = L.FeedForward(wf.copy(), bf.copy(), dtype=np.float64)
    out.update(ff_w=wf, ff_b=bf, ff_y=ff.forward(xr), lsm_y=L.softmax(xr, 1))
    # multi_cross_entropy_error with labels and one-hot
    probs = L.softmax(rng.standard_normal((6, 5)), 1)
    lab = rng.integers(0, 5, 6)
    out.update(ce_p=probs, ce_t=lab, ce_loss=np.float64(U[2].multi_cross_entropy_error(probs, lab)),
               ce_loss_onehot=np.float64(U[2].multi_cross_entropy_error(probs, np.eye(5)[lab])))
    save("layers", **out)


def gen_optim(L, U):
    rng = np.random.default_rng(106)
    shapes = [(5, 7), (7,), (1, 3), (4100,)]
    p0 = [rng.standard_normal(s) for s in shapes]
    gs = [[rng.standard_normal(s) for s in shapes] for _ in range(3)]
    out = {}
    for name, opt in (("adam", U[0].Adam(lr=0.01)), ("adam_layers", L.Adam(lr=0.01)), ("sgd", L.SGD(lr=0.1))):
        params = [p.copy() for p in p0]
        for step in range(3):
            opt.update(params, [g.copy() for g in gs[step]])
        for i, p in enumerate(params):
            out["%s_p%d" % (name, i)] = p
    for i in range(len(shapes)):
        out["p0_%d" % i] = p0[i]
        for s in range(3):
            out["g%d_%d" % (s, i)] = gs[s][i]
    # clip_grads: one case that clips and one that does not
    for tag, mx in (("clip", 1.0), ("noclip", 1e6)):
        g = [a.copy() for a in gs[0]]
        L.clip_grads(g, mx)
        for i, a in enumerate(g):
            out["%s_%d" % (tag, i)] = a
    save("optim", **out)


def gen_attention(L):
    rng = np.random.default_rng(107)
    N, Te, Td, H = 3, 6, 4, 5
    hs = rng.standard_normal((N, Te, H)); h = rng.standard_normal((N, H))
    out = {}
    ws = L.WeightSum(); a = L.softmax(rng.standard_normal((N, Te)), 1)
    c = ws.forward(hs, a); dc = rng.standard_normal(c.shape)
    dhs, da = ws.backward(dc)
    out.update(ws_hs=hs, ws_a=a, ws_c=c, ws_dc=dc, ws_dhs=dhs, ws_da=da)
    aw = L.AttentionWeight(); w = aw.forward(hs, h); dw = rng.standard_normal(w.shape)
    dhs2, dh2 = aw.backward(dw)
    out.update(aw_h=h, aw_w=w, aw_dw=dw, aw_dhs=dhs2, aw_dh=dh2)
    at = L.Attention(); c2 = at.forward(hs, h); dc2 = rng.standard_normal(c2.shape)
    dhs3, dh3 = at.backward(dc2)
    out.update(at_c=c2, at_dc=dc2, at_dhs=dhs3, at_dh=dh3)
    hs_dec = rng.standard_normal((N, Td, H))
    ta = L.TimeAttention(); o = ta.forward(hs, hs_dec); do = rng.standard_normal(o.shape)
    dhs_enc, dhs_dec = ta.backward(do)
    out.update(ta_dec=hs_dec, ta_out=o, ta_do=do, ta_dhs_enc=dhs_enc, ta_dhs_dec=dhs_dec,
               ta_weights=np.stack(ta.hs_weight, axis=1))
    save("attention", **out)


def gen_seq2seq(L):
    """AttentionSeq2seq (layers.py:1171-1181): one forward + backward of the full encoder / attention decoder / loss
    composition, plus clip + one Adam step as `fit` performs them (:1192-1208)."""
    np.random.seed(4321)            # the reference draws its initial weights from the global NumPy RNG
    V, Dm, H = 13, 8, 12
    m = L.AttentionSeq2seq({}, V, Dm, H)
    w0 = [np.array(w, dtype=np.float64) for w in m.weights]
    rng = np.random.default_rng(108)
    N, Te, Td = 3, 7, 5
    xs = rng.integers(0, V, (N, Te))
    ts = rng.integers(0, V, (N, Td + 1))
    loss = m.forward(xs, ts)
    m.backward()
    grads = [np.array(g, dtype=np.float64) for g in m.grads]
    out = dict(xs=xs, ts=ts, loss=np.float64(loss), n=np.int64(len(w0)), dims=np.array([V, Dm, H]))
    for i, (w, g) in enumerate(zip(w0, grads)):
        out["w%d" % i] = w
        out["g%d" % i] = g
    # greedy decoding from the trained-for-zero-steps model (AttentionDecoder.generate :1138-1156)
    m2_question = xs[:1]
    enc_hs = m.encoder.forward(m2_question)
    out["generated"] = np.array(m.decoder.generate(enc_hs, 3, 6), dtype=np.int64)
    save("seq2seq", **out)


def gen_rnnlm(L):
    """Rnnlm + RnnlmTrainer (layers.py:699-818): two iterations of the trainer's inner loop (forward with dropout ON,
    backward, remove_duplicate, clip_grads, SGD) on the tied-weight two-LSTM language model, an eval-mode forward, and
    the trainer's batch walk over a corpus."""
    np.random.seed(777)             # initial weights come from the global NumPy RNG (:702-712)
    V, H = 23, 12
    m = L.Rnnlm(V, H, H)
    w0 = [np.array(w, dtype=np.float64) for w in m.params]
    rng = np.random.default_rng(109)
    N, T, steps, max_grad, lr = 3, 5, 2, 0.25, 0.5
    xs = rng.integers(0, V, (steps, N, T))
    ts = rng.integers(0, V, (steps, N, T))
    out = dict(xs=xs, ts=ts, dims=np.array([V, H]), n=np.int64(len(w0)), max_grad=np.float64(max_grad), lr=np.float64(lr),
               mask_seed=np.int64(4242))
    for i, w in enumerate(w0):
        out["w%d" % i] = w
    opt = L.SGD(lr)
    np.random.seed(4242)            # dropout masks: np.random.rand in layer order, every forward (:687)
    losses = []
    for it in range(steps):
        loss = m.forward(xs[it], ts[it])
        m.backward()
        params, grads = L.remove_duplicate(m.params, m.grads)
        if it == 0:
            for i, g in enumerate(grads):
                out["g%d" % i] = np.array(g, dtype=np.float64)       # merged, before clipping
            out["n_merged"] = np.int64(len(grads))
        L.clip_grads(grads, max_grad)
        opt.update(params, grads)
        losses.append(float(loss))
    out["losses"] = np.array(losses)
    params, _ = L.remove_duplicate(m.params, m.grads)
    for i, p in enumerate(params):
        out["p%d" % i] = np.array(p, dtype=np.float64)
    # eval mode (dropout off), state carried on from the two training steps, then after reset_state
    for layer in m.layers:
        if hasattr(layer, "train_flg"):
            layer.train_flg = False
    out["eval_loss_stateful"] = np.float64(m.forward(xs[0], ts[0]))
    m.reset_state()
    out["eval_loss_reset"] = np.float64(m.forward(xs[0], ts[0]))
    out["eval_score"] = np.array(m.predict(xs[1]), dtype=np.float64)
    # the trainer's corpus walk (:765-777)
    tr = L.RnnlmTrainer(m, opt)
    tr.time_idx = 0
    corpus = np.arange(50) * 3 % 23
    bx1, bt1 = tr.get_batch(corpus[:-1], corpus[1:], 4, 5)
    bx2, bt2 = tr.get_batch(corpus[:-1], corpus[1:], 4, 5)
    out.update(corpus=corpus, bx1=bx1, bt1=bt1, bx2=bx2, bt2=bt2)
    save("rnnlm", **out)


def gen_peeky_seq2seq(L):
    """Seq2Seq with the peeky Decoder (layers.py:866-958): the encoder's last state is concatenated to every decoder LSTM
    input and to every affine input.  One forward + backward, and greedy decoding."""
    np.random.seed(2468)
    V, Dm, H = 11, 6, 10
    m = L.Seq2Seq({"_": 4}, V, Dm, H)
    w0 = [np.array(w, dtype=np.float64) for w in m.weights]
    rng = np.random.default_rng(111)
    N, Te, Td = 3, 6, 4
    xs = rng.integers(0, V, (N, Te))
    ts = rng.integers(0, V, (N, Td + 1))
    loss = m.forward(xs, ts)
    m.backward()
    out = dict(xs=xs, ts=ts, loss=np.float64(loss), n=np.int64(len(w0)), dims=np.array([V, Dm, H]))
    for i, (w, g) in enumerate(zip(w0, m.grads)):
        out["w%d" % i] = w
        out["g%d" % i] = np.array(g, dtype=np.float64)
    out["generated"] = np.array(m.generate(xs[:1], {"_": 4}, 5), dtype=np.int64)
    save("peeky_seq2seq", **out)


def main():
    if not ref_loader.available():
        raise SystemExit("reference not found at %s" % ref_loader.REFERENCE_ROOT)
    os.makedirs(OUT, exist_ok=True)
    ad = ref_loader.load_tape()
    L = ref_loader.load_layers()
    U = ref_loader.load_utils()
    gen_unary(ad)
    gen_unary_extra(ad)
    gen_binary(ad)
    gen_structural(ad)
    gen_mlp(ad, U)
    gen_mha(ad, L)
    gen_layers(ad, L, U)
    gen_optim(L, U)
    gen_attention(L)
    gen_seq2seq(L)
    gen_rnnlm(L)
    gen_peeky_seq2seq(L)


if __name__ == "__main__":
    main()
