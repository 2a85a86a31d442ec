"""Pins oracle/np_oracle.py against the committed golden vectors (produced by the UNMODIFIED reference,
oracle/make_golden.py) and, where /root/reference exists, against the live reference on fresh inputs."""
import numpy as np
import pytest

from conftest import golden
from oracle import np_oracle as O
from oracle import ref_loader

TOL = dict(rtol=1e-12, atol=1e-12)
UN = {"sigmoid": "sigmoid", "abs": "abs"}


def _np(a):
    return np.asarray(a)


def test_unary_ops():
    g = golden("unary")
    names = sorted({k.rsplit("_", 1)[0] for k in g})
    assert len(names) >= 20
    for name in names:
        x, up = g[name + "_x"], g[name + "_g"]
        X = O.leaf(x)
        if name == "log":
            node = O.op("log", X, base=3.0)
        elif name == "pow":
            node = O.op("pow", X, power=3)
        elif name == "neg":
            node = O.op("mul", O.leaf(-1), X)
        else:
            node = O.op(name, X)
        O.backward(node, up)
        np.testing.assert_allclose(node.val, g[name + "_y"], err_msg=name, **TOL)
        np.testing.assert_allclose(X.grad, g[name + "_gx"], err_msg=name, **TOL)


def test_unary_extra_ops():
    """cot / arcos (backward expressions as written in the reference), trace; csc forward (its reference backward raises)."""
    g = golden("unary_extra")
    for name in ("cot", "arcos", "trace"):
        X = O.leaf(g[name + "_x"])
        node = O.op(name, X)
        O.backward(node, g[name + "_g"])
        np.testing.assert_allclose(node.val, g[name + "_y"], err_msg=name, **TOL)
        np.testing.assert_allclose(X.grad, g[name + "_gx"], err_msg=name, **TOL)
    X = O.leaf(g["csc_x"])
    node = O.op("csc", X)
    np.testing.assert_allclose(node.val, g["csc_y"], **TOL)
    O.backward(node, np.ones_like(g["csc_x"]))                  # d/dx (1/sin x) = -cos x / sin^2 x
    np.testing.assert_allclose(X.grad, -np.cos(g["csc_x"]) / np.sin(g["csc_x"]) ** 2, **TOL)


def test_binary_ops_all_broadcast_patterns():
    g = golden("binary")
    keys = sorted({k.rsplit("_", 1)[0] for k in g})
    assert len(keys) == 4 * 14 * 2
    for key in keys:
        opname = key.split("_")[0]
        A, B = O.leaf(g[key + "_a"]), O.leaf(g[key + "_b"])
        node = O.op(opname, A, B)
        O.backward(node, g[key + "_g"])
        np.testing.assert_allclose(node.val, g[key + "_y"], err_msg=key, **TOL)
        assert A.grad.shape == g[key + "_ga"].shape and B.grad.shape == g[key + "_gb"].shape, key
        np.testing.assert_allclose(A.grad, g[key + "_ga"], err_msg=key, **TOL)
        np.testing.assert_allclose(B.grad, g[key + "_gb"], err_msg=key, **TOL)
        # the general reduce-to-shape coincides with the reference heuristic on every pattern here
        np.testing.assert_allclose(O.reduce_to_shape(g[key + "_g"], A.shape), O.check_shape(g[key + "_g"], A.shape))


def test_sub_sign_inversion_is_replicated():
    g = golden("binary")
    np.testing.assert_allclose(g["sub_0_0_ga"], -g["sub_0_0_g"])   # reference returns (-g, +g)  tgrad.py:754
    np.testing.assert_allclose(g["sub_0_0_gb"], g["sub_0_0_g"])


def test_structural_ops():
    g = golden("structural")
    for tag, opname in (("mm2", "matmul"), ("mm4", "matmul")):
        A, B = O.leaf(g[tag + "_a"]), O.leaf(g[tag + "_b"])
        node = O.op(opname, A, B)
        O.backward(node, g[tag + "_g"])
        np.testing.assert_allclose(node.val, g[tag + "_y"], **TOL)
        np.testing.assert_allclose(A.grad, g[tag + "_ga"], **TOL)
        np.testing.assert_allclose(B.grad, g[tag + "_gb"], **TOL)
    for tag, axes in (("td", ([-1], [-2])), ("td2", 2)):
        A, B = O.leaf(g[tag + "_a"]), O.leaf(g[tag + "_b"])
        node = O.op("tensordot", A, B, axes=axes)
        O.backward(node, g[tag + "_g"])
        np.testing.assert_allclose(node.val, g[tag + "_y"], **TOL)
        np.testing.assert_allclose(A.grad, g[tag + "_ga"], **TOL)
        np.testing.assert_allclose(B.grad, g[tag + "_gb"], **TOL)
    for ax, code in ((0, 0), (1, 1), (2, 2), (-1, 9)):
        for kd in (True, False):
            key = "sum_%d_%d" % (code, int(kd))
            X = O.leaf(g["sum_x"])
            node = O.op("sum", X, axis=ax, keepdims=kd)
            O.backward(node, g[key + "_g"])
            np.testing.assert_allclose(node.val, g[key + "_y"], **TOL)
            np.testing.assert_allclose(X.grad, g[key + "_gx"], **TOL)
    X = O.leaf(g["sm_x"]); node = O.op("softmax", X, dim=2); O.backward(node, g["sm_g"])
    np.testing.assert_allclose(node.val, g["sm_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["sm_gx"], **TOL)
    X = O.leaf(g["sm4_x"]); node = O.op("softmax", X, dim=4); O.backward(node, g["sm4_g"])
    np.testing.assert_allclose(node.val, g["sm4_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["sm4_gx"], **TOL)
    X = O.leaf(g["sm_x"]); node = O.op("softmax_with_loss", X, dim=2); O.backward(node, g["sm_g"])
    np.testing.assert_allclose(node.val, g["swl_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["swl_gx"], **TOL)
    X = O.leaf(g["tr_x"]); node = O.op("transpose", X, axes=(0, 2, 1, 3)); O.backward(node, g["tr_g"])
    np.testing.assert_allclose(node.val, g["tr_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["tr_gx"], **TOL)
    X = O.leaf(g["tr_x"]); node = O.op("reshape", X, shape=(12, 20)); O.backward(node, g["rs_g"])
    np.testing.assert_allclose(node.val, g["rs_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["rs_gx"], **TOL)
    X = O.leaf(g["sl_x"]); node = O.op("slice", X, index=slice(1, 4)); O.backward(node, g["sl_g"])
    np.testing.assert_allclose(node.val, g["sl_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["sl_gx"], **TOL)
    X = O.leaf(g["sl_x"]); node = O.op("mean", X); O.backward(node, np.array(2.5))
    np.testing.assert_allclose(node.val, g["mean_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["mean_gx"], **TOL)
    X = O.leaf(g["rep_x"]); node = O.op("repeat", X, times=5, axis=1); O.backward(node, g["rep_g"])
    np.testing.assert_allclose(node.val, g["rep_y"], **TOL)
    np.testing.assert_allclose(X.grad, g["rep_gx"], **TOL)
    assert O.argmax_flat(g["arg_x"]) == g["arg_max"] and O.argmin_flat(g["arg_x"]) == g["arg_min"]
    assert O.argmax_flat(g["arg_x"]).dtype == np.int64
    # diamond graph + accumulation across two gradient() calls
    X = O.leaf(g["dia_x"])
    e = O.op("exp", X)
    node = O.op("add", O.op("mul", e, e), O.op("tanh", e))
    O.backward(node, g["dia_g"])
    np.testing.assert_allclose(X.grad, g["dia_gx1"], **TOL)
    O.backward(node, g["dia_g"])
    np.testing.assert_allclose(X.grad, g["dia_gx2"], **TOL)
    np.testing.assert_allclose(node.val, g["dia_y"], **TOL)


def test_mlp_config1_shape_reduced():
    g = golden("mlp")
    loss, y, grads, gx = O.mlp_tape_step(g["x"], g["labels"], g["W1"], g["b1"], g["W2"], g["b2"])
    np.testing.assert_allclose(loss, g["loss"], **TOL)
    np.testing.assert_allclose(y, g["y"], **TOL)
    for got, key in zip(grads, ("gW1", "gb1", "gW2", "gb2")):
        assert got.shape == g[key].shape
        np.testing.assert_allclose(got, g[key], **TOL)
    np.testing.assert_allclose(gx, g["gx"], **TOL)
    params = [g["W1"].copy(), g["b1"].copy(), g["W2"].copy(), g["b2"].copy()]
    opt = O.Adam(lr=1e-3)
    opt.update(params, grads)
    opt.update(params, grads)
    for got, key in zip(params, ("aW1", "ab1", "aW2", "ab2")):
        np.testing.assert_allclose(got, g[key], **TOL)


@pytest.mark.parametrize("tag", ["raw", "scaled"])
def test_mha_tape_twin(tag):
    g = golden("mha_" + tag)
    out, leaves = O.mha_tape(g["query"], g["key"], g["value"], g["wq"], g["wk"], g["wv"], g["wo"], int(g["heads"]))
    np.testing.assert_allclose(out.val, g["out"], rtol=1e-10, atol=1e-10)
    O.backward(out, g["dout"])
    for name in ("wq", "wk", "wv", "wo", "query", "key", "value"):
        ref = g["g" + name]
        np.testing.assert_allclose(leaves[name].grad, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max(), err_msg=name)
    # the layer forward (naive softmax) equals the tape forward
    np.testing.assert_allclose(O.mha_forward(g["query"], g["key"], g["value"], g["wq"], g["wk"], g["wv"], g["wo"],
                                             int(g["heads"])), g["out"], rtol=1e-10, atol=1e-10)


def test_mha_layer_forward():
    g = golden("mha_layer")
    np.testing.assert_allclose(O.mha_forward(g["x"], g["x"], g["x"], g["wq"], g["wk"], g["wv"], g["wo"], 4), g["out"],
                               **TOL)


def test_layers():
    g = golden("layers")
    np.testing.assert_allclose(O.time_affine_forward(g["ta_x"], g["ta_W"], g["ta_b"]), g["ta_y"], **TOL)
    dx, dW, db = O.time_affine_backward(g["ta_x"], g["ta_W"], g["ta_d"])
    np.testing.assert_allclose(dx, g["ta_dx"], **TOL)
    np.testing.assert_allclose(dW, g["ta_dW"], **TOL)
    np.testing.assert_allclose(db, g["ta_db"], **TOL)
    ct, ht, cache = O.lstm_forward(g["lc_x"], g["lc_h"], g["lc_c"], g["lc_wx"], g["lc_wh"], g["lc_b"])
    np.testing.assert_allclose(ct, g["lc_ct"], **TOL)
    np.testing.assert_allclose(ht, g["lc_ht"], **TOL)
    dx, dh, dc, dWx, dWh, db = O.lstm_backward(cache, g["lc_wx"], g["lc_wh"], g["lc_dh"], g["lc_dc"])
    for got, key in ((dx, "lc_dx"), (dh, "lc_dhp"), (dc, "lc_dcp"), (dWx, "lc_dWx"), (dWh, "lc_dWh"), (db, "lc_db")):
        np.testing.assert_allclose(got, g[key], err_msg=key, **TOL)
    # TimeLSTM: the reference keeps hs / dxs / h / c in float32 buffers (layers.py:566-571) -> 1e-6
    hs, caches, h, c = O.time_lstm_forward(g["tl_xs"], g["lc_wx"], g["lc_wh"], g["lc_b"])
    np.testing.assert_allclose(hs, g["tl_hs"], rtol=2e-6, atol=2e-6)
    dxs, grads, dh = O.time_lstm_backward(caches, g["lc_wx"], g["lc_wh"], g["tl_dhs"])
    np.testing.assert_allclose(dxs, g["tl_dxs"], rtol=1e-5, atol=1e-5)
    for got, key in zip(grads, ("tl_dWx", "tl_dWh", "tl_db")):
        np.testing.assert_allclose(got, g[key], rtol=1e-5, atol=1e-5, err_msg=key)
    np.testing.assert_allclose(dh, g["tl_dh"], rtol=1e-5, atol=1e-5)
    # tanh RNN: oracle cell formulas == what the reference TAPE yields for the same two-step expression
    h0 = np.zeros_like(g["rnn_h1"])
    h1 = O.rnn_forward(g["rnn_x0"], h0, g["rnn_wx"], g["rnn_wh"], g["rnn_b"])
    h2 = O.rnn_forward(g["rnn_x1"], h1, g["rnn_wx"], g["rnn_wh"], g["rnn_b"])
    np.testing.assert_allclose(h1, g["rnn_h1"], **TOL)
    np.testing.assert_allclose(h2, g["rnn_h2"], **TOL)
    dh1, dx1, dWx1, dWh1, db1 = O.rnn_backward(g["rnn_x1"], h1, h2, g["rnn_wx"], g["rnn_wh"], g["rnn_g2"])
    dh0, dx0, dWx0, dWh0, db0 = O.rnn_backward(g["rnn_x0"], h0, h1, g["rnn_wx"], g["rnn_wh"], dh1)
    np.testing.assert_allclose(dWx0 + dWx1, g["rnn_dWx"], **TOL)
    np.testing.assert_allclose(dWh0 + dWh1, g["rnn_dWh"], **TOL)
    np.testing.assert_allclose((db0 + db1).reshape(1, -1), g["rnn_db"], **TOL)
    np.testing.assert_allclose(dx0, g["rnn_dx0"], **TOL)
    np.testing.assert_allclose(dx1, g["rnn_dx1"], **TOL)
    # TimeSoftmaxWithLoss
    loss, cache = O.time_softmax_with_loss_forward(g["sl_logits"], g["sl_ts"])
    np.testing.assert_allclose(loss, g["sl_loss"], **TOL)
    np.testing.assert_allclose(O.time_softmax_with_loss_backward(cache), g["sl_dx"], **TOL)
    loss2, cache2 = O.time_softmax_with_loss_forward(g["sl_logits"], g["sl_onehot"])
    np.testing.assert_allclose(loss2, g["sl_loss2"], **TOL)
    np.testing.assert_allclose(O.time_softmax_with_loss_backward(cache2, 2.0), g["sl_dx2"], **TOL)
    # embedding (the reference's float32 output buffer, layers.py:493)
    np.testing.assert_allclose(O.embedding_forward(g["em_W"], g["em_ids"]), g["em_out"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(O.time_embedding_backward(g["em_W"].shape, g["em_ids"], g["em_d"]), g["em_dW"], **TOL)
    np.testing.assert_allclose(O.relu_forward(g["relu_x"]), g["relu_y"], **TOL)
    np.testing.assert_allclose(O.relu_backward(g["relu_x"], g["relu_d"]), g["relu_dx"], **TOL)
    np.testing.assert_allclose(O.feed_forward_forward(g["relu_x"], g["ff_w"], g["ff_b"]), g["ff_y"], **TOL)
    np.testing.assert_allclose(O._stable_softmax(g["relu_x"], 1), g["lsm_y"], **TOL)
    np.testing.assert_allclose(O.multi_cross_entropy_error(g["ce_p"], g["ce_t"]), g["ce_loss"], **TOL)
    np.testing.assert_allclose(O.multi_cross_entropy_error(g["ce_p"], np.eye(5)[g["ce_t"]]), g["ce_loss_onehot"], **TOL)


def test_optimizers():
    g = golden("optim")
    for name, opt in (("adam", O.Adam(lr=0.01)), ("adam_layers", O.Adam(lr=0.01)), ("sgd", O.SGD(lr=0.1))):
        params = [g["p0_%d" % i].copy() for i in range(4)]
        for step in range(3):
            opt.update(params, [g["g%d_%d" % (step, i)].copy() for i in range(4)])
        for i, p in enumerate(params):
            np.testing.assert_allclose(p, g["%s_p%d" % (name, i)], err_msg=name, **TOL)
    for tag, mx in (("clip", 1.0), ("noclip", 1e6)):
        gr = [g["g0_%d" % i].copy() for i in range(4)]
        O.clip_grads(gr, mx)
        for i, a in enumerate(gr):
            np.testing.assert_allclose(a, g["%s_%d" % (tag, i)], **TOL)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference_fresh_inputs():
    """Second pin: fresh random shapes through the live reference (only in the build container)."""
    ad = ref_loader.load_tape()
    rng = np.random.default_rng(2024)
    for trial in range(5):
        n, d, h = rng.integers(2, 9), rng.integers(2, 9), rng.integers(2, 9)
        x, w, b = rng.standard_normal((n, d)), rng.standard_normal((d, h)), rng.standard_normal((1, h))
        g = rng.standard_normal((n, h))
        X, W, B = ad.Matrix(x), ad.Matrix(w), ad.Matrix(b)
        node = ad.Sigmoid(ad.tanh(X @ W + B) * ad.exp(X @ W))
        node.gradient(g)
        Xo, Wo, Bo = O.leaf(x), O.leaf(w), O.leaf(b)
        no = O.op("sigmoid", O.op("mul", O.op("tanh", O.op("add", O.op("matmul", Xo, Wo), Bo)),
                                  O.op("exp", O.op("matmul", Xo, Wo))))
        O.backward(no, g)
        np.testing.assert_allclose(no.val, node.val, **TOL)
        for a, r in ((Xo, X), (Wo, W), (Bo, B)):
            np.testing.assert_allclose(a.grad, r.grad, rtol=1e-11, atol=1e-11)
