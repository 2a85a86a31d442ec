"""Layer / optimizer parity (layers.py, utils/optimizer.py mirrors) vs the reference's golden vectors and vs
the oracle at the config shapes (reduced where the CPU oracle would take minutes)."""
import numpy as np
import pytest

from conftest import assert_close, golden
from oracle import np_oracle as O


def _L():
    from auto_differentiation_b200.layers import layers
    return layers


def test_time_affine_relu_feedforward(backend):
    L = _L()
    g = golden("layers")
    lay = L.TimeAffine(g["ta_W"], g["ta_b"])
    assert_close(lay.forward(g["ta_x"]), g["ta_y"], what="TimeAffine fwd")
    w_id, g_id = id(lay.weights[0]), id(lay.grads[0])
    assert_close(lay.backward(g["ta_d"]), g["ta_dx"], what="TimeAffine dx")
    assert_close(lay.grads[0], g["ta_dW"], what="TimeAffine dW")
    assert_close(lay.grads[1], g["ta_db"], what="TimeAffine db")
    assert id(lay.weights[0]) == w_id and id(lay.grads[0]) == g_id      # grads[i][...] = : identity preserved
    r = L.ReLU()
    assert_close(r.forward(g["relu_x"]), g["relu_y"], what="ReLU fwd")
    assert_close(r.backward(g["relu_d"]), g["relu_dx"], what="ReLU bwd (0 at x == 0)")
    ff = L.FeedForward(g["ff_w"], g["ff_b"])
    y = ff.forward(g["relu_x"])
    assert_close(y, g["ff_y"], what="FeedForward fwd")
    d = np.random.default_rng(3).standard_normal(g["ff_y"].shape)
    dx = ff.backward(d)
    odx, odw, odb = O.feed_forward_backward(g["relu_x"], g["ff_w"], g["ff_y"], d)
    assert_close(dx, odx, what="FeedForward dx"); assert_close(ff.grads[0], odw, what="FeedForward dW")
    assert_close(ff.grads[1], odb, what="FeedForward db")
    assert_close(L.softmax(g["relu_x"], 1), g["lsm_y"], what="layers.softmax")


def test_lstm_cell_and_time_lstm(backend):
    L = _L()
    g = golden("layers")
    cell = L.LSTM(g["lc_wx"], g["lc_wh"], g["lc_b"])
    ct, ht = cell.forward(g["lc_x"], g["lc_h"], g["lc_c"])          # returns (c', h') like the reference
    assert_close(ct, g["lc_ct"], what="LSTM c'"); assert_close(ht, g["lc_ht"], what="LSTM h'")
    dx, dh, dc = cell.backward(g["lc_dh"], g["lc_dc"])
    for got, key in ((dx, "lc_dx"), (dh, "lc_dhp"), (dc, "lc_dcp"), (cell.grads[0], "lc_dWx"),
                     (cell.grads[1], "lc_dWh"), (cell.grads[2], "lc_db")):
        assert_close(got, g[key], what=key)
    tl = L.TimeLSTM(g["lc_wx"], g["lc_wh"], g["lc_b"])
    hs = tl.forward(g["tl_xs"])
    # the reference's own hs/dxs buffers are float32 (layers.py:566,586): golden carries ~1e-7 rounding
    assert_close(hs, g["tl_hs"], rtol=1e-5, atol=2e-6, what="TimeLSTM hs")
    dxs = tl.backward(g["tl_dhs"])
    assert_close(dxs, g["tl_dxs"], rtol=1e-5, atol=2e-6, what="TimeLSTM dxs")
    for got, key in ((tl.grads[0], "tl_dWx"), (tl.grads[1], "tl_dWh"), (tl.grads[2], "tl_db"), (tl.dh, "tl_dh")):
        assert_close(got, g[key], rtol=1e-5, atol=2e-6, what=key)


def test_time_lstm_config4_shape_reduced_vs_oracle(backend):
    """Config 4 pipeline pieces at parity size (T=16, N=4; H reduced to 64 for the CPU oracle)."""
    L = _L()
    rng = np.random.default_rng(1238)
    N, T, Dm, H = 4, 16, 32, 64
    xs = rng.standard_normal((N, T, Dm))
    wx = rng.standard_normal((Dm, 4 * H)) / np.sqrt(Dm); wh = rng.standard_normal((H, 4 * H)) / np.sqrt(H)
    b = rng.standard_normal(4 * H) * 0.1
    dhs = rng.standard_normal((N, T, H))
    hs_o, caches, _, _ = O.time_lstm_forward(xs, wx, wh, b)
    dxs_o, grads_o, dh_o = O.time_lstm_backward(caches, wx, wh, dhs)
    tl = L.TimeLSTM(wx, wh, b)
    assert_close(tl.forward(xs), hs_o, what="hs")
    assert_close(tl.backward(dhs), dxs_o, rtol=2e-5, atol=2e-6, what="dxs")
    for got, ref, name in zip(tl.grads, grads_o, ("dWx", "dWh", "db")):
        assert_close(got, ref, rtol=2e-5, atol=2e-6, what=name)
    assert_close(tl.dh, dh_o, rtol=2e-5, atol=2e-6, what="dh")
    # stateful: a second forward continues from h/c
    tl2 = L.TimeLSTM(wx, wh, b, stateful=True)
    tl2.forward(xs[:, :8]); hs2 = tl2.forward(xs[:, 8:])
    assert_close(hs2, hs_o[:, 8:], what="stateful continuation")


def test_rnn_cell_and_time_rnn(backend):
    L = _L()
    g = golden("layers")
    # two steps by hand with the RNN cell == the reference tape's gradients for the same expression
    c0 = L.RNN(g["rnn_wx"], g["rnn_wh"], g["rnn_b"]); c1 = L.RNN(g["rnn_wx"], g["rnn_wh"], g["rnn_b"])
    h0 = np.zeros_like(g["rnn_h1"])
    h1 = c0.forward(g["rnn_x0"], h0); h2 = c1.forward(g["rnn_x1"], h1)
    assert_close(h1, g["rnn_h1"], what="rnn h1"); assert_close(h2, g["rnn_h2"], what="rnn h2")
    dh1, dx1 = c1.backward(g["rnn_g2"]); dh0, dx0 = c0.backward(dh1)
    assert_close(dx0, g["rnn_dx0"], what="rnn dx0"); assert_close(dx1, g["rnn_dx1"], what="rnn dx1")
    assert_close(np.asarray(c0.grads[0]) + np.asarray(c1.grads[0]), g["rnn_dWx"], what="rnn dWx")
    assert_close(np.asarray(c0.grads[1]) + np.asarray(c1.grads[1]), g["rnn_dWh"], what="rnn dWh")
    assert_close(np.asarray(c0.grads[2]) + np.asarray(c1.grads[2]), g["rnn_db"], what="rnn db")
    # TimeRNN vs oracle (config-4 cell, reduced)
    rng = np.random.default_rng(1239)
    N, T, Dm, H = 4, 16, 24, 48
    xs = rng.standard_normal((N, T, Dm)); wx = rng.standard_normal((Dm, H)) / np.sqrt(Dm)
    wh = rng.standard_normal((H, H)) / np.sqrt(H); b = rng.standard_normal(H) * 0.1
    dhs = rng.standard_normal((N, T, H))
    hs_o, hprevs = O.time_rnn_forward(xs, wx, wh, b)
    dxs_o, (gwx, gwh, gb), dh_o = O.time_rnn_backward(xs, hs_o, hprevs, wx, wh, dhs)
    tr = L.TimeRNN(wx, wh, b)
    assert_close(tr.forward(xs), hs_o, what="TimeRNN hs")
    assert_close(tr.backward(dhs), dxs_o, rtol=2e-5, atol=2e-6, what="TimeRNN dxs")
    for got, ref, name in zip(tr.grads, (gwx, gwh, gb), ("dWx", "dWh", "db")):
        assert_close(got, ref, rtol=2e-5, atol=2e-6, what="TimeRNN " + name)


def test_time_softmax_with_loss_embedding_ce(backend):
    L = _L()
    from auto_differentiation_b200.utils.activation_func import multi_cross_entropy_error
    g = golden("layers")
    sl = L.TimeSoftmaxWithLoss()
    loss = sl.forward(g["sl_logits"], g["sl_ts"])                    # labels with ignore (-1)
    assert abs(float(loss) - float(g["sl_loss"])) < 1e-5
    assert_close(sl.backward(), g["sl_dx"], what="TSWL dx")
    sl2 = L.TimeSoftmaxWithLoss()
    loss2 = sl2.forward(g["sl_logits"], g["sl_onehot"])              # one-hot -> argmax(axis=2), bit-exact labels
    assert abs(float(loss2) - float(g["sl_loss2"])) < 1e-5
    assert_close(sl2.backward(2.0), g["sl_dx2"], what="TSWL dx (one-hot, dout=2)")
    te = L.TimeEmbedding(g["em_W"])
    out = te.forward(g["em_ids"])
    np.testing.assert_array_equal(np.asarray(out), g["em_W"].astype(np.float32)[g["em_ids"]])   # gather is exact
    te.backward(g["em_d"])
    assert_close(te.grads[0], g["em_dW"], what="TimeEmbedding dW (repeated ids accumulate)")
    e = L.Embedding(g["em_W"])
    np.testing.assert_array_equal(np.asarray(e.forward(g["em_ids"][:, 0])), g["em_W"].astype(np.float32)[g["em_ids"][:, 0]])
    assert abs(float(multi_cross_entropy_error(g["ce_p"], g["ce_t"])) - float(g["ce_loss"])) < 1e-5
    assert abs(float(multi_cross_entropy_error(g["ce_p"], np.eye(5)[g["ce_t"]])) - float(g["ce_loss_onehot"])) < 1e-5
    # vocabulary-scale rows (config 4: V = 1000) vs the oracle, with ignored positions
    rng = np.random.default_rng(1240)
    N, T, V = 4, 16, 1000
    logits = rng.standard_normal((N, T, V)) * 3
    ts = rng.integers(0, V, (N, T)); ts[1, 3] = -1; ts[3, 15] = -1
    lo, cache = O.time_softmax_with_loss_forward(logits, ts)
    sl3 = L.TimeSoftmaxWithLoss()
    assert abs(float(sl3.forward(logits, ts)) - lo) < 2e-5
    assert_close(sl3.backward(), O.time_softmax_with_loss_backward(cache), what="TSWL V=1000")


@pytest.mark.parametrize("tag", ["raw", "scaled"])
def test_mha_layer_vs_tape_twin_golden(backend, tag):
    """layers.MultiHeadAttention forward (layers.py:1342-1364) + the backward the reference only has on the
    tape (Test/MultiHeadAttentionTest.py:49-91): outputs and all seven gradients vs the reference golden."""
    L = _L()
    g = golden("mha_" + tag)
    heads = int(g["heads"])
    layer = L.MultiHeadAttention(g["wq"].shape[0], g["wq"].shape[1], heads,
                                 weights=[g["wq"], g["wk"], g["wv"], g["wo"]])
    out = layer.forward(g["query"], g["key"], g["value"])
    # "raw": un-normalised randn weights -> |scores| up to ~1e2, outputs/grads up to ~1e4; the fp32 score
    # rounding (|s|*2^-24) is amplified by exp(): tolerance 5e-4 relative to the tensor scale, stated here.
    rtol, atol = (5e-4, 5e-4) if tag == "raw" else (1e-5, 1e-6)
    assert_close(out, g["out"], rtol=rtol, atol=atol, what="MHA out")
    dq, dk, dv = layer.backward(g["dout"])
    for got, name in ((layer.grads[0], "wq"), (layer.grads[1], "wk"), (layer.grads[2], "wv"), (layer.grads[3], "wo"),
                      (dq, "query"), (dk, "key"), (dv, "value")):
        assert_close(got, g["g" + name], rtol=rtol, atol=atol, what="MHA g" + name)


def test_mha_layer_forward_golden_and_config3_reduced(backend):
    L = _L()
    g = golden("mha_layer")
    layer = L.MultiHeadAttention(16, 16, 4, weights=[g["wq"], g["wk"], g["wv"], g["wo"]])
    assert_close(layer.forward(g["x"], g["x"], g["x"]), g["out"], what="MHA layer fwd")
    # config-3 structure (8 heads, d_k = 64) at a size the CPU oracle finishes in seconds: B=2, S=64, d=512
    rng = np.random.default_rng(1237)
    B, S, d = 2, 64, 512
    x = rng.standard_normal((B, S, d))
    ws = [rng.standard_normal((d, d)) / np.sqrt(d) for _ in range(4)]
    dout = rng.standard_normal((B, S, d))
    out_o, leaves = O.mha_tape(x, x, x, *ws, 8)
    O.backward(out_o, dout)
    layer = L.MultiHeadAttention(d, d, 8, weights=ws)
    # A chain of four GEMMs + softmax accumulates fp32 rounding: the yardstick for "fp32-faithful" is the
    # reference algorithm itself evaluated in float32 (NumPy/OpenBLAS sgemm) against its float64 result.
    x32, ws32 = x.astype(np.float32), [w.astype(np.float32) for w in ws]
    out32 = O.mha_forward(x32, x32, x32, *ws32, 8)
    fp32_err = float(np.abs(out32.astype(np.float64) - out_o.val).max())
    atol = max(2e-6, 3.0 * fp32_err / max(1.0, float(np.abs(out_o.val).max())))
    assert atol < 2e-5, "float32 yardstick unexpectedly large: %g" % atol
    assert_close(layer.forward(x, x, x), out_o.val, rtol=2e-5, atol=atol, what="C3 out")
    dq, dk, dv = layer.backward(dout)
    for got, name in zip(layer.grads, ("wq", "wk", "wv", "wo")):
        assert_close(got, leaves[name].grad, rtol=2e-5, atol=atol, what="C3 g" + name)
    assert_close(np.asarray(dq) + np.asarray(dk) + np.asarray(dv),
                 leaves["query"].grad + leaves["key"].grad + leaves["value"].grad, rtol=2e-5, atol=2 * atol,
                 what="C3 gx (sum of three input gradients: three more GEMMs in the chain)")


def test_optimizers_golden(backend):
    L = _L()
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200.utils.optimizer import Adam
    g = golden("optim")
    for name, opt in (("adam", Adam(lr=0.01)), ("adam_layers", L.Adam(lr=0.01)), ("sgd", L.SGD(lr=0.1))):
        params = [pkg.asarray(g["p0_%d" % i]) for i in range(4)]
        ids = [id(p) for p in params]
        for step in range(3):
            opt.update(params, [pkg.asarray(g["g%d_%d" % (step, i)]) for i in range(4)])
        assert ids == [id(p) for p in params]                       # updated in place
        for i, p in enumerate(params):
            assert_close(p, g["%s_p%d" % (name, i)], what="%s p%d" % (name, i))
    for tag, mx in (("clip", 1.0), ("noclip", 1e6)):
        gr = [pkg.asarray(g["g0_%d" % i]) for i in range(4)]
        L.clip_grads(gr, mx)
        for i, a in enumerate(gr):
            assert_close(a, g["%s_%d" % (tag, i)], what="%s %d" % (tag, i))
    with pytest.raises(TypeError):                                   # host arrays are rejected loudly, no fallback
        Adam().update([np.zeros(3)], [np.zeros(3)])


def test_remove_duplicate_and_trainer_contract(backend):
    """Trainer.fit -> model.forward/backward -> remove_duplicate -> optimizer.update (utils/Train.py:15-64)."""
    L = _L()
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200.utils.Train import Trainer
    from auto_differentiation_b200.utils.optimizer import Adam
    from auto_differentiation_b200.utils.useful_func import remove_duplicate
    w = pkg.asarray(np.arange(12, dtype=np.float64).reshape(3, 4))
    g1, g2 = pkg.asarray(np.ones((3, 4))), pkg.asarray(np.full((4, 3), 2.0))
    params, grads = remove_duplicate([w, w.T], [g1, g2])              # transposed weight tying (layers.py:711)
    assert len(params) == 1 and params[0] is w
    assert_close(grads[0], np.full((3, 4), 3.0), what="tied grad")
    params, grads = remove_duplicate([w, w], [pkg.asarray(np.ones((3, 4))), pkg.asarray(np.ones((3, 4)))])
    assert len(params) == 1
    assert_close(grads[0], np.full((3, 4), 2.0), what="shared grad")

    class Model:
        def __init__(self):
            rng = np.random.default_rng(0)
            self.aff = L.TimeAffine(rng.standard_normal((8, 5)) * 0.1, np.zeros(5))
            self.loss = L.TimeSoftmaxWithLoss()
            self.weights, self.grads = self.aff.weights, self.aff.grads

        def forward(self, x, t, q):
            return self.loss.forward(self.aff.forward(x), t)

        def backward(self):
            return self.aff.backward(self.loss.backward())

    rng = np.random.default_rng(1)
    x = rng.standard_normal((64, 3, 8)); t = (x.sum(axis=2) > 0).astype(np.int64) * 2
    tr = Trainer(Model(), Adam(lr=0.05))
    tr.fit(x, t, np.zeros(64), max_epoch=6, batch_size=16, eval_interval=1, verbose=False)
    assert tr.loss_list[-1] < tr.loss_list[0] * 0.7


def test_repeat_and_sum_helpers(backend):
    """`Repeat` (layers.py:113-121) and `Sum` (:124-129): np.repeat along rows / columns, column sums and their backward."""
    L = _L()
    rng = np.random.default_rng(8)
    b = rng.standard_normal((2, 5))
    np.testing.assert_array_equal(np.asarray(L.Repeat(b, 3, "Y").repeated), np.repeat(b.astype(np.float32), 3, axis=0))
    np.testing.assert_array_equal(np.asarray(L.Repeat(b, 4, "x").repeated), np.repeat(b.astype(np.float32), 4, axis=1))
    with pytest.raises(ValueError):
        L.Repeat(b, 2, "z")
    x = rng.standard_normal((7, 5))
    sm = L.Sum()
    assert_close(sm.forward(x), np.sum(x, axis=0, keepdims=True), what="Sum fwd")
    dy = rng.standard_normal((1, 5))
    np.testing.assert_array_equal(np.asarray(sm.backward(dy, 7)), np.repeat(dy.astype(np.float32), 7, axis=0))
