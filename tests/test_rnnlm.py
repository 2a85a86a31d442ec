"""Rnnlm / TimeDropout / RnnlmTrainer (reference layers.py:678-818): two stacked stateful LSTMs
with dropout and the output projection tied to the embedding (`affine_W = embed_W.T`), trained through
remove_duplicate -> clip_grads -> SGD.  Pinned to a golden run of the UNMODIFIED reference
(oracle/make_golden.py gen_rnnlm -> tests/golden/rnnlm.npz); the dropout masks are reproduced exactly by drawing
from the same seeded host stream (`reference_rng=True`)."""
import numpy as np
import pytest

from conftest import assert_close, golden
from oracle import np_oracle as O


def _merged_weights(g):
    """The reference's 9 params (w7 = the tied transpose view) -> the 8 that remove_duplicate keeps."""
    ws = [g["w%d" % i] for i in range(int(g["n"]))]
    assert np.array_equal(ws[7], ws[0].T)
    return ws[:7] + [ws[8]]


def _shapes(g):
    V, H = (int(v) for v in g["dims"])
    N, T = g["xs"].shape[1:]
    return [(N, T, H)] * 3


def test_oracle_rnnlm_matches_reference_golden():
    g = golden("rnnlm")
    ws = [w.copy() for w in _merged_weights(g)]
    np.random.seed(int(g["mask_seed"]))
    opt, state, losses = O.SGD(float(g["lr"])), None, []
    for it in range(g["xs"].shape[0]):
        masks = O.dropout_masks(_shapes(g))
        loss, grads, state, _ = O.rnnlm_step(g["xs"][it], g["ts"][it], ws, masks, state)
        if it == 0:
            assert int(g["n_merged"]) == len(grads)
            for i, gr in enumerate(grads):
                # the reference holds fp32 weights and fp32 LSTM buffers (layers.py:705-713, :566-571)
                assert_close(gr, g["g%d" % i], rtol=2e-5, atol=2e-6, what="merged grad %d" % i)
        O.clip_grads(grads, float(g["max_grad"]))
        opt.update(ws, grads)
        losses.append(loss)
    assert np.allclose(losses, g["losses"], rtol=1e-6, atol=1e-7)
    for i, w in enumerate(ws):
        assert_close(w, g["p%d" % i], rtol=2e-5, atol=2e-6, what="param %d after 2 steps" % i)
    loss_s, _, _, _ = O.rnnlm_step(g["xs"][0], g["ts"][0], ws, None, state)
    assert abs(loss_s - float(g["eval_loss_stateful"])) < 2e-6
    loss_r, _, st, _ = O.rnnlm_step(g["xs"][0], g["ts"][0], ws, None, None)
    assert abs(loss_r - float(g["eval_loss_reset"])) < 2e-6
    _, _, _, score = O.rnnlm_step(g["xs"][1], g["ts"][1], ws, None, st)
    assert_close(score, g["eval_score"], rtol=2e-5, atol=2e-6, what="eval score")


def test_oracle_get_batch_matches_reference_golden():
    g = golden("rnnlm")
    c = g["corpus"]
    bx1, bt1, idx = O.rnnlm_get_batch(c[:-1], c[1:], 4, 5, 0)
    bx2, bt2, _ = O.rnnlm_get_batch(c[:-1], c[1:], 4, 5, idx)
    for got, key in ((bx1, "bx1"), (bt1, "bt1"), (bx2, "bx2"), (bt2, "bt2")):
        np.testing.assert_array_equal(got, g[key])


def _model(g, **kw):
    from auto_differentiation_b200.layers import layers as Ls
    V, H = (int(v) for v in g["dims"])
    return Ls.Rnnlm(V, H, H, weights=[w.copy() for w in _merged_weights(g)], **kw)


def test_rnnlm_training_steps_golden(backend):
    """The trainer's inner loop, two iterations with dropout on, against the reference run."""
    from auto_differentiation_b200.layers import layers as Ls
    g = golden("rnnlm")
    m = _model(g, reference_rng=True)
    assert len(m.params) == 9 and m.params[7].shape == m.params[0].shape[::-1]
    assert m.params[7].ptr == m.params[0].ptr                      # one buffer, two views
    opt = Ls.SGD(float(g["lr"]))
    np.random.seed(int(g["mask_seed"]))
    losses = []
    for it in range(g["xs"].shape[0]):
        loss = m.forward(g["xs"][it], g["ts"][it])
        m.backward()
        params, grads = Ls.remove_duplicate(m.params, m.grads)
        assert len(params) == int(g["n_merged"])
        if it == 0:
            for i, gr in enumerate(grads):
                assert_close(np.asarray(gr), g["g%d" % i], rtol=2e-5, atol=2e-6, what="merged grad %d" % i)
        Ls.clip_grads(grads, float(g["max_grad"]))
        opt.update(params, grads)
        losses.append(float(loss))
    assert np.allclose(losses, g["losses"], rtol=2e-5, atol=2e-6), (losses, g["losses"])
    params, _ = Ls.remove_duplicate(m.params, m.grads)
    for i, p in enumerate(params):
        assert_close(np.asarray(p), g["p%d" % i], rtol=5e-5, atol=5e-6, what="param %d after 2 steps" % i)
    # the tied view follows the embedding update
    np.testing.assert_array_equal(np.asarray(m.params[7]), np.asarray(m.params[0]).T)
    # eval mode: carried state, then reset
    m.train_mode(False)
    assert abs(float(m.forward(g["xs"][0], g["ts"][0])) - float(g["eval_loss_stateful"])) < 5e-5
    m.reset_state()
    assert abs(float(m.forward(g["xs"][0], g["ts"][0])) - float(g["eval_loss_reset"])) < 5e-5
    assert_close(np.asarray(m.predict(g["xs"][1])), g["eval_score"], rtol=5e-5, atol=5e-6, what="eval score")


def test_time_dropout_device_rng(backend):
    """Device-generated masks: values in {0, 1/(1-ratio)}, keep rate near 1-ratio, backward = dout * mask, eval = identity."""
    from auto_differentiation_b200.layers import layers as Ls
    from auto_differentiation_b200 import device as D
    rng = np.random.default_rng(0)
    x = rng.standard_normal((8, 16, 257)).astype(np.float32)       # odd size: exercises the scalar tail
    drop = Ls.TimeDropout(0.25)
    y = np.asarray(drop.forward(x))
    mask = np.asarray(drop.mask)
    scale = np.float32(1 / (1.0 - 0.25))
    assert set(np.unique(mask)).issubset({np.float32(0.0), scale})
    assert abs((mask > 0).mean() - 0.75) < 0.02
    np.testing.assert_array_equal(y, x * mask)
    dout = rng.standard_normal(x.shape).astype(np.float32)
    np.testing.assert_array_equal(np.asarray(drop.backward(dout)), dout * mask)
    drop.train_flg = False
    np.testing.assert_array_equal(np.asarray(drop.forward(x)), x)
    # strided (transposed) input, as TimeLSTM hands it over
    xt = D.asarray(x).transpose(1, 0, 2)
    drop.train_flg = True
    yt = np.asarray(drop.forward(xt))
    np.testing.assert_array_equal(yt, x.transpose(1, 0, 2) * np.asarray(drop.mask))


def test_rnnlm_trainer_fit(backend, tmp_path):
    """fit(): corpus walk identical to the reference's, perplexity history equal to the same loop on the oracle
    (dropout off so that the two sides need no shared RNG), save() writes the de-duplicated parameters."""
    from auto_differentiation_b200.layers import layers as Ls
    g = golden("rnnlm")
    V, H = (int(v) for v in g["dims"])
    corpus = (np.arange(241) * 7 + 3) % V
    xs, ts = corpus[:-1], corpus[1:]
    m = _model(g)
    m.train_mode(False)
    tr = Ls.RnnlmTrainer(m, Ls.SGD(0.5))
    tr.time_idx = 0
    c = g["corpus"]
    bx1, bt1 = tr.get_batch(c[:-1], c[1:], 4, 5)
    bx2, bt2 = tr.get_batch(c[:-1], c[1:], 4, 5)
    for got, key in ((bx1, "bx1"), (bt1, "bt1"), (bx2, "bx2"), (bt2, "bt2")):
        np.testing.assert_array_equal(got, g[key])
        assert got.dtype == g[key].dtype
    ppl = tr.fit(xs, ts, max_epoch=2, batch_size=4, time_size=6, max_grad=0.25, eval_interval=3)
    # the same loop on the oracle
    ws = [w.copy() for w in _merged_weights(g)]
    opt, state, idx, ref_ppl, tot, cnt = O.SGD(0.5), None, 0, [], 0.0, 0
    max_iters = len(xs) // (4 * 6)
    for _ in range(2):
        for it in range(max_iters):
            bx, bt, idx = O.rnnlm_get_batch(xs, ts, 4, 6, idx)
            loss, grads, state, _ = O.rnnlm_step(bx, bt, ws, None, state)
            O.clip_grads(grads, 0.25)
            opt.update(ws, grads)
            tot, cnt = tot + loss, cnt + 1
            if not it % 3:
                ref_ppl.append(float(np.exp(tot / cnt)))
                tot, cnt = 0.0, 0
    assert len(ppl) == len(ref_ppl) and np.allclose(ppl, ref_ppl, rtol=1e-4), (ppl, ref_ppl)
    assert tr.current_epoch == 2
    tr.save(str(tmp_path / "weights"))
    for i, w in enumerate(ws):
        saved = np.load(str(tmp_path / ("weights%d.npy" % i)))
        assert_close(saved, w, rtol=2e-4, atol=2e-5, what="saved param %d" % i)
