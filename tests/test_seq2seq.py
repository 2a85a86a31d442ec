"""AttentionSeq2seq (reference layers.py:1086-1240): encoder / attention-decoder / loss composition,
greedy decoding and the fit loop (global-norm clip + Adam), pinned to a golden run of the UNMODIFIED reference
(oracle/make_golden.py gen_seq2seq -> tests/golden/seq2seq.npz)."""
import numpy as np
import pytest

from conftest import assert_close, golden
from oracle import np_oracle as O


def _weights(g):
    return [g["w%d" % i] for i in range(int(g["n"]))]


def test_oracle_seq2seq_matches_reference_golden():
    g = golden("seq2seq")
    ws = _weights(g)
    loss, grads = O.attention_seq2seq_step(g["xs"], g["ts"], ws)
    assert abs(loss - float(g["loss"])) < 1e-9
    for i, gr in enumerate(grads):
        ref = g["g%d" % i]
        # the reference keeps float32 buffers inside TimeLSTM / the decoder's weights (layers.py:566-571, :1093-1098)
        assert_close(gr.reshape(ref.shape), ref, rtol=1e-5, atol=1e-6, what="grad %d" % i)
    assert O.attention_seq2seq_generate(g["xs"][:1], ws, 3, 6) == [int(v) for v in g["generated"]]


def test_seq2seq_forward_backward_generate_golden(backend):
    from auto_differentiation_b200.layers import layers as Ls
    g = golden("seq2seq")
    V, Dm, H = (int(v) for v in g["dims"])
    m = Ls.AttentionSeq2seq({}, V, Dm, H, weights=_weights(g))
    loss = m.forward(g["xs"], g["ts"])
    m.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-5
    for i, gr in enumerate(m.grads):
        ref = g["g%d" % i]
        assert_close(np.asarray(gr).reshape(ref.shape), ref, rtol=2e-5, atol=2e-6, what="grad %d" % i)
    # greedy decoding: integer outputs, bit-exact
    enc_hs = m.encoder.forward(g["xs"][:1])
    assert m.decoder.generate(enc_hs, 3, 6) == [int(v) for v in g["generated"]]


def test_seq2seq_fit_matches_oracle_loop(backend):
    """Three fit iterations (forward, backward, clip_grad, Adam) vs the same loop on the oracle."""
    from auto_differentiation_b200.layers import layers as Ls
    from auto_differentiation_b200.utils.optimizer import Adam
    g = golden("seq2seq")
    V, Dm, H = (int(v) for v in g["dims"])
    rng = np.random.default_rng(3)
    N, Te, Td, steps = 4, 6, 5, 3
    xs = rng.integers(0, V, (N * steps, Te))
    ts = rng.integers(0, V, (N * steps, Td + 1))
    max_grad = 0.05                                        # small enough that the clip is active
    m = Ls.AttentionSeq2seq({}, V, Dm, H, weights=[w.copy() for w in _weights(g)])
    hist = m.fit(xs, ts, optimizer=Adam(lr=0.01), max_epoch=1, batch_size=N, max_grad=max_grad)
    ws = [w.astype(np.float64).copy() for w in _weights(g)]
    opt = O.Adam(lr=0.01)
    ref_hist = []
    for it in range(steps):
        loss, grads = O.attention_seq2seq_step(xs[it * N:(it + 1) * N], ts[it * N:(it + 1) * N], ws)
        grads = [gr.reshape(w.shape).astype(np.float64) for gr, w in zip(grads, ws)]
        O.seq2seq_clip_grad(grads, max_grad)
        opt.update(ws, grads)
        ref_hist.append(loss)
    assert np.allclose(hist, ref_hist, rtol=2e-5, atol=2e-6), (hist, ref_hist)
    for i, (w, ref) in enumerate(zip(m.weights, ws)):
        assert_close(np.asarray(w).reshape(ref.shape), ref, rtol=5e-5, atol=5e-6, what="weight %d after fit" % i)


# ---- Seq2Seq with the peeky Decoder (reference layers.py:866-958) --------------------------------------------------
def test_oracle_peeky_seq2seq_matches_reference_golden():
    g = golden("peeky_seq2seq")
    ws = _weights(g)
    loss, grads = O.peeky_seq2seq_step(g["xs"], g["ts"], ws)
    assert abs(loss - float(g["loss"])) < 1e-9
    for i, gr in enumerate(grads):
        ref = g["g%d" % i]
        assert_close(gr.reshape(ref.shape), ref, rtol=1e-5, atol=1e-6, what="grad %d" % i)
    assert O.peeky_seq2seq_generate(g["xs"][:1], ws, 4, 5) == [int(v) for v in g["generated"]]


def test_peeky_seq2seq_forward_backward_generate_golden(backend):
    from auto_differentiation_b200.layers import layers as Ls
    g = golden("peeky_seq2seq")
    V, Dm, H = (int(v) for v in g["dims"])
    m = Ls.Seq2Seq({"_": 4}, V, Dm, H, weights=_weights(g))
    loss = m.forward(g["xs"], g["ts"])
    m.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-5
    for i, gr in enumerate(m.grads):
        ref = g["g%d" % i]
        assert_close(np.asarray(gr).reshape(ref.shape), ref, rtol=2e-5, atol=2e-6, what="grad %d" % i)
    assert m.generate(g["xs"][:1], {"_": 4}, 5) == [int(v) for v in g["generated"]]      # integer outputs, bit-exact
    m.reset_state()
    assert m.decoder.layers[1].h is None
