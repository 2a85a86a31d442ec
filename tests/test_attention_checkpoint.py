"""Bahdanau-style attention layers (vs golden + oracle) and checkpoint save/load."""
import os

import numpy as np

from conftest import assert_close, golden
from oracle import np_oracle as O


def test_oracle_attention_matches_reference_golden():
    g = golden("attention")
    np.testing.assert_allclose(O.weight_sum_forward(g["ws_hs"], g["ws_a"]), g["ws_c"], rtol=1e-12, atol=1e-12)
    dhs, da = O.weight_sum_backward(g["ws_hs"], g["ws_a"], g["ws_dc"])
    np.testing.assert_allclose(dhs, g["ws_dhs"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(da, g["ws_da"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.attention_weight_forward(g["ws_hs"], g["aw_h"]), g["aw_w"], rtol=1e-12, atol=1e-12)
    dhs, dh = O.attention_weight_backward(g["ws_hs"], g["aw_h"], g["aw_dw"])
    np.testing.assert_allclose(dhs, g["aw_dhs"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dh, g["aw_dh"], rtol=1e-12, atol=1e-12)
    out, ws = O.time_attention_forward(g["ws_hs"], g["ta_dec"])
    np.testing.assert_allclose(out, g["ta_out"], rtol=1e-12, atol=1e-12)
    de, dd = O.time_attention_backward(g["ws_hs"], g["ta_dec"], ws, g["ta_do"])
    np.testing.assert_allclose(de, g["ta_dhs_enc"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dd, g["ta_dhs_dec"], rtol=1e-12, atol=1e-12)


def test_attention_layers_golden(backend):
    from auto_differentiation_b200.layers import layers as L
    g = golden("attention")
    hs = g["ws_hs"]
    ws = L.WeightSum()
    assert_close(ws.forward(hs, g["ws_a"]), g["ws_c"], what="WeightSum fwd")
    dhs, da = ws.backward(g["ws_dc"])
    assert_close(dhs, g["ws_dhs"], what="WeightSum dhs"); assert_close(da, g["ws_da"], what="WeightSum da")
    aw = L.AttentionWeight()
    assert_close(aw.forward(hs, g["aw_h"]), g["aw_w"], what="AttentionWeight fwd")
    dhs, dh = aw.backward(g["aw_dw"])
    assert_close(dhs, g["aw_dhs"], what="AttentionWeight dhs"); assert_close(dh, g["aw_dh"], what="AttentionWeight dh")
    at = L.Attention()
    assert_close(at.forward(hs, g["aw_h"]), g["at_c"], what="Attention fwd")
    dhs, dh = at.backward(g["at_dc"])
    assert_close(dhs, g["at_dhs"], what="Attention dhs"); assert_close(dh, g["at_dh"], what="Attention dh")
    ta = L.TimeAttention()
    assert_close(ta.forward(hs, g["ta_dec"]), g["ta_out"], what="TimeAttention fwd")
    assert_close(ta.hs_weight, g["ta_weights"], what="TimeAttention weights")
    de, dd = ta.backward(g["ta_do"])
    assert_close(de, g["ta_dhs_enc"], what="TimeAttention dhs_enc"); assert_close(dd, g["ta_dhs_dec"], what="TimeAttention dhs_dec")
    # a larger, tensor-core sized case against the oracle
    rng = np.random.default_rng(1241)
    N, Te, Td, H = 4, 96, 64, 128
    enc, dec, do = rng.standard_normal((N, Te, H)) / 4, rng.standard_normal((N, Td, H)) / 4, rng.standard_normal((N, Td, H))
    out_o, w_o = O.time_attention_forward(enc, dec)
    de_o, dd_o = O.time_attention_backward(enc, dec, w_o, do)
    ta = L.TimeAttention()
    assert_close(ta.forward(enc, dec), out_o, rtol=2e-5, atol=2e-6, what="TimeAttention fwd (large)")
    de, dd = ta.backward(do)
    assert_close(de, de_o, rtol=2e-5, atol=4e-6, what="dhs_enc (large)"); assert_close(dd, dd_o, rtol=2e-5, atol=4e-6, what="dhs_dec (large)")


def test_checkpoint_roundtrip(backend, tmp_path):
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200.utils import checkpoint
    from auto_differentiation_b200.utils.optimizer import Adam
    rng = np.random.default_rng(9)
    shapes = [(5, 7), (7,), (7, 5)]
    params = [pkg.asarray(rng.standard_normal(s)) for s in shapes]
    grads = [pkg.asarray(rng.standard_normal(s)) for s in shapes]
    opt = Adam(lr=0.01)
    opt.update(params, grads); opt.update(params, grads)
    before = [np.asarray(p).copy() for p in params]
    n = checkpoint.save(params, grads, directory=str(tmp_path), optimizer=opt)
    assert n == 3 and os.path.isfile(tmp_path / "weights0.npy")            # the reference's file naming (layers.py:818)
    np.testing.assert_array_equal(np.load(tmp_path / "weights2.npy"), before[2])
    opt.update(params, grads)                                               # diverge, then restore
    after_one_more = [np.asarray(p).copy() for p in params]
    opt2 = Adam(lr=0.01)
    ids = [id(p) for p in params]
    checkpoint.load(params, grads, directory=str(tmp_path), optimizer=opt2)
    assert ids == [id(p) for p in params] and opt2.iter == 2
    for p, b in zip(params, before):
        np.testing.assert_array_equal(np.asarray(p), b)
    opt2.update(params, grads)                                              # resumed run == uninterrupted run
    for p, a in zip(params, after_one_more):
        assert_close(p, a, rtol=1e-6, atol=1e-7, what="resume")


def test_checkpoint_capturable_adam_counter_and_tied_grads(backend, tmp_path):
    """(1) a capturable Adam's real step counter is the DEVICE scalar (graph replays never touch the host
    attribute): save() must read it, load() must write it back into a live optimizer; (2) save()/load() must not fold
    the gradients of tied weights (they only select parameters)."""
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200.utils import checkpoint
    from auto_differentiation_b200.utils.optimizer import Adam
    rng = np.random.default_rng(10)
    params = [pkg.asarray(rng.standard_normal((6, 4))), pkg.asarray(rng.standard_normal((4,)))]
    grads = [pkg.asarray(rng.standard_normal((6, 4))), pkg.asarray(rng.standard_normal((4,)))]
    opt = Adam(lr=0.01, capturable=True)
    for _ in range(5):
        opt.update(params, grads)
    opt.iter = 2                                   # what the host attribute looks like after graph REPLAYS: stale
    assert opt.step_count() == 5
    checkpoint.save(params, grads, directory=str(tmp_path), optimizer=opt)
    assert int(np.load(tmp_path / "adam_state.npz")["iter"]) == 5
    ref = Adam(lr=0.01, capturable=True)           # uninterrupted run: 6 updates
    p_ref = [pkg.asarray(np.asarray(p)) for p in params]
    # rebuild the same state by restoring into a fresh optimizer and stepping once; compare with stepping the live one
    m_ids = [id(m) for m in opt.m]
    opt.update(params, grads)
    expected = [np.asarray(p).copy() for p in params]
    for _ in range(3):
        opt.update(params, grads)                  # diverge the live optimizer (counter 9, m/v moved on) ...
    checkpoint.load(params, grads, directory=str(tmp_path), optimizer=opt)     # ... and restore it IN PLACE
    assert opt.step_count() == 5 and m_ids == [id(m) for m in opt.m]
    opt.update(params, grads)
    for p, e in zip(params, expected):
        assert_close(p, e, rtol=1e-6, atol=1e-7, what="live capturable resume")
    fresh = Adam(lr=0.01, capturable=True)
    checkpoint.load(p_ref, None, directory=str(tmp_path), optimizer=fresh)
    fresh.update(p_ref, grads)
    assert fresh.step_count() == 6
    for p, e in zip(p_ref, expected):
        assert_close(p, e, rtol=1e-6, atol=1e-7, what="fresh capturable resume")
    # tied weights: W and W.T share storage; checkpointing between backward() and update() leaves both gradients alone
    w = pkg.asarray(rng.standard_normal((3, 5)))
    g1, g2 = pkg.asarray(np.ones((3, 5))), pkg.asarray(np.full((5, 3), 2.0))
    d2 = tmp_path / "tied"
    assert checkpoint.save([w, w.T], [g1, g2], directory=str(d2)) == 1
    checkpoint.load([w, w.T], [g1, g2], directory=str(d2))
    np.testing.assert_array_equal(np.asarray(g1), np.ones((3, 5)))
    np.testing.assert_array_equal(np.asarray(g2), np.full((5, 3), 2.0))
