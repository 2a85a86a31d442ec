"""The persistent tanh-RNN time-loop kernel (csrc/rnn_seq.cu): forward and BPTT in one launch each, against fp64 NumPy
statements of the reference loops (TimeRNN.forward layers/layers.py:203-207; backward :210-225 as intended),
against the oracle through the layer, against the per-step loop it replaces, bit-reproducibility, output guards and
CUDA-graph replay.  Tolerance: the state is a chain of T dependent 3xTF32 products, each error passing through
tanh' <= 1 and Wh (spectral norm ~2 for the N(0,1)/sqrt(H) init): rtol 2e-5 / atol 2e-6 of the tensor scale per
output, like every multi-GEMM graph in this suite (achieved margins are printed by the ledger)."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import np_oracle as O


def _fwd_ref(h0, pre, wh):
    T = pre.shape[0]
    hs = np.empty((T + 1,) + h0.shape)
    hs[0] = h0
    for t in range(T):
        hs[t + 1] = np.tanh(pre[t] + hs[t] @ wh)
    return hs


def _bwd_ref(dhs, hs, wh):
    N, T, H = dhs.shape
    dA = np.empty((T, N, H))
    dh = np.zeros((N, H))
    for t in reversed(range(T)):
        dA[t] = (dhs[:, t, :] + dh) * (1.0 - hs[t + 1] ** 2)
        dh = dA[t] @ wh.T
    return dA, dh


@pytest.mark.parametrize("H,N,T", [(256, 5, 6), (512, 33, 4), (1024, 64, 16), (1024, 1, 3), (1024, 32, 5), (256, 64, 1)])
def test_rnn_seq_kernels_vs_fp64(backend, H, N, T):
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200 import device as D
    if not D.rnn_seq_supported(N, H):
        pytest.skip("persistent recurrence not available for N=%d H=%d on this device" % (N, H))
    rng = np.random.default_rng(H + N + T)
    wh = (rng.standard_normal((H, H)) / np.sqrt(H)).astype(np.float32)
    h0 = (rng.standard_normal((N, H)) * 0.5).astype(np.float32)
    pre = rng.standard_normal((T, N, H)).astype(np.float32)
    hs_ref = _fwd_ref(h0.astype(np.float64), pre.astype(np.float64), wh.astype(np.float64))
    hs_ext = pkg.asarray(np.concatenate([h0[None], np.zeros_like(pre)], axis=0))
    D.rnn_tanh_seq_fwd_(hs_ext, pkg.asarray(pre), pkg.asarray(wh))
    assert_close(hs_ext, hs_ref, rtol=2e-5, atol=2e-6, what="rnn_seq fwd H=%d N=%d T=%d" % (H, N, T))
    # BPTT on the fp64 states (so that the backward error is measured alone), upstream gradient as a strided (N,T,H) view
    dhs_full = rng.standard_normal((N, T, H + 8)).astype(np.float32)
    dhs = pkg.asarray(dhs_full)[:, :, :H]
    hs_dev = pkg.asarray(hs_ref)
    dA, dh0 = D.rnn_tanh_seq_bwd(dhs, hs_dev, pkg.asarray(wh))
    hs32 = np.asarray(hs_dev, dtype=np.float64)
    dA_ref, dh0_ref = _bwd_ref(dhs_full[:, :, :H].astype(np.float64), hs32, wh.astype(np.float64))
    assert_close(dA, dA_ref, rtol=2e-5, atol=2e-6, what="rnn_seq bwd dA H=%d N=%d T=%d" % (H, N, T))
    assert_close(dh0, dh0_ref, rtol=2e-5, atol=2e-6, what="rnn_seq bwd dh0 H=%d N=%d T=%d" % (H, N, T))


def test_time_rnn_layer_persistent_equals_loop_and_oracle(backend):
    """TimeRNN at a width the persistent kernel takes (H = 256): forward / backward vs the oracle, and vs the per-step
    loop (`set_persistent_recurrence(False)`) it replaces."""
    from auto_differentiation_b200 import device as D
    from auto_differentiation_b200.layers import layers as Ls
    rng = np.random.default_rng(12)
    N, T, Dm, H = 6, 7, 24, 256
    xs = rng.standard_normal((N, T, Dm))
    Wx, Wh, b = rng.standard_normal((Dm, H)) / np.sqrt(Dm), rng.standard_normal((H, H)) / np.sqrt(H), rng.standard_normal(H) * 0.1
    dhs = rng.standard_normal((N, T, H))
    hs_o, hprevs = O.time_rnn_forward(xs, Wx, Wh, b)
    dxs_o, (gWx, gWh, gb), _ = O.time_rnn_backward(xs, hs_o, hprevs, Wx, Wh, dhs)
    outs = {}
    for persistent in (True, False):
        D.set_persistent_recurrence(persistent)
        try:
            layer = Ls.TimeRNN(Wx, Wh, b)
            hs = layer.forward(xs)
            dxs = layer.backward(dhs)
            outs[persistent] = [np.asarray(a, dtype=np.float64) for a in (hs, dxs, *layer.grads, layer.dh)]
        finally:
            D.set_persistent_recurrence(True)
        for got, ref, name in zip(outs[persistent], (hs_o, dxs_o, gWx, gWh, gb), ("hs", "dxs", "dWx", "dWh", "db")):
            assert_close(got, ref, rtol=2e-5, atol=2e-6, what="TimeRNN persistent=%s %s" % (persistent, name))
    for a, b_, name in zip(outs[True], outs[False], ("hs", "dxs", "dWx", "dWh", "db", "dh0")):
        assert_close(a, b_, rtol=2e-5, atol=2e-6, what="TimeRNN persistent vs loop " + name)


@pytest.mark.gpu
def test_rnn_seq_bit_reproducible_guards_and_graph_replay():
    """Run-to-run bit equality (fixed reduction order), nothing written outside the output windows, and a CUDA-graph
    replay of forward + backward equal to the eager launches bit for bit (same kernel either way)."""
    import torch
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200 import device as D
    from auto_differentiation_b200.graph import GraphedStep
    H, N, T = 1024, 37, 9
    if not D.rnn_seq_supported(N, H):
        pytest.skip("persistent recurrence not available on this device")
    rng = np.random.default_rng(3)
    wh = pkg.asarray(rng.standard_normal((H, H)) / np.sqrt(H))
    init = np.concatenate([rng.standard_normal((1, N, H)) * 0.5, rng.standard_normal((T, N, H))], axis=0).astype(np.float32)
    dhs = pkg.asarray(rng.standard_normal((N, T, H)))
    SENT = 12345.0
    guard = pkg.full(((T + 1) * N * H + 2 * 4096,), SENT)
    hs_ext = guard[4096:4096 + (T + 1) * N * H].reshape(T + 1, N, H)

    pre_dev = pkg.asarray(init[1:])

    def run():
        D.copy_into_(pkg.asarray(init), hs_ext)
        D.rnn_tanh_seq_fwd_(hs_ext, pre_dev, wh)
        return D.rnn_tanh_seq_bwd(dhs, hs_ext, wh)
    results = []
    for _ in range(3):
        dA, dh0 = run()
        torch.cuda.synchronize()
        results.append((np.asarray(hs_ext).copy(), np.asarray(dA).copy(), np.asarray(dh0).copy()))
    for r in results[1:]:
        for a, b in zip(results[0], r):
            np.testing.assert_array_equal(a, b)
    g = np.asarray(guard)
    assert np.all(g[:4096] == SENT) and np.all(g[-4096:] == SENT), "rnn_seq wrote outside hs_ext"
    assert np.all(np.isfinite(results[0][0])) and np.all(np.abs(results[0][0][1:]) <= 1.0)
    static_in = pkg.asarray(init)

    def graphed():
        D.copy_into_(static_in, hs_ext)
        D.rnn_tanh_seq_fwd_(hs_ext, pre_dev, wh)
        return D.rnn_tanh_seq_bwd(dhs, hs_ext, wh)
    gs = GraphedStep(graphed, warmup=1)
    for _ in range(2):
        dA_g, dh0_g = gs()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(np.asarray(hs_ext), results[0][0])
    np.testing.assert_array_equal(np.asarray(dA_g), results[0][1])
    np.testing.assert_array_equal(np.asarray(dh0_g), results[0][2])


@pytest.mark.gpu
def test_rnn_seq_two_streams():
    """Two persistent launches issued on different streams (each needs every one of its CTAs resident): the library chains
    them, and both produce exactly what a lone launch produces."""
    import torch
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200 import device as D
    H, N, T = 1024, 64, 24
    if not D.rnn_seq_supported(N, H):
        pytest.skip("persistent recurrence not available on this device")
    rng = np.random.default_rng(8)
    wh = pkg.asarray(rng.standard_normal((H, H)) / np.sqrt(H))
    pres = [pkg.asarray(rng.standard_normal((T, N, H))) for _ in range(2)]
    lone = []
    for p in pres:
        hs = D.zeros((T + 1, N, H))
        D.rnn_tanh_seq_fwd_(hs, p, wh)
        torch.cuda.synchronize()
        lone.append(np.asarray(hs).copy())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [D.zeros((T + 1, N, H)) for _ in range(2)]
    torch.cuda.synchronize()
    for _ in range(3):
        for st, hs, p in zip(streams, outs, pres):
            with torch.cuda.stream(st):
                D.rnn_tanh_seq_fwd_(hs, p, wh)
    torch.cuda.synchronize()
    for got, ref in zip(outs, lone):
        np.testing.assert_array_equal(np.asarray(got), ref)


@pytest.mark.parametrize("N,T,Dm,H", [(16, 6, 24, 64), (64, 5, 32, 256), (33, 3, 16, 128), (1, 1, 8, 256), (2, 7, 12, 512)])
def test_time_lstm_sequence_call_equals_loop_and_oracle(backend, N, T, Dm, H):
    """TimeLSTM with the time loops issued from one C call (csrc/lstm_seq.cu: tcgen05 split-K recurrent product + a cell
    kernel that sums the partials itself) against the oracle (TimeLSTM.forward / backward, layers/layers.py:561-597) and
    against the per-step loop over b200ad_gemm + b200ad_lstm_cell_* it replaces (`set_persistent_recurrence(False)`)."""
    from auto_differentiation_b200 import device as D
    from auto_differentiation_b200.layers import layers as Ls
    if not D.lstm_seq_supported(N, H):
        pytest.skip("LSTM sequence call not available for N=%d H=%d" % (N, H))
    rng = np.random.default_rng(N + T + H)
    xs = rng.standard_normal((N, T, Dm))
    Wx = rng.standard_normal((Dm, 4 * H)) / np.sqrt(Dm)
    Wh = rng.standard_normal((H, 4 * H)) / np.sqrt(H)
    b = rng.standard_normal(4 * H) * 0.1
    dhs = rng.standard_normal((N, T, H))
    hs_o, caches, _, _ = O.time_lstm_forward(xs, Wx, Wh, b)
    dxs_o, (gWx, gWh, gb), dh_o = O.time_lstm_backward(caches, Wx, Wh, dhs)
    outs = {}
    for seq in (True, False):
        D.set_persistent_recurrence(seq)
        try:
            layer = Ls.TimeLSTM(Wx, Wh, b)
            hs = layer.forward(xs)
            dxs = layer.backward(dhs)
            outs[seq] = [np.asarray(a, dtype=np.float64) for a in (hs, dxs, *layer.grads, layer.dh)]
        finally:
            D.set_persistent_recurrence(True)
        for got, ref, name in zip(outs[seq], (hs_o, dxs_o, gWx, gWh, gb, dh_o), ("hs", "dxs", "dWx", "dWh", "db", "dh0")):
            assert_close(got, ref, rtol=2e-5, atol=2e-6, what="TimeLSTM %s (%s)" % (name, "one call" if seq else "per-step loop"))
    for a, c, name in zip(outs[True], outs[False], ("hs", "dxs", "dWx", "dWh", "db", "dh0")):
        assert_close(a, c, rtol=2e-5, atol=2e-6, what="TimeLSTM one call vs per-step loop: " + name)
