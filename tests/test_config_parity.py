"""Parity ON THE PATHS THE BENCH NUMBERS COME FROM ("next round" item 1): the config-size models through
the same kernels / dispatch branches / CUDA-graph replays that `bench.py` times, against `oracle/np_oracle.py`.

  (a) config 4 at its widths (V=1000, D=256, H=1024; T=16; N in {4, 64}), tanh RNN and LSTM: eager AND replayed from
      a CUDA graph (`GraphedStep`), each against the oracle and against each other; the per-step reference loop
      (`persistent=False`) and both recurrent GEMM dispatch branches are pinned explicitly.
  (b) config 3 at S=512 (B=2, d=512, 8 heads, d_k=64 -> the fused tcgen05 attention passes) vs the tape twin.
  (c) config 5 model at d_model=512 / 8 heads with ragged micro-batches vs `O.c5_model_step`.
  (d) NCCL data-parallel gradients of W ranks == single-rank gradients of the same global batch (needs >= 2 GPUs;
      the same check also runs inside `bench.py --gpus N` as the `dp_parity` key of its JSON line).

Tolerances (stated per assertion, achieved errors are printed by the ledger in conftest.py): single kernels rtol
1e-5 / atol 1e-6 of the tensor scale; multi-GEMM chains 2e-5 / 2e-6, with the absolute part tied to a float32
NumPy evaluation of the same reference algorithm for the attention graph (capped at 2e-5)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_close
from oracle import np_oracle as O


def _params_np(model):
    return [np.asarray(w, dtype=np.float64) for w in model.weights]


def _rnnlm_oracle(cell, params, xs, ts):
    E, Wx, Wh, b, Wa, ba = params
    emb = O.embedding_forward(E, xs)
    if cell == "lstm":
        hs, caches, _, _ = O.time_lstm_forward(emb, Wx, Wh, b)
    else:
        hs, hprevs = O.time_rnn_forward(emb, Wx, Wh, b)
    score = O.time_affine_forward(hs, Wa, ba)
    loss, cache = O.time_softmax_with_loss_forward(score, ts)
    dscore = O.time_softmax_with_loss_backward(cache)
    dhs, dWa, dba = O.time_affine_backward(hs, Wa, dscore)
    if cell == "lstm":
        demb, (gWx, gWh, gb), _ = O.time_lstm_backward(caches, Wx, Wh, dhs)
    else:
        demb, (gWx, gWh, gb), _ = O.time_rnn_backward(emb, hs, hprevs, Wx, Wh, dhs)
    gE = O.time_embedding_backward(E.shape, xs, demb)
    return loss, (gE, gWx, gWh, gb, dWa, dba)


_NAMES = ("E", "Wx", "Wh", "b", "Wa", "ba")


@pytest.mark.gpu
@pytest.mark.parametrize("n_batch", [4, 64])
@pytest.mark.parametrize("cell", ["rnn", "lstm"])
def test_config4_widths_eager_and_graph_vs_oracle(cell, n_batch):
    """RnnLM(hidden=1024, wordvec=256, vocab=1000), T=16: loss + all six gradients vs the oracle, eagerly and replayed
    from a CUDA graph; and graph == eager to the same tolerance (the two may take different kernels)."""
    import torch
    from auto_differentiation_b200 import device as D
    from auto_differentiation_b200.graph import GraphedStep
    from auto_differentiation_b200.models.model import RnnLM
    rng = np.random.default_rng(1238)
    V, T = 1000, 16
    xs = rng.integers(0, V, (n_batch, T))
    ts = np.roll(xs, -1, axis=1)
    ts[0, 3] = -1                                             # one ignored position (TimeSoftmaxWithLoss mask)
    m = RnnLM(vocab=V, wordvec=256, hidden=1024, cell=cell)
    loss_o, grads_o = _rnnlm_oracle(cell, _params_np(m), xs, ts)

    loss = m.forward(xs, ts)
    m.backward()
    assert abs(float(loss) - loss_o) <= 2e-5 * max(1.0, abs(loss_o)), (float(loss), loss_o)
    eager = [np.asarray(g, dtype=np.float64) for g in m.grads]
    for got, ref, name in zip(eager, grads_o, _NAMES):
        assert_close(got, ref, rtol=2e-5, atol=2e-6, what="C4 %s N=%d eager g%s" % (cell, n_batch, name))

    m2 = RnnLM(vocab=V, wordvec=256, hidden=1024, cell=cell)
    xs_d, ts_d = D.as_index(xs), D.as_index(ts)

    def step():
        out = m2.forward(xs_d, ts_d)
        m2.backward()
        return out
    g = GraphedStep(step, warmup=2)
    for _ in range(2):
        loss_g = g()
    torch.cuda.synchronize()
    assert abs(float(loss_g) - loss_o) <= 2e-5 * max(1.0, abs(loss_o)), (float(loss_g), loss_o)
    for got, ref, e, name in zip(m2.grads, grads_o, eager, _NAMES):
        got = np.asarray(got, dtype=np.float64)
        assert_close(got, ref, rtol=2e-5, atol=2e-6, what="C4 %s N=%d graph g%s" % (cell, n_batch, name))
        assert_close(got, e, rtol=2e-5, atol=2e-6, what="C4 %s N=%d graph-vs-eager g%s" % (cell, n_batch, name))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["device_time", "single_launch"])
def test_recurrent_gemm_dispatch_branches_pinned(mode):
    """The M <= 64 recurrent-step GEMM has two implementations (tcgen05 64-column tiles + split-K, and the exact-fp32
    single-launch cluster kernel).  `auto` picks by capture state; here each branch is pinned through
    `b200ad_gemm_set_recurrent_dispatch` and checked against fp64 at the config-4 step shape, eagerly."""
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200 import _lib as L, device as D
    rng = np.random.default_rng(5)
    h = rng.standard_normal((64, 1024))
    w = rng.standard_normal((1024, 1024)) / 32.0
    pre = rng.standard_normal((64, 1024))
    z = pre + h @ w
    lib = L.load()
    prev = lib.b200ad_gemm_set_recurrent_dispatch({"device_time": 1, "single_launch": 2}[mode])
    try:
        # the product itself, to rtol 1e-5 / atol 1e-6 of ITS scale (max |z| ~ 6): the tcgen05 branch sits at ~0.6e-6 of
        # that scale (fp32 accumulation inside the tensor core truncates; chains are folded every 128 K-elements)
        lin = pkg.asarray(pre).copy()
        D.matmul(pkg.asarray(h), pkg.asarray(w), out=lin, beta=1.0)
        assert_close(lin, z, rtol=1e-5, atol=1e-6, what="recurrent step pre-activation (%s)" % mode)
        # with the fused tanh: tanh is 1-Lipschitz, so the same ABSOLUTE error bound carries over (atol stated in units
        # of the pre-activation scale, not of tanh's unit range)
        out = pkg.asarray(pre).copy()
        D.matmul(pkg.asarray(h), pkg.asarray(w), out=out, beta=1.0, act=L.ACT_TANH)
        assert_close(out, np.tanh(z), rtol=1e-5, atol=1e-6 * float(np.abs(z).max()), what="recurrent step (%s)" % mode)
        outT = D.matmul(pkg.asarray(h), pkg.asarray(w), trans_b=True)
        assert_close(outT, h @ w.T, rtol=1e-5, atol=1e-6, what="BPTT step (%s)" % mode)
    finally:
        lib.b200ad_gemm_set_recurrent_dispatch(prev)


@pytest.mark.gpu
def test_config3_layer_s512_vs_tape_twin():
    """MultiHeadAttention at B=2, S=512, d=512, 8 heads (d_k = 64: the fused attention passes + gen3 projections) vs
    the oracle's tape twin of Test/MultiHeadAttentionTest.py:57-91 — output and all seven gradients."""
    from auto_differentiation_b200.layers import layers as Ls
    rng = np.random.default_rng(1237)
    B, S, d = 2, 512, 512
    q, k, v = (rng.standard_normal((B, S, d)) for _ in range(3))
    ws = [rng.standard_normal((d, d)) / np.sqrt(d) for _ in range(4)]
    dout = rng.standard_normal((B, S, d))
    out_o, leaves = O.mha_tape(q, k, v, *ws, 8)
    O.backward(out_o, dout)
    # yardstick for "fp32-faithful" on a 4-GEMM + softmax chain: the reference algorithm itself in float32 NumPy
    out32 = O.mha_forward(q.astype(np.float32), k.astype(np.float32), v.astype(np.float32),
                          *[w.astype(np.float32) for w in ws], 8)
    fp32_err = float(np.abs(out32.astype(np.float64) - out_o.val).max())
    atol = max(2e-6, 3.0 * fp32_err / max(1.0, float(np.abs(out_o.val).max())))
    assert atol < 2e-5, "float32 yardstick unexpectedly large: %g" % atol
    layer = Ls.MultiHeadAttention(d, d, 8, weights=ws)
    out = layer.forward(q, k, v)
    assert layer.cache[6] is None, "fused attention path not taken at d_k = 64"
    assert_close(out, out_o.val, rtol=2e-5, atol=atol, what="C3 S=512 out")
    dq, dk, dv = layer.backward(dout)
    # gradients sit at the end of a 5-7 GEMM chain (forward 3-4 deep, backward 2-4 more) against 4 for the output the
    # yardstick was taken on: their absolute part is stated as twice the output's
    for got, name in zip(layer.grads, ("wq", "wk", "wv", "wo")):
        assert_close(got, leaves[name].grad, rtol=2e-5, atol=2 * atol, what="C3 S=512 g" + name)
    for got, name in ((dq, "query"), (dk, "key"), (dv, "value")):
        assert_close(got, leaves[name].grad, rtol=2e-5, atol=2 * atol, what="C3 S=512 g" + name)


@pytest.mark.gpu
def test_config5_model_widths_ragged_microbatches():
    """AttentionClassifier(d_model=512, heads=8, hidden=256, classes=10) — the bench model, fused d_k = 64 attention —
    B=5 sequences of S=512 in micro-batches of 2 (2+2+1) with on-device gradient accumulation vs `O.c5_model_step`."""
    from auto_differentiation_b200.models.model import AttentionClassifier
    from auto_differentiation_b200.utils.optimizer import Adam
    rng = np.random.default_rng(1239)
    B, S, d = 5, 512, 512
    x = rng.standard_normal((B, S, d))
    t = rng.integers(0, 10, B)
    m = AttentionClassifier(d_model=d, heads=8, hidden=256, classes=10, micro_batch=2)
    params = _params_np(m)
    loss_o, grads_o, _ = O.c5_model_step(x, t, params, heads=8)
    loss = m.forward(x.astype(np.float32), t)
    m.backward()
    assert m.mha.cache[6] is None, "fused attention path not taken"
    assert abs(float(loss) - loss_o) <= 2e-5 * max(1.0, abs(loss_o)), (float(loss), loss_o)
    for got, ref, name in zip(m.grads, grads_o, ("wq", "wk", "wv", "wo", "W1", "b1", "W2", "b2")):
        assert_close(got, ref.reshape(got.shape), rtol=2e-5, atol=2e-6, what="C5 widths g" + name)
    opt, ropt = Adam(lr=1e-3), O.Adam(lr=1e-3)
    opt.update(m.weights, m.grads)
    ropt.update(params, [g.reshape(p.shape) for g, p in zip(grads_o, params)])
    for got, ref, name in zip(m.weights, params, ("wq", "wk", "wv", "wo", "W1", "b1", "W2", "b2")):
        assert_close(got, ref, rtol=1e-5, atol=1e-6, what="C5 widths Adam " + name)


@pytest.mark.gpu
def test_nccl_data_parallel_gradients_equal_single_rank():
    """SURVEY §8(e) parity: reduced gradients from W NCCL ranks == one rank on the same global batch (rtol 1e-5)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (bench.py --gpus N runs the same check as `dp_parity`)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "dp_parity_nccl.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "dp_parity ok" in res.stdout, res.stdout[-2000:]
