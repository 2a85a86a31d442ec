"""Config models (1, 4, 5) vs the oracle, micro-batch accumulation, and the data-parallel path
(world_size-2 gloo on CPU: sharding + bucketed all-reduce == single-process gradients)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import assert_close
from oracle import np_oracle as O


def _params_np(model):
    return [np.asarray(w, dtype=np.float64) for w in model.weights]


def test_attention_classifier_matches_oracle_and_microbatching(backend):
    from auto_differentiation_b200.models.model import AttentionClassifier
    rng = np.random.default_rng(1239)
    B, S, d = 7, 12, 32
    x = rng.standard_normal((B, S, d)); t = rng.integers(0, 5, B)
    ref_model = AttentionClassifier(d_model=d, heads=4, hidden=16, classes=5, micro_batch=B)
    params = _params_np(ref_model)
    loss_o, grads_o, _ = O.c5_model_step(x, t, params, heads=4)
    for mb in (B, 3):                                   # one chunk, and ragged micro-batches (3+3+1)
        m = AttentionClassifier(d_model=d, heads=4, hidden=16, classes=5, micro_batch=mb)
        loss = m.forward(x, t)
        m.backward()
        assert abs(float(loss) - loss_o) < 2e-5, (float(loss), loss_o)
        for got, ref, name in zip(m.grads, grads_o, ("wq", "wk", "wv", "wo", "W1", "b1", "W2", "b2")):
            assert_close(got, ref.reshape(got.shape), rtol=2e-5, atol=2e-6, what="mb=%d g%s" % (mb, name))
    # one Adam step == the oracle's Adam on the same gradients
    from auto_differentiation_b200.utils.optimizer import Adam
    opt, ropt = Adam(lr=1e-3), O.Adam(lr=1e-3)
    opt.update(m.weights, m.grads)
    rp = [p.copy() for p in params]
    ropt.update(rp, [g.reshape(p.shape) for g, p in zip(grads_o, rp)])
    for got, ref in zip(m.weights, rp):
        assert_close(got, ref, rtol=2e-5, atol=2e-6, what="adam step")


def test_tape_mlp_trains_and_matches_oracle(backend):
    from auto_differentiation_b200.models.model import TapeMLP
    from auto_differentiation_b200.utils.Train import Trainer
    from auto_differentiation_b200.utils.optimizer import Adam
    rng = np.random.default_rng(1235)
    x = rng.standard_normal((64, 784)); t = rng.integers(0, 10, 64)
    m = TapeMLP()
    W1, b1, W2, b2 = _params_np(m)
    loss_o, y_o, grads_o, _ = O.mlp_tape_step(x, t, W1, b1, W2, b2)
    loss = m.forward(x, t)
    m.backward()
    assert abs(float(loss) - loss_o) < 2e-5
    for got, ref in zip(m.grads, grads_o):
        assert_close(got, ref, rtol=2e-5, atol=2e-6, what="TapeMLP grad")
    np.testing.assert_array_equal(np.asarray(m.predict(x)), y_o.argmax(axis=1))     # bit-exact argmax
    xs = rng.standard_normal((256, 784)); ts = (xs[:, :10].argmax(axis=1)).astype(np.int64)
    tr = Trainer(TapeMLP(), Adam(lr=1e-3))
    tr.fit(xs, ts, np.zeros(256), max_epoch=8, batch_size=64, eval_interval=1, verbose=False)
    assert tr.loss_list[-1] < tr.loss_list[0]


@pytest.mark.parametrize("cell", ["rnn", "lstm"])
def test_rnn_lm_config4_reduced(backend, cell):
    from auto_differentiation_b200.models.model import RnnLM
    rng = np.random.default_rng(1238)
    V, Dm, H, N, T = 50, 16, 24, 4, 16
    m = RnnLM(vocab=V, wordvec=Dm, hidden=H, cell=cell)
    xs = rng.integers(0, V, (N, T)); ts = np.roll(xs, -1, axis=1)
    E, Wx, Wh, b, Wa, ba = _params_np(m)
    emb = O.embedding_forward(E, xs)
    if cell == "lstm":
        hs, caches, _, _ = O.time_lstm_forward(emb, Wx, Wh, b)
    else:
        hs, hprevs = O.time_rnn_forward(emb, Wx, Wh, b)
    score = O.time_affine_forward(hs, Wa, ba)
    loss_o, cache = O.time_softmax_with_loss_forward(score, ts)
    dscore = O.time_softmax_with_loss_backward(cache)
    dhs, dWa, dba = O.time_affine_backward(hs, Wa, dscore)
    if cell == "lstm":
        demb, (gWx, gWh, gb), _ = O.time_lstm_backward(caches, Wx, Wh, dhs)
    else:
        demb, (gWx, gWh, gb), _ = O.time_rnn_backward(emb, hs, hprevs, Wx, Wh, dhs)
    gE = O.time_embedding_backward(E.shape, xs, demb)
    loss = m.forward(xs, ts)
    m.backward()
    assert abs(float(loss) - loss_o) < 2e-5
    for got, ref, name in zip(m.grads, (gE, gWx, gWh, gb, dWa, dba), ("E", "Wx", "Wh", "b", "Wa", "ba")):
        assert_close(got, ref, rtol=3e-5, atol=3e-6, what="RnnLM(%s) g%s" % (cell, name))


# ---- data parallel: two gloo ranks on CPU -------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from abi_emulator import Emulator
    from auto_differentiation_b200 import _lib, device
    _lib._lib = Emulator()                       # CPU test: host logic only (see tests/abi_emulator.py)
    device._DEVICE = "cpu"
    device.require_cuda = lambda: None
    device.stream_ptr = lambda: 0
    from auto_differentiation_b200.data_parallel import GradientReducer, shard_rows
    from auto_differentiation_b200.models.model import AttentionClassifier
    rng = np.random.default_rng(77)
    B, S, d = 10, 8, 16
    x = rng.standard_normal((B, S, d)); t = rng.integers(0, 5, B)
    m = AttentionClassifier(d_model=d, heads=2, hidden=8, classes=5, micro_batch=2, seed=5)
    m.reducer = GradientReducer(m.flat_grads.t, m.bucket_slices())
    m.global_batch = B
    lo, hi = shard_rows(B, rank, world)
    loss = m.forward(x[lo:hi], t[lo:hi])
    m.backward()
    lt = torch.tensor([float(loss)], dtype=torch.float64)
    dist.all_reduce(lt)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=lt.numpy(), rows=np.array([lo, hi]),
             reduced=np.array([m.reducer.reduced_bytes]), **{"g%d" % i: np.asarray(g) for i, g in enumerate(m.grads)})
    dist.destroy_process_group()


def test_data_parallel_two_ranks_equal_single_process(tmp_path, backend):
    if backend != "emu":
        pytest.skip("the gloo/CPU DP test runs on the emulator; the NCCL path is exercised by bench.py --gpus N")
    import torch.multiprocessing as mp
    from auto_differentiation_b200.data_parallel import shard_rows
    from auto_differentiation_b200.models.model import AttentionClassifier
    assert [shard_rows(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert tuple(r0["rows"]) == (0, 5) and tuple(r1["rows"]) == (5, 10)
    rng = np.random.default_rng(77)
    B, S, d = 10, 8, 16
    x = rng.standard_normal((B, S, d)); t = rng.integers(0, 5, B)
    m = AttentionClassifier(d_model=d, heads=2, hidden=8, classes=5, micro_batch=4, seed=5)
    loss = m.forward(x, t)
    m.backward()
    assert abs(float(r0["loss"][0]) - float(loss)) < 1e-5
    nbytes = sum(g.size for g in m.grads) * 4
    assert int(r0["reduced"][0]) == nbytes                       # every gradient byte crossed exactly once
    for i, g in enumerate(m.grads):
        np.testing.assert_array_equal(r0["g%d" % i], r1["g%d" % i])          # ranks agree bit-for-bit
        assert_close(r0["g%d" % i], np.asarray(g, dtype=np.float64), rtol=1e-5, atol=1e-6, what="DP grad %d" % i)


def test_capturable_adam_matches_host_adam(backend):
    """The device-side step scalar (CUDA-graph friendly) gives the same updates as the host expression."""
    import auto_differentiation_b200 as pkg
    from auto_differentiation_b200.utils.optimizer import Adam
    rng = np.random.default_rng(4)
    p0 = [rng.standard_normal((33, 7)), rng.standard_normal((5000,))]
    gs = [[rng.standard_normal(p.shape) for p in p0] for _ in range(4)]
    a, b = Adam(lr=0.01), Adam(lr=0.01, capturable=True)
    pa, pb = [pkg.asarray(p) for p in p0], [pkg.asarray(p) for p in p0]
    for step in range(4):
        a.update(pa, [pkg.asarray(g) for g in gs[step]])
        b.update(pb, [pkg.asarray(g) for g in gs[step]])
    for x, y in zip(pa, pb):
        assert_close(x, np.asarray(y, dtype=np.float64), rtol=1e-6, atol=1e-7, what="capturable adam")


@pytest.mark.gpu
def test_cuda_graph_replay_equals_eager():
    """config-1 step captured in a CUDA graph: N replays == N eager steps (same seeds, same data)."""
    import torch
    from auto_differentiation_b200 import asarray
    from auto_differentiation_b200.graph import GraphedStep
    from auto_differentiation_b200.models.model import TapeMLP
    from auto_differentiation_b200.utils.optimizer import Adam
    rng = np.random.default_rng(1235)
    x = asarray(rng.standard_normal((64, 784)))
    t = rng.integers(0, 10, 64)
    onehot = np.zeros((64, 10), dtype=np.float32); onehot[np.arange(64), t] = 1.0
    td = asarray(onehot)

    def make():
        m, o = TapeMLP(), Adam(lr=1e-3, capturable=True)

        def step():
            loss = m.forward(x, td)
            m.backward()
            o.update(m.weights, m.grads)
            return loss
        return m, step
    m1, eager = make()
    for _ in range(3 + 1 + 5):           # GraphedStep runs 3 warm-up steps + 1 capture pass (not executed) ...
        eager()
    m2, step2 = make()
    g = GraphedStep(step2, warmup=3)     # 3 eager steps, then capture (capture itself does not execute)
    for _ in range(6):
        g()
    torch.cuda.synchronize()
    for a, b in zip(m1.weights, m2.weights):
        assert_close(b, np.asarray(a, dtype=np.float64), rtol=1e-6, atol=1e-7, what="graph vs eager weights")


def test_trainer_fit_graph_mode_equals_eager(backend):
    """`Trainer.fit(..., graph=True)`: static input buffers + one captured iteration replayed per batch must train
    exactly like the eager loop (same permutation stream, same updates); on the CPU backend it exercises the static
    buffers / device-label path without a capture."""
    from auto_differentiation_b200.models.model import TapeMLP
    from auto_differentiation_b200.utils.Train import Trainer
    from auto_differentiation_b200.utils.optimizer import Adam
    rng = np.random.default_rng(3)
    xs = rng.standard_normal((256, 784)).astype(np.float32)
    ts = xs[:, :10].argmax(axis=1).astype(np.int64)
    runs = []
    for graph in (False, True):
        np.random.seed(11)
        tr = Trainer(TapeMLP(seed=7), Adam(lr=1e-3))
        tr.fit(xs, ts, max_epoch=3, batch_size=64, eval_interval=2, verbose=False, graph=graph)
        runs.append((tr, [np.asarray(w, dtype=np.float64) for w in tr.model.weights]))
    (eager, w_e), (graphed, w_g) = runs
    assert len(eager.loss_list) == len(graphed.loss_list) == 6
    np.testing.assert_allclose(graphed.loss_list, eager.loss_list, rtol=1e-5, atol=1e-6)
    for a, b in zip(w_g, w_e):
        assert_close(a, b, rtol=1e-5, atol=1e-6, what="Trainer graph vs eager weights")
    assert graphed.optimizer.step_count() == 12
    if backend == "cuda":
        assert graphed.graphed.graph is not None, "the iteration was never captured"
