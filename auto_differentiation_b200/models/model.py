"""Models for the benchmark configs, written against the reference's model contract
(utils/Train.py:40-48): `.forward(x, t, q) -> loss`, `.backward()`, list attributes `.weights` / `.grads`
(updated in place), usable with `Trainer.fit` and `Adam.update` unchanged.

The reference's own `models/model.py` is unimportable as shipped and holds word2vec
models that are out of scope; these are the compositions SURVEY §8(d) defines for configs 1, 4 and 5.
"""
import math

import numpy as np

from .. import _lib as L
from .. import device as D
from ..device import DeviceArray, asarray
from ..auto_grad import tgrad as ad
from ..layers import layers as Ls


def _flatten_into(arrays):
    """Re-home `arrays` as consecutive views of ONE flat buffer (one fill / few all-reduces)."""
    total = sum(a.size for a in arrays)
    flat = D.empty((total,))
    views, off = [], 0
    for a in arrays:
        v = flat[off:off + a.size].reshape(a.shape)
        D.copy_into_(a.contiguous(), v)
        views.append(v)
        off += a.size
    return flat, views


def chunk_plan(n_rows, max_rows, first=32, growth=1.3, taper=2):
    """Row counts of the micro-batches a HOST-resident batch is streamed in.  The first chunk is small (its copy is the
    only one nothing overlaps), the sizes then grow geometrically (a chunk's copy must fit under the previous chunk's
    compute: growth < copy rate / compute rate) up to `max_rows`, and shrink again towards the end (each chunk at most
    1/taper of what remains) so that little compute is left once the last copy has landed — that tail is what a
    copy-bound rank (8 ranks sharing the host fabric) otherwise pays on top of its transfer time."""
    sizes, remaining, grow = [], int(n_rows), float(first)
    while remaining > 0:
        size = min(int(max_rows), int(grow) // 8 * 8 or int(grow), max(int(first), -(-remaining // taper)))
        size = max(1, min(size, remaining))
        sizes.append(size)
        remaining -= size
        grow = grow * growth
    return sizes


class _InputStreamer:
    """Host -> device streaming of row chunks of a host array, owned by the model and reused by every `forward`
    (nothing is allocated per step).  Each chunk is copied as sub-copies of <= `piece` rows on `n_streams` copy streams
    (default 1: with 8 ranks sharing the host fabric a second stream per rank measured 8 % slower, 54.7 vs 50.6 ms per
    step, and alone one stream already runs at the PCIe line rate, profiles/r2_h2d_probe_n8.json) into a ring of `depth`
    staging buffers; compute waits on the chunk's events, the copy streams wait
    on the event that says the kernels of the chunk that last used the buffer are done.  On a CPU-only test backend
    it degenerates to synchronous per-chunk uploads."""

    def __init__(self, max_rows, row_shape, depth=3):
        import os
        import torch
        self.torch = torch
        self.max_rows, self.row_shape, self.depth = int(max_rows), tuple(row_shape), depth
        self.piece = max(1, int(os.environ.get("B200AD_H2D_PIECE_ROWS", "32")))
        n_streams = max(1, int(os.environ.get("B200AD_H2D_STREAMS", "1")))
        self.streams = [torch.cuda.Stream() for _ in range(n_streams)]
        self.bufs = [torch.empty((self.max_rows,) + self.row_shape, dtype=torch.float32, device="cuda")
                     for _ in range(depth)]
        self.free = [None] * depth
        self.ready = {}
        self.src = None
        self.bounds = None

    def begin(self, x, bounds):
        """x: C-contiguous float32 host array (page-locked memory makes the copies asynchronous); bounds: [(lo, hi)]."""
        self.src = self.torch.from_numpy(x)
        self.bounds = bounds
        self.ready = {}
        self._issue(0)
        self._issue(1)

    def _issue(self, c):
        if c >= len(self.bounds) or c in self.ready:
            return
        torch, slot = self.torch, c % self.depth
        lo, hi = self.bounds[c]
        buf = self.bufs[slot]
        events = []
        for i, st in enumerate(self.streams):
            with torch.cuda.stream(st):
                if self.free[slot] is not None:
                    st.wait_event(self.free[slot])          # the kernels of chunk c-depth are done with this buffer
                for p0 in range(lo + i * self.piece, hi, self.piece * len(self.streams)):
                    p1 = min(hi, p0 + self.piece)
                    buf[p0 - lo:p1 - lo].copy_(self.src[p0:p1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
                events.append(ev)
        self.ready[c] = events

    def get(self, c):
        self._issue(c + 1)
        self._issue(c + 2)
        cur = self.torch.cuda.current_stream()
        for ev in self.ready.pop(c):
            cur.wait_event(ev)
        lo, hi = self.bounds[c]
        return DeviceArray(self.bufs[c % self.depth][:hi - lo])

    def release(self, c):
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.current_stream())
        self.free[c % self.depth] = ev


class TapeMLP:
    """Config 1: softmax(tanh(x@W1+b1)@W2+b2, 2) + mean cross-entropy, built op by op on the tape
    (`auto_grad` Matrix / gradient()) every step, like a user of the reference would."""

    def __init__(self, d_in=784, d_hidden=256, d_out=10, seed=1235):
        rng = np.random.default_rng(seed)
        self.W1 = ad.Matrix((rng.standard_normal((d_in, d_hidden)) / np.sqrt(d_in)), "W1")
        self.b1 = ad.Matrix(np.zeros((1, d_hidden)), "b1")
        self.W2 = ad.Matrix((rng.standard_normal((d_hidden, d_out)) / np.sqrt(d_hidden)), "W2")
        self.b2 = ad.Matrix(np.zeros((1, d_out)), "b2")
        self.leaves = [self.W1, self.b1, self.W2, self.b2]
        self.weights = [m.val for m in self.leaves]
        self.grads = [D.zeros(w.shape) for w in self.weights]
        self.d_out = d_out
        self._eye = None
        self._loss_node = None
        self.y = None

    def forward(self, x, t, q=None):
        n = len(x)
        X = ad.Matrix(x, requires_grad=False)
        if isinstance(t, DeviceArray) and t.ndim == 2:
            onehot = t                                                           # device-resident one-hot (graph capture)
        elif isinstance(t, DeviceArray):
            # device-resident integer labels (Trainer.fit(graph=True)): rows of the identity = the exact one-hot
            if self._eye is None:
                self._eye = asarray(np.eye(self.d_out, dtype=np.float32))
            onehot = D.embedding_gather(self._eye, D.as_index(t).reshape(n))
        else:
            onehot = np.zeros((n, self.d_out), dtype=np.float32)
            onehot[np.arange(n), np.asarray(t).astype(np.int64)] = 1.0           # exact one-hot (host, 64x10)
        T = ad.Matrix(onehot, requires_grad=False)
        self.y = ad.softmax(ad.tanh(X @ self.W1 + self.b1) @ self.W2 + self.b2, 2)
        picked = ad.sum(self.y * T, axis=1, keepdims=True) + 1e-7                # y[i, t_i] + 1e-7
        self._loss_node = ad.sum(ad.ln(picked), axis=0, keepdims=True) * (-1.0 / n)   # host scalar: folded into a kernel
        return self._loss_node.val.reshape(())

    def backward(self):
        for m in self.leaves:
            m.grad = None
        self._loss_node.gradient()
        for m, g in zip(self.leaves, self.grads):
            D.copy_into_(m.grad, g)

    def predict(self, x):
        X = ad.Matrix(x, requires_grad=False)
        y = ad.softmax(ad.tanh(X @ self.W1 + self.b1) @ self.W2 + self.b2, 2)
        return y.val.argmax(axis=1)


class RnnLM:
    """Config 4: TimeEmbedding -> tanh TimeRNN (or TimeLSTM) -> TimeAffine -> TimeSoftmaxWithLoss
    (pipeline of RNN_Implementation.ipynb:404-411; init of SURVEY §8d)."""

    def __init__(self, vocab=1000, wordvec=256, hidden=1024, cell="rnn", seed=1238):
        rng = np.random.default_rng(seed)
        V, Dm, H = vocab, wordvec, hidden
        G = 4 * H if cell == "lstm" else H
        embed_W = rng.standard_normal((V, Dm)) / 100
        Wx = rng.standard_normal((Dm, G)) / np.sqrt(Dm)
        Wh = rng.standard_normal((H, G)) / np.sqrt(H)
        b = np.zeros(G)
        affine_W = rng.standard_normal((H, V)) / np.sqrt(H)
        affine_b = np.zeros(V)
        self.embed = Ls.TimeEmbedding(embed_W)
        self.rnn = Ls.TimeLSTM(Wx, Wh, b) if cell == "lstm" else Ls.TimeRNN(Wx, Wh, b)
        self.affine = Ls.TimeAffine(affine_W, affine_b)
        self.loss_layer = Ls.TimeSoftmaxWithLoss()
        self.layers = [self.embed, self.rnn, self.affine]
        self.weights, self.grads = [], []
        for layer in self.layers:
            self.weights += layer.weights
            self.grads += layer.grads
        self.params = self.weights

    def forward(self, xs, ts, q=None):
        h = self.embed.forward(xs)
        h = self.rnn.forward(h)
        score = self.affine.forward(h)
        return self.loss_layer.forward(score, ts)

    def backward(self, dout=1):
        d = self.loss_layer.backward(dout)
        d = self.affine.backward(d)
        d = self.rnn.backward(d)
        self.embed.backward(d)
        return None


class AttentionClassifier:
    """Config 5: MultiHeadAttention(d_model, heads) on x (B,S,d) -> mean over the sequence -> affine
    d->hidden + tanh -> affine hidden->classes -> softmax cross-entropy (1,182,474 params at
    the default sizes).

    `forward` walks the local batch in micro-batches (the S x S score tensors of 8192 sequences do not
    fit HBM at once): each micro-batch runs forward AND backward, accumulating into the gradient buffers
    on the device; `backward()` then only completes the step (data-parallel all-reduce of the gradient
    buckets, overlapped with the tail of the last micro-batch).  Gradients are those of the mean loss
    over `global_batch` rows, so a sum all-reduce over ranks gives the single-process gradient."""

    def __init__(self, d_model=512, heads=8, hidden=256, classes=10, seed=1239, micro_batch=512, reducer=None,
                 input_grads=True):
        rng = np.random.default_rng(seed)
        d = d_model
        ws = [rng.standard_normal((d, d)) / np.sqrt(d) for _ in range(4)]
        W1 = rng.standard_normal((d, hidden)) / np.sqrt(d)
        W2 = rng.standard_normal((hidden, classes)) / np.sqrt(hidden)
        params = [asarray(w) for w in ws] + [asarray(W1), D.zeros((hidden,)), asarray(W2), D.zeros((classes,))]
        self.flat_params, self.weights = _flatten_into(params)
        self.flat_grads, self.grads = _flatten_into([D.zeros(p.shape) for p in params])
        self.mha = Ls.MultiHeadAttention(d, d, heads, weights=self.weights[:4])
        self.mha.grads = self.grads[:4]
        self.mha.accumulate = True
        self.mha.input_grads = input_grads
        self.micro_batch = micro_batch
        self.reducer = reducer
        self.global_batch = None
        self.loss_buf = D.zeros((1,))
        self.last_logits = None
        self._streamer, self._streamer_key = None, None

    # gradient buckets in the order they become final during the LAST micro-batch's backward
    def bucket_slices(self):
        sizes = [g.size for g in self.grads]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        head = (int(offs[4]), int(offs[8]))
        return [head, (int(offs[3]), int(offs[4])), (int(offs[2]), int(offs[3])), (int(offs[1]), int(offs[2])),
                (int(offs[0]), int(offs[1]))]

    def zero_grads(self):
        D.fill_(self.flat_grads, 0.0)
        D.fill_(self.loss_buf, 0.0)

    def chunk_step(self, xc, tc, scale, notify=None):
        """forward + backward of one micro-batch; grads += d(scale * mean-loss of the chunk)."""
        W1, b1, W2, b2 = self.weights[4:]
        gW1, gb1, gW2, gb2 = self.grads[4:]
        mb, S, d = xc.shape
        a = self.mha.forward(xc, xc, xc)                                         # (mb,S,d)
        pooled = D.reduce_mid(a.contiguous(), mb, S, d, 1.0 / S).reshape(mb, d)  # mean over the sequence
        h1 = D.matmul(pooled, W1, bias=b1, act=L.ACT_TANH)
        logits = D.matmul(h1, W2, bias=b2)
        lc, probs = D.xent_fwd(logits, tc, -1, 0.0, 0)
        D.axpby_(lc[0:1], self.loss_buf, scale, 1.0)
        dlogits = D.xent_bwd(probs, tc, lc, scale, -1, out=probs)
        D.matmul(h1, dlogits, trans_a=True, out=gW2, beta=1.0)
        D.axpby_(D.reduce_mid(dlogits, 1, mb, dlogits.shape[1]), gb2, 1.0, 1.0)
        dz1 = D.act_bwd(D.matmul(dlogits, W2, trans_b=True), h1, L.ACT_TANH)
        D.matmul(pooled, dz1, trans_a=True, out=gW1, beta=1.0)
        D.axpby_(D.reduce_mid(dz1, 1, mb, dz1.shape[1]), gb1, 1.0, 1.0)
        if notify:
            notify(0)                                                            # head bucket is final
        dpooled = D.matmul(dz1, W1, trans_b=True)
        da = D.repeat_axis(dpooled.reshape(mb, 1, d), S, 1, 1.0 / S)
        self.mha.backward(da, notify=notify)
        self.last_logits = logits

    def forward(self, x, t, q=None):
        """x: (B_local, S, d) DeviceArray or host array; t: (B_local,) labels. Returns the mean loss
        contribution of this rank (sum over ranks = global mean loss) as a 0-d DeviceArray.

        Host input (NumPy fp32; page-locked memory makes the copies asynchronous) is streamed: the
        host->device copy of micro-batch c+1 runs on a copy stream while micro-batch c computes."""
        B = len(x)
        gb = self.global_batch or B
        self.zero_grads()
        if B == 0:
            # a rank whose row shard is empty (global batch < world size) contributes zero gradients and zero loss, and
            # still joins every all-reduce in `backward()` — the other ranks would otherwise hang there
            return self.loss_buf.reshape(())
        mb = min(self.micro_batch, B)
        t_dev = D.as_index(t)
        streamed = (D._DEVICE == "cuda" and isinstance(x, np.ndarray) and x.dtype == np.float32
                    and x.flags.c_contiguous)
        if streamed:
            sizes = chunk_plan(B, mb)
        else:
            sizes = [mb] * (B // mb) + ([B % mb] if B % mb else [])
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        bounds = [(int(offs[i]), int(offs[i + 1])) for i in range(len(sizes))]
        feeder = None
        if streamed:
            key = (mb, tuple(x.shape[1:]))
            if self._streamer is None or self._streamer_key != key:
                self._streamer, self._streamer_key = _InputStreamer(mb, x.shape[1:]), key
            feeder = self._streamer
            feeder.begin(x, bounds)
        for c, (lo, hi) in enumerate(bounds):
            if feeder is not None:
                xc = feeder.get(c)
            elif isinstance(x, DeviceArray):
                xc = x[lo:hi]
            else:
                xc = asarray(np.asarray(x[lo:hi]))
            last = (c == len(bounds) - 1)
            notify = self.reducer.bucket_ready if (last and self.reducer is not None) else None
            self.chunk_step(xc, t_dev[lo:hi], (hi - lo) / gb, notify)
            if feeder is not None:
                feeder.release(c)
        return self.loss_buf.reshape(())

    def backward(self):
        if self.reducer is not None:
            self.reducer.wait()
        return None
