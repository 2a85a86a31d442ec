"""Mini-batch training driver with the call surface of the reference's `utils.Train.Trainer` (reference :7-64):
`Trainer(model, optimizer).fit(x, t, q, max_epoch=, batch_size=, max_grad=, eval_interval=)` and the attributes
`loss_list`, `eval_interval`, `current_epoch`.

Behaviour kept: every epoch draws ONE permutation from the global NumPy RNG and reorders `x` and `t` cumulatively
(the optional third array is sliced in its original order, as in the reference); an iteration is forward ->
backward -> merge of shared weights -> optional global-norm clip -> optimizer update; the running mean of the loss
is appended to `loss_list` whenever `iteration % eval_interval == 0`.  Different on purpose: the loss is a device
scalar that is only read back at those evaluation points (one host sync per `eval_interval`, not per step), and
the third positional array is optional (without it the model receives `batch_size`, the reference's else-branch).

`fit(..., graph=True)` (an addition; the default keeps the eager loop): the whole iteration — forward, backward,
weight merge, clip, optimizer update and the running-loss accumulation — is captured ONCE into a CUDA graph over
static device input buffers after a few eager warm-up iterations, and every later iteration is one host->device
copy of the batch plus one graph replay.  This is what the launch-latency-bound configs need (config 1: ~40 kernels
for 78 MFLOP; config 4: hundreds of dependent small launches).  Requirements: the model accepts device-resident
inputs (fp32 `x`, int64 labels) and builds the same launch sequence every iteration; the optimizer is switched to
its `capturable` form (step counter on the device)."""
from time import time

import numpy as np

from .useful_func import remove_duplicate
from ..optim_host import clip_grads


class _RunningLoss:
    """Sum of device-resident losses since the last read-out."""

    def __init__(self):
        self.total, self.count = 0, 0

    def add(self, loss):
        self.total = self.total + loss
        self.count += 1

    def pop_mean(self):
        mean = float(self.total) / self.count          # the only device -> host read of the loop
        self.total, self.count = 0, 0
        return mean


class _GraphedIteration:
    """Static input buffers + one captured training iteration (see the module docstring)."""

    def __init__(self, trainer, x, t, extra, batch_size, max_grad, warmup=3):
        import torch
        from .. import device as D
        self.torch, self.D = torch, D
        self.on_gpu = D._DEVICE == "cuda"
        self.x_is_index = np.asarray(x).dtype.kind in "iu"            # token ids (config 4) stay int64 on the device
        self.x_buf = D.empty((batch_size,) + tuple(x.shape[1:]), torch.int64 if self.x_is_index else torch.float32)
        t = np.asarray(t)
        self.t_is_index = t.dtype.kind in "iu"
        self.t_buf = D.empty((batch_size,) + tuple(t.shape[1:]), torch.int64 if self.t_is_index else torch.float32)
        self.extra_buf = None
        if extra is not None and np.asarray(extra).dtype.kind == "f":
            self.extra_buf = D.empty((batch_size,) + tuple(np.asarray(extra).shape[1:]))
        self.third = self.extra_buf if self.extra_buf is not None else batch_size
        self.loss_sum = D.zeros((1,))
        self.count = 0
        self.trainer, self.max_grad = trainer, max_grad
        self.warmup, self.calls, self.graph = warmup, 0, None
        if hasattr(trainer.optimizer, "capturable"):
            trainer.optimizer.capturable = True
        self.epoch_x = self.epoch_t = self.epoch_extra = None

    def _x_dtype(self):
        return np.int64 if self.x_is_index else np.float32

    def _iteration(self):
        loss = self.trainer._train_step(self.x_buf, self.t_buf, self.third, self.max_grad)
        self.D.axpby_(loss.reshape(1), self.loss_sum, 1.0, 1.0)

    def stage_epoch(self, x, t, extra):
        """Datasets up to 2 GiB are uploaded once per epoch (the reference reshuffles per epoch) and sliced on the device;
        larger ones are copied batch by batch from the host arrays."""
        torch = self.torch
        if x.nbytes + np.asarray(t).nbytes <= (2 << 30):
            dev = self.x_buf.t.device
            self.epoch_x = torch.from_numpy(np.ascontiguousarray(x, dtype=self._x_dtype())).to(dev, non_blocking=True)
            tt = np.ascontiguousarray(t, dtype=np.int64 if self.t_is_index else np.float32)
            self.epoch_t = torch.from_numpy(tt).to(dev, non_blocking=True)
            if self.extra_buf is not None:
                self.epoch_extra = torch.from_numpy(np.ascontiguousarray(extra, dtype=np.float32)).to(dev, non_blocking=True)
        else:
            self.epoch_x = torch.from_numpy(np.ascontiguousarray(x, dtype=self._x_dtype()))
            self.epoch_t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.int64 if self.t_is_index else np.float32))
            if self.extra_buf is not None:
                self.epoch_extra = torch.from_numpy(np.ascontiguousarray(extra, dtype=np.float32))

    def run(self, rows):
        self.x_buf.t.copy_(self.epoch_x[rows], non_blocking=True)
        self.t_buf.t.copy_(self.epoch_t[rows], non_blocking=True)
        if self.extra_buf is not None:
            self.extra_buf.t.copy_(self.epoch_extra[rows], non_blocking=True)
        self.count += 1
        if self.graph is not None:
            self.graph.replay()
            return
        self._iteration()
        self.calls += 1
        if self.on_gpu and self.calls == self.warmup:
            # first-use work (kernel attributes, optimizer tables, workspace sizes) is done: capture the next iteration.
            # Capture records without executing, so the captured iteration is replayed by the NEXT call to run().
            torch = self.torch
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._iteration()

    def pop_mean(self):
        mean = float(self.loss_sum) / max(self.count, 1)      # the only device -> host read of the loop
        self.D.fill_(self.loss_sum, 0.0)
        self.count = 0
        return mean


class Trainer:
    def __init__(self, model, optimizer):
        self.model = model
        self.optimizer = optimizer
        self.loss_list = []
        self.eval_interval = None
        self.current_epoch = 0
        self.graphed, self._graphed_key = None, None

    def _train_step(self, bx, bt, third, max_grad):
        loss = self.model.forward(bx, bt, third)
        self.model.backward()
        params, grads = remove_duplicate(self.model.weights, self.model.grads)
        if max_grad is not None:
            clip_grads(grads, max_grad)
        self.optimizer.update(params, grads)
        return loss

    def _report(self, mean_loss, iteration, n_iterations, seconds, verbose):
        if verbose:
            print("\r| epoch %-5d |  iter %-5d / %-5d | time %10d[s] | loss %.2f"
                  % (self.current_epoch + 1, iteration + 1, n_iterations, seconds, mean_loss), end="", flush=True)
        self.loss_list.append(mean_loss)

    def fit(self, x, t, *args, max_epoch=10, batch_size=32, max_grad=None, eval_interval=20, verbose=True,
            graph=False):
        n_rows = len(x)
        n_iterations = n_rows // batch_size
        self.eval_interval = eval_interval
        extra = args[0] if args else None
        graphed = None
        if graph:
            # the captured iteration is kept across fit() calls of the same batch geometry (capture costs tens of ms)
            key = (batch_size, tuple(np.shape(x)[1:]), np.asarray(x).dtype.kind, tuple(np.shape(t)[1:]),
                   np.asarray(t).dtype.kind, None if extra is None else tuple(np.shape(extra)[1:]), max_grad)
            if getattr(self, "_graphed_key", None) != key:
                self.graphed, self._graphed_key = _GraphedIteration(self, x, t, extra, batch_size, max_grad), key
            graphed = self.graphed
        meter = graphed if graphed is not None else _RunningLoss()
        began = time()
        for _ in range(max_epoch):
            order = np.random.permutation(np.arange(n_rows))
            x, t = x[order], t[order]
            if graphed is not None:
                graphed.stage_epoch(x, t, extra)
            for it in range(n_iterations):
                rows = slice(it * batch_size, (it + 1) * batch_size)
                if graphed is not None:
                    graphed.run(rows)
                else:
                    third = extra[rows] if extra is not None else batch_size
                    meter.add(self._train_step(x[rows], t[rows], third, max_grad))
                if eval_interval is not None and it % eval_interval == 0:
                    self._report(meter.pop_mean(), it, n_iterations, time() - began, verbose)
            self.current_epoch += 1
