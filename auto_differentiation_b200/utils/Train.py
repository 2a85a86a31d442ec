"""Mini-batch training driver with the call surface of the reference's `utils.Train.Trainer` (reference :7-64):
`Trainer(model, optimizer).fit(x, t, q, max_epoch=, batch_size=, max_grad=, eval_interval=)` and the attributes
`loss_list`, `eval_interval`, `current_epoch`.

Behaviour kept: every epoch draws ONE permutation from the global NumPy RNG and reorders `x` and `t` cumulatively
(the optional third array is sliced in its original order, as in the reference); an iteration is forward ->
backward -> merge of shared weights -> optional global-norm clip -> optimizer update; the running mean of the loss
is appended to `loss_list` whenever `iteration % eval_interval == 0`.  Different on purpose: the loss is a device
scalar that is only read back at those evaluation points (one host sync per `eval_interval`, not per step), and
the third positional array is optional (without it the model receives `batch_size`, the reference's else-branch)."""
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


class Trainer:
    def __init__(self, model, optimizer):
        self.model = model
        self.optimizer = optimizer
        self.loss_list = []
        self.eval_interval = None
        self.current_epoch = 0

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

    def fit(self, x, t, *args, max_epoch=10, batch_size=32, max_grad=None, eval_interval=20, verbose=True):
        n_rows = len(x)
        n_iterations = n_rows // batch_size
        self.eval_interval = eval_interval
        extra = args[0] if args else None
        meter = _RunningLoss()
        began = time()
        for _ in range(max_epoch):
            order = np.random.permutation(np.arange(n_rows))
            x, t = x[order], t[order]
            for it in range(n_iterations):
                rows = slice(it * batch_size, (it + 1) * batch_size)
                third = extra[rows] if extra is not None else batch_size
                meter.add(self._train_step(x[rows], t[rows], third, max_grad))
                if eval_interval is not None and it % eval_interval == 0:
                    self._report(meter.pop_mean(), it, n_iterations, time() - began, verbose)
            self.current_epoch += 1
