"""Checkpointing.  `save` writes the reference's on-disk format — one `weights{index}.npy`
per de-duplicated parameter (RnnlmTrainer.save, layers/layers.py:815-818) — and, beyond the reference (which has no
load and drops the optimizer state), `adam_state.npz` with m / v / iter; `load` restores both in place."""
import os

import numpy as np

from .. import device as D
from ..optim_host import remove_duplicate


def save(params, grads=None, directory=".", optimizer=None):
    if grads is not None:
        params, _ = remove_duplicate(list(params), list(grads))
    os.makedirs(directory, exist_ok=True)
    for index, p in enumerate(params):
        np.save(os.path.join(directory, f"weights{index}.npy"), np.asarray(p))
    if optimizer is not None and getattr(optimizer, "m", None) is not None:
        state = {"iter": np.int64(optimizer.iter)}
        for i, (m, v) in enumerate(zip(optimizer.m, optimizer.v)):
            state[f"m{i}"], state[f"v{i}"] = np.asarray(m), np.asarray(v)
        np.savez(os.path.join(directory, "adam_state.npz"), **state)
    return len(params)


def load(params, grads=None, directory=".", optimizer=None):
    """Restores `params` in place (identity preserved, so optimizer tables and weight tying stay valid)."""
    if grads is not None:
        params, _ = remove_duplicate(list(params), list(grads))
    for index, p in enumerate(params):
        a = np.load(os.path.join(directory, f"weights{index}.npy"))
        if tuple(a.shape) != tuple(p.shape):
            raise ValueError("weights%d.npy has shape %s, parameter has %s" % (index, a.shape, p.shape))
        D.copy_into_(D.asarray(a), p if p.t.is_contiguous() else p)
    path = os.path.join(directory, "adam_state.npz")
    if optimizer is not None and os.path.isfile(path):
        st = np.load(path)
        optimizer.iter = int(st["iter"])
        optimizer.m = [D.asarray(st[f"m{i}"]) for i in range(len(params))]
        optimizer.v = [D.asarray(st[f"v{i}"]) for i in range(len(params))]
        optimizer._table = None
    return len(params)
