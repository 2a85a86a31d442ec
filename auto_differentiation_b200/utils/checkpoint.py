"""Checkpointing.  `save` writes the reference's on-disk format — one `weights{index}.npy`
per de-duplicated parameter (RnnlmTrainer.save, layers/layers.py:815-818) — and, beyond the reference (which has no
load and drops the optimizer state), `adam_state.npz` with m / v / iter; `load` restores both in place."""
import os

import numpy as np

from .. import device as D
from ..optim_host import unique_params


def save(params, grads=None, directory=".", optimizer=None):
    """`grads` is accepted for call-compatibility with the trainer's `remove_duplicate(params, grads)` idiom; the
    de-duplication here only SELECTS parameters and never touches a gradient buffer."""
    params = unique_params(list(params))
    os.makedirs(directory, exist_ok=True)
    for index, p in enumerate(params):
        np.save(os.path.join(directory, f"weights{index}.npy"), np.asarray(p))
    if optimizer is not None and getattr(optimizer, "m", None) is not None:
        count = optimizer.step_count() if hasattr(optimizer, "step_count") else optimizer.iter
        state = {"iter": np.int64(count)}      # capturable Adam: the device counter (graph replays advance only that)
        for i, (m, v) in enumerate(zip(optimizer.m, optimizer.v)):
            state[f"m{i}"], state[f"v{i}"] = np.asarray(m), np.asarray(v)
        np.savez(os.path.join(directory, "adam_state.npz"), **state)
    return len(params)


def load(params, grads=None, directory=".", optimizer=None):
    """Restores `params` in place (identity preserved, so optimizer tables and weight tying stay valid)."""
    params = unique_params(list(params))
    for index, p in enumerate(params):
        a = np.load(os.path.join(directory, f"weights{index}.npy"))
        if tuple(a.shape) != tuple(p.shape):
            raise ValueError("weights%d.npy has shape %s, parameter has %s" % (index, a.shape, p.shape))
        D.copy_into_(D.asarray(a), p)
    path = os.path.join(directory, "adam_state.npz")
    if optimizer is not None and os.path.isfile(path):
        st = np.load(path)
        if hasattr(optimizer, "set_step_count"):
            optimizer.set_step_count(int(st["iter"]))          # also rewrites a live capturable device counter
        else:
            optimizer.iter = int(st["iter"])
        if optimizer.m is not None and len(optimizer.m) == len(params):
            # restore IN PLACE: a captured CUDA graph (and the optimizer's device table) keep pointing at these buffers
            for i in range(len(params)):
                D.copy_into_(D.asarray(st[f"m{i}"]), optimizer.m[i])
                D.copy_into_(D.asarray(st[f"v{i}"]), optimizer.v[i])
        else:
            optimizer.m = [D.asarray(st[f"m{i}"]) for i in range(len(params))]
            optimizer.v = [D.asarray(st[f"v{i}"]) for i in range(len(params))]
            optimizer._table = None
    return len(params)
