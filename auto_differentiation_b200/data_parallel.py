"""Data-parallel plumbing: one process per GPU, the minibatch sharded by rows, parameters replicated,
gradient buckets all-reduced (sum) over NCCL/NVLink on a side stream as soon as the backward pass has
finished writing them.  The reference has no distributed code at all; this is the one
exchange step the path has.  `torch.distributed` is the transport (NCCL on GPUs; gloo for the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_rows(n_rows, rank, world):
    """Rows [lo, hi) of the global batch owned by `rank` (even split; the remainder goes to the first ranks)."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradientReducer:
    """All-reduces slices of ONE flat gradient buffer, bucket by bucket, in the order the backward pass
    finalises them.  `bucket_ready(i)` is called by the model right after the last kernel that writes
    bucket i was enqueued: the comm stream waits on an event recorded at that point, so the all-reduce
    of bucket i overlaps the rest of the backward; `wait()` makes the compute stream (and thus the
    optimizer) wait for all buckets."""

    def __init__(self, flat_grads, bucket_slices, group=None):
        self.flat = flat_grads              # torch tensor (1-D)
        self.slices = list(bucket_slices)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.on_gpu = flat_grads.is_cuda
        self.comm_stream = torch.cuda.Stream() if (self.on_gpu and self.world > 1) else None
        self._pending = []
        self._done = set()
        self.reduced_bytes = 0

    def bucket_ready(self, i):
        if self.world == 1 or i in self._done:
            return
        self._done.add(i)
        lo, hi = self.slices[i]
        view = self.flat[lo:hi]
        self.reduced_bytes += (hi - lo) * 4
        if self.on_gpu:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self._pending.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        """Reduce whatever was not announced, then join the comm stream into the compute stream."""
        for i in range(len(self.slices)):
            self.bucket_ready(i)
        if self.world > 1:
            if self.on_gpu:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            else:
                for w in self._pending:
                    w.wait()
        self._pending = []
        self._done = set()


def dp_parity(make_model, x_global, t_global, group=None):
    """SURVEY §8(e) parity check, run by every rank: the sum-all-reduced gradients of the W ranks (rows sharded by
    `shard_rows`) against ONE rank running the same global batch alone.  `make_model()` must build identical models
    (shared seed) exposing forward / backward / grads / flat_grads / bucket_slices / global_batch / reducer.
    Returns {"max_err_over_scale", "max_rel_err", ...} on rank 0 (None elsewhere); raises nothing — callers assert."""
    import numpy as np
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(x_global)
    lo, hi = shard_rows(n, rank, world)
    m = make_model()
    m.global_batch = n
    m.reducer = GradientReducer(m.flat_grads.t, m.bucket_slices(), group) if world > 1 else None
    loss = m.forward(x_global[lo:hi], t_global[lo:hi])
    m.backward()
    lt = loss.t.detach().clone().reshape(1)
    if world > 1:
        dist.all_reduce(lt, group=group)
    sharded = [np.asarray(g, dtype=np.float64) for g in m.grads]
    if rank != 0:
        return None
    single = make_model()
    single.global_batch = n
    single.reducer = None
    loss1 = single.forward(x_global, t_global)
    single.backward()
    worst_scaled, worst_rel, per = 0.0, 0.0, []
    for a, g in zip(sharded, single.grads):
        b = np.asarray(g, dtype=np.float64)
        scale = max(float(np.abs(b).max()), 1e-30)
        err = np.abs(a - b)
        # elementwise |a-b| <= atol*scale + rtol*|b| with rtol 1e-5, atol 1e-6 (summation order differs between W and 1)
        ratio = float(np.max(err / (1e-6 * scale + 1e-5 * np.abs(b))))
        per.append(ratio)
        worst_scaled = max(worst_scaled, float(err.max()) / scale)
        worst_rel = max(worst_rel, ratio)
    return {"world": world, "global_batch": int(n), "rtol": 1e-5, "atol_of_scale": 1e-6,
            "max_err_over_scale": worst_scaled, "worst_err_over_tol": worst_rel,
            "loss_sharded": float(lt.item()), "loss_single": float(loss1),
            "ok": bool(worst_rel <= 1.0 and abs(float(lt.item()) - float(loss1)) <= 1e-5 * max(1.0, abs(float(loss1))))}
