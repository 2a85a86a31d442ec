"""Secondary workloads (BASELINE.json configs 1-4) — `python bench.py --workload c1|c2|c3|c4`.
Each prints ONE JSON line; the collected tables live in profiles/r1_workloads.md."""
import json
import os
import time

import numpy as np


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time(torch, fn, warmup=3, min_ms=200.0, max_reps=200):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    reps, total = 0, 0.0
    while total < min_ms and reps < max_reps:
        e0, e1 = _events(torch)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
        reps += 1
    return total / reps


def run_c2(args, torch, pk):
    """elementwise + activation gradient sweep: forward + backward of each op, fp32, 1e3..1e9 elements,
    operand A (N/256, 256), broadcast operand b (1,256); algorithmic bytes/element from SURVEY §8(d)."""
    from auto_differentiation_b200 import _lib as L, device as D
    sizes = [10 ** e for e in range(3, 10)]
    if args.max_elems:
        sizes = [s for s in sizes if s <= args.max_elems]
    if getattr(args, "c2_sizes", None):
        sizes = list(args.c2_sizes)
    # bytes per element the B200 path MOVES (fwd + bwd) — the basis of GBps / frac_hbm below.  For mul and the unary ops
    # this equals SURVEY 8(d)'s out-of-place count (32 / 20 / 20).  Add differs: the reference's backward writes two
    # copies of g (24 B/elem same-shape, 16 broadcast), this path returns g itself for a same-shape operand and only
    # reads g once for the broadcast operand's column sum, so it moves 12 B/elem in both variants; quoting the
    # reference's byte count over this path's time would overstate the bandwidth (the 182 % of round 1).
    ops = {"add_bcast": 12, "mul_bcast": 20, "add_same": 12, "mul_same": 32, "exp": 20, "tanh": 20, "sigmoid": 20,
           "relu": 20}
    ref_bytes = {"add_bcast": 16, "mul_bcast": 20, "add_same": 24, "mul_same": 32, "exp": 20, "tanh": 20, "sigmoid": 20,
                 "relu": 20}
    table = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2
    for n_req in sizes:
        n = max(256, n_req // 256 * 256)
        rows = n // 256
        gen = torch.Generator(device="cuda").manual_seed(1236)
        a = D.DeviceArray(torch.rand((rows, 256), generator=gen, device="cuda") + 0.5)
        g = D.DeviceArray(torch.randn((rows, 256), generator=gen, device="cuda"))
        b_row = D.DeviceArray(torch.rand((1, 256), generator=gen, device="cuda") + 0.5)
        b_full = D.DeviceArray(torch.rand((rows, 256), generator=gen, device="cuda") + 0.5)
        x = D.DeviceArray(torch.randn((rows, 256), generator=gen, device="cuda"))
        for name, bpe in ops.items():
            if name.endswith("_bcast") or name.endswith("_same"):
                op = L.B_ADD if name.startswith("add") else L.B_MUL
                bb = b_row if name.endswith("_bcast") else b_full

                def fn(op=op, bb=bb):
                    D.binary(op, a, bb)
                    D.binary_bwd(op, g, a, bb, True, True, True)
            else:
                uop = {"exp": L.U_EXP, "tanh": L.U_TANH, "sigmoid": L.U_SIGMOID, "relu": L.U_RELU}[name]

                def fn(uop=uop):
                    y = D.unary(uop, x)
                    D.unary_bwd(uop, g, x=x, y=y)

            def fn_flushed(fn=fn):
                if n * 4 < (200 << 20):
                    flush.zero_()          # evict L2 between iterations for working sets that would fit
                fn()
            ms_with_flush = _time(torch, fn_flushed, min_ms=100.0)
            ms_flush = _time(torch, lambda: flush.zero_(), min_ms=50.0) if n * 4 < (200 << 20) else 0.0
            ms = max(ms_with_flush - ms_flush, 1e-4)
            row = {"ms": ms, "GBps": n * bpe / ms / 1e6, "frac_hbm": n * bpe / ms / 1e6 / pk["hbm"],
                   "bytes_per_elem_moved": bpe, "bytes_per_elem_reference": ref_bytes[name]}
            if n * 4 < (200 << 20) and not getattr(args, "c2_no_graph", False):
                # the same pair replayed from a CUDA graph (8 flush+fwd+bwd iterations per replay): removes the host
                # launch gaps that dominate the eager numbers below ~1e8 elements
                from auto_differentiation_b200.graph import GraphedStep

                def eight(f):
                    def run():
                        for _ in range(8):
                            f()
                    return run
                g_fn = GraphedStep(eight(fn_flushed), warmup=1)
                g_fl = GraphedStep(eight(lambda: flush.zero_()), warmup=1)
                g_ms = max((_time(torch, g_fn, min_ms=50.0) - _time(torch, g_fl, min_ms=50.0)) / 8, 1e-4)
                row.update(graph_ms=g_ms, graph_GBps=n * bpe / g_ms / 1e6, graph_frac_hbm=n * bpe / g_ms / 1e6 / pk["hbm"])
                del g_fn, g_fl
            table.setdefault(name, {})[str(n_req)] = row
        del a, g, b_full, x
        torch.cuda.empty_cache()
    big = str(sizes[-1])
    best = {k: v[big]["frac_hbm"] for k, v in table.items()}
    return {"metric": "elementwise fwd+bwd algorithmic GB/s", "unit": "GB/s", "value": float(np.mean([v[big]["GBps"] for v in table.values()])),
            "config": {"workload": "c2: elementwise+activation gradient sweep, fp32, %s..%s elements" % (sizes[0], big),
                       "l2": "L2 flushed (256 MB write) between iterations when the working set is < 200 MB"},
            "roofline": {"bound": "hbm", "peak": pk["hbm"], "unit": "GB/s", "frac_at_largest": best, "peak_source": pk["src"],
                         "frac": float(np.mean(list(best.values()))), "achieved": float(np.mean([v[big]["GBps"] for v in table.values()])),
                         "byte_basis": "bytes this path moves per element (fwd+bwd): add 12, mul same-shape 32 / broadcast 20, "
                                       "unary 20; equal to SURVEY 8(d) except add (reference: 24 / 16, it copies g twice)"},
            "table": table}


def run_c1(args, torch, pk):
    from auto_differentiation_b200.models.model import TapeMLP
    from auto_differentiation_b200.utils.optimizer import Adam
    from auto_differentiation_b200 import _lib as L
    rng = np.random.default_rng(1235)
    x = rng.standard_normal((64, 784)).astype(np.float32)
    t = rng.integers(0, 10, 64)
    m, opt = TapeMLP(), Adam(lr=1e-3)
    from auto_differentiation_b200 import asarray
    xd = asarray(x)

    def step():
        m.forward(xd, t)
        m.backward()
        opt.update(m.weights, m.grads)
    n0 = L.launch_count()
    step()
    per_step = L.launch_count() - n0
    ms = _time(torch, step, warmup=5, min_ms=500.0, max_reps=500)
    t0 = time.perf_counter()
    for _ in range(50):
        step()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) / 50 * 1e3
    # the same step captured once in a CUDA graph (device-resident one-hot labels, capturable Adam)
    from auto_differentiation_b200.graph import GraphedStep
    m2, opt2 = TapeMLP(), Adam(lr=1e-3, capturable=True)
    onehot = np.zeros((64, 10), dtype=np.float32)
    onehot[np.arange(64), t] = 1.0
    td = asarray(onehot)

    def gstep():
        loss = m2.forward(xd, td)
        m2.backward()
        opt2.update(m2.weights, m2.grads)
        return loss
    g = GraphedStep(gstep)
    for _ in range(20):
        g()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500):
        g()
    torch.cuda.synchronize()
    graph_ms = (time.perf_counter() - t0) / 500 * 1e3
    graph_loss = float(g.result)
    from oracle import np_oracle as O
    W1, b1, W2, b2 = [np.asarray(w, dtype=np.float64) for w in m.weights]
    ropt = O.Adam(lr=1e-3)
    ts = []
    for i in range(30):
        t0 = time.perf_counter()
        _, _, grads, _ = O.mlp_tape_step(x.astype(np.float64), t, W1, b1, W2, b2)
        ropt.update([W1, b1, W2, b2], grads)
        ts.append(time.perf_counter() - t0)
    cpu_ms = float(np.median(ts[5:])) * 1e3
    # through the reference's own training entry point: Trainer.fit (utils/Train.py:15-64), eager and graph=True
    from auto_differentiation_b200.utils.Train import Trainer
    xs = rng.standard_normal((64 * 64, 784)).astype(np.float32)
    ts = rng.integers(0, 10, 64 * 64)
    api = {}
    for mode, epochs in (("eager", 4), ("graph", 40)):
        tr = Trainer(TapeMLP(), Adam(lr=1e-3))
        tr.fit(xs[:64 * 8], ts[:64 * 8], max_epoch=1, batch_size=64, eval_interval=10 ** 9, verbose=False, graph=(mode == "graph"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr.fit(xs, ts, max_epoch=epochs, batch_size=64, eval_interval=10 ** 9, verbose=False, graph=(mode == "graph"))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        api[mode] = {"steps_per_s": 64 * epochs / dt, "iterations": 64 * epochs, "final_loss": tr.loss_list[-1]}
    flops = 78053376.0
    return {"metric": "fwd+bwd+update steps/sec", "unit": "steps/s", "value": api["graph"]["steps_per_s"],
            "ms_per_step": 1e3 / api["graph"]["steps_per_s"],
            "through_api": {"call": "Trainer(TapeMLP(), Adam()).fit(x, t, batch_size=64, graph=True) — wall clock incl. the per-batch "
                                    "input copy, one loss read-back per epoch", "trainer_fit_graph": api["graph"],
                            "trainer_fit_eager": api["eager"], "manual_step_eager_steps_per_s": 1e3 / wall_ms},
            "device_ms_per_step": ms, "gpu_launches_per_step": int(per_step),
            "roofline": {"bound": "launch latency (78 MFLOP/step: < 1 us of tensor time; SURVEY 8d says to report steps/s)",
                         "achieved": flops * api["graph"]["steps_per_s"] / 1e12, "unit": "TFLOP/s", "peak": pk["bf16_sustained"],
                         "frac": flops * api["graph"]["steps_per_s"] / 1e12 / pk["bf16_sustained"]},
            "cuda_graph": {"steps_per_s": 1e3 / graph_ms, "ms_per_step": graph_ms, "loss_after_523_steps": graph_loss},
            "config": {"workload": "c1: tape MLP 784-256-10, batch 64, Adam (eager, launch-latency bound: 78 MFLOP/step)"},
            "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "same step, fp64 NumPy port of the reference tape"}}


def run_c3(args, torch, pk):
    from auto_differentiation_b200.layers import layers as Ls
    from auto_differentiation_b200 import device as D
    B, S, d = args.c3_batch, 512, 512
    rng = np.random.default_rng(1237)
    ws = [rng.standard_normal((d, d)) / np.sqrt(d) for _ in range(4)]
    layer = Ls.MultiHeadAttention(d, d, 8, weights=ws)
    gen = torch.Generator(device="cuda").manual_seed(1237)
    x = D.DeviceArray(torch.randn((B, S, d), generator=gen, device="cuda"))
    dout = D.DeviceArray(torch.randn((B, S, d), generator=gen, device="cuda"))

    def step():
        layer.forward(x, x, x)
        layer.backward(dout)
    ms = _time(torch, step, min_ms=500.0)
    flops = 4831838208.0 * B
    # the same forward + backward replayed from a CUDA graph (43 launches: the eager loop is partly host-bound at B = 16)
    from auto_differentiation_b200.graph import GraphedStep
    g = GraphedStep(step)
    gms = _time(torch, g, min_ms=300.0)
    from oracle import np_oracle as O
    xs = rng.standard_normal((2, S, d))
    t0 = time.perf_counter()
    out, _ = O.mha_tape(xs, xs, xs, *ws, 8)
    O.backward(out, np.ones_like(out.val))
    cpu_s = time.perf_counter() - t0
    return {"metric": "MHA fwd+bwd steps/sec", "unit": "steps/s", "value": 1e3 / ms, "ms_per_step": ms,
            "through_api": {"call": "layer.forward(x, x, x); layer.backward(dout) (eager, device-timed)", "steps_per_s": 1e3 / ms,
                            "graphed_call": "GraphedStep(lambda: (layer.forward(x, x, x), layer.backward(dout)))()",
                            "graphed_steps_per_s": 1e3 / gms},
            "config": {"workload": "c3: MultiHeadAttention fwd+bwd, B=%d, S=512, d=512, 8 heads (3xTF32)" % B},
            "roofline": {"bound": "tensor", "achieved": flops / ms / 1e9, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": flops / ms / 1e9 / pk["bf16_sustained"], "peak_source": pk["src"] + " bf16 sustained"},
            "cuda_graph": {"steps_per_s": 1e3 / gms, "ms_per_step": gms, "TFLOPs": flops / gms / 1e9},
            "cpu_baseline": {"value": 1.0 / (cpu_s / 2 * B), "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "2 of %d sequences, fp64 tape twin, scaled linearly" % B}}


def run_c4(args, torch, pk):
    """Config 4 through the reference's training entry point: `Trainer.fit` on RnnLM (TimeEmbedding -> TimeRNN / TimeLSTM
    -> TimeAffine -> TimeSoftmaxWithLoss), eager and graph=True; one step = forward + backward + Adam over N=64, T=256."""
    from auto_differentiation_b200.models.model import RnnLM
    from auto_differentiation_b200.utils.optimizer import Adam
    from auto_differentiation_b200.utils.Train import Trainer
    from auto_differentiation_b200 import _lib as L
    out = {}
    rng = np.random.default_rng(1238)
    N, T, V = 64, 256, 1000
    n_batches = 4
    xs = rng.integers(0, V, (N * n_batches, T))
    ts = np.roll(xs, -1, axis=1)
    cells = getattr(args, "c4_cells", None) or ("rnn", "lstm")
    for cell in cells:
        m, opt = RnnLM(vocab=V, wordvec=256, hidden=1024, cell=cell), Adam(lr=1e-3)

        def step():
            m.forward(xs[:N], ts[:N])
            m.backward()
            opt.update(m.weights, m.grads)
        step()
        n0 = L.launch_count()
        step()
        launches = L.launch_count() - n0
        torch.cuda.synchronize()
        res = {"gpu_launches_per_step": int(launches)}
        for mode, epochs in (("eager", 2), ("graph", 4)):
            tr = Trainer(RnnLM(vocab=V, wordvec=256, hidden=1024, cell=cell), Adam(lr=1e-3))
            tr.fit(xs, ts, max_epoch=1, batch_size=N, eval_interval=10 ** 9, verbose=False, graph=(mode == "graph"))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tr.fit(xs, ts, max_epoch=epochs, batch_size=N, eval_interval=10 ** 9, verbose=False, graph=(mode == "graph"))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            res["trainer_fit_%s_ms_per_step" % mode] = dt / (epochs * n_batches) * 1e3
            res["trainer_fit_%s_final_loss" % mode] = tr.loss_list[-1]
        out[cell] = res
    flops = 229512314880.0
    first = out[cells[0]]
    best_ms = min(first["trainer_fit_eager_ms_per_step"], first["trainer_fit_graph_ms_per_step"])
    return {"metric": "fwd+bwd+update steps/sec", "unit": "steps/s", "value": 1e3 / best_ms, "ms_per_step": best_ms,
            "through_api": {"call": "Trainer(RnnLM(hidden=1024), Adam()).fit(xs, ts, batch_size=64[, graph=True]) — wall clock incl. "
                                    "the per-batch input copy", "cells": out},
            "config": {"workload": "c4: RNN LM V=1000 D=256 H=1024 T=256 N=64 (tanh cell; LSTM variant in `cells`)"},
            "cells": out,
            "roofline": {"bound": "tensor", "achieved": flops / best_ms / 1e9, "unit": "TFLOP/s",
                         "peak": pk["bf16_sustained"], "frac": flops / best_ms / 1e9 / pk["bf16_sustained"],
                         "note": "229.5 GFLOP/step (SURVEY 8d); the M=64 recurrence is a latency-bound dependent chain, "
                                 "the input projection / affine / weight-gradient GEMMs run on tcgen05 (3xTF32)"}}


def run_all(args, torch, pk, ClockSampler):
    """BASELINE configs[0..3] in one pass for the `workloads` block of the `bench.py --gpus 1` line: each entry carries
    its own clocks record (nvidia-smi sampled while it runs), roofline fraction and through-the-API number."""
    import copy
    out = {}
    a = copy.copy(args)
    a.c2_sizes, a.c2_no_graph, a.c3_batch = (10 ** 8, 10 ** 9), True, 16
    for name, fn in (("c1", run_c1), ("c2", run_c2), ("c3", run_c3), ("c4", run_c4)):
        sampler = ClockSampler(torch.cuda.current_device())
        sampler.start()
        t0 = time.perf_counter()
        try:
            line = fn(a, torch, pk)
        except Exception as e:                        # noqa: BLE001 - one failing side workload must not lose the headline
            line = {"error": "%s: %s" % (type(e).__name__, e)}
        torch.cuda.synchronize()
        line["clocks"] = sampler.stop()
        line["wall_s"] = time.perf_counter() - t0
        if name == "c2" and "table" in line:           # keep the line small: the per-op fractions are in roofline
            line["table"] = {k: {sz: {kk: round(vv, 4) if isinstance(vv, float) else vv for kk, vv in r.items()}
                                 for sz, r in v.items()} for k, v in line["table"].items()}
        out[name] = line
        torch.cuda.empty_cache()
    return out


def run(args):
    import torch
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import peaks
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    import auto_differentiation_b200 as pkg
    pkg.set_gemm_precision(args.precision)
    pk = peaks()
    fn = {"c1": run_c1, "c2": run_c2, "c3": run_c3, "c4": run_c4}[args.workload]
    from bench import ClockSampler
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    line = fn(args, torch, pk)
    torch.cuda.synchronize()
    line["clocks"] = sampler.stop()
    line.update({"n_gpus": 1, "dtype": "f32", "data": "synthetic", "gemm_precision": args.precision})
    print(json.dumps(line))
