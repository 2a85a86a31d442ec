"""Launches every memory-bound kernel family once per repetition at ~1e8 elements, for an ncu pass

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
        --clock-control none -k regex:'unary_|binary_|softmax_|xent_|adam_multi|sgd_multi|colsum|rowsum|reduce_narrow|lstm_cell|act_bwd|bias_act|embedding_|axpby|fill_' \
        --csv --log-file gpurun_out/r2_membound_ncu.csv python profiles/membound_kernels.py

(achieved HBM GB/s for the elementwise / softmax / cross-entropy / optimizer / column-sum kernels).
Run without ncu it prints CUDA-event timings and GB/s of the same launches (L2 flushed between them)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from auto_differentiation_b200 import _lib as L, device as D      # noqa: E402
from auto_differentiation_b200.utils.optimizer import Adam        # noqa: E402
from auto_differentiation_b200.layers.layers import SGD           # noqa: E402


def main():
    n = int(os.environ.get("MEMBOUND_ELEMS", str(10 ** 8))) // 1024 * 1024
    reps = int(os.environ.get("MEMBOUND_REPS", "2"))
    rows256 = n // 256
    gen = torch.Generator(device="cuda").manual_seed(1)
    rand = lambda *s: D.DeviceArray(torch.randn(s, generator=gen, device="cuda"))
    x, g, a, b = rand(rows256, 256), rand(rows256, 256), rand(rows256, 256), rand(rows256, 256)
    brow = rand(1, 256)
    V = 1000
    rows_v = n // V
    logits, gsm = rand(rows_v, V), rand(rows_v, V)
    labels = D.as_index(np.random.default_rng(0).integers(0, V, rows_v))
    wide = rand(n // 1024, 1024)
    H = 1024
    nl = n // (4 * H)
    gates_a, c_prev = rand(nl, 4 * H), rand(nl, H)
    emb_w = rand(V, 256)
    emb_idx = D.as_index(np.random.default_rng(1).integers(0, V, n // 256))
    params = [rand(n // 4) for _ in range(4)]
    grads = [rand(n // 4) for _ in range(4)]
    adam, sgd = Adam(lr=1e-3), SGD(lr=0.1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    lib = L.load()
    results = {}
    for _ in range(200):                     # bring the clocks up before the first timed launch
        flush.zero_()
    torch.cuda.synchronize()

    def timed(name, nbytes, fn):
        best = None
        if not os.environ.get("MEMBOUND_NO_WARM"):
            fn()                      # first use of a kernel pays its lazy load: not part of the measurement
            torch.cuda.synchronize()
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        results[name] = {"ms": best, "algorithmic_bytes": nbytes, "GBps": nbytes / best / 1e6}
        return out

    y = timed("unary_fwd exp", 8 * n, lambda: D.unary(L.U_EXP, x))
    timed("unary_bwd exp", 12 * n, lambda: D.unary_bwd(L.U_EXP, g, x=x, y=y))
    y = timed("unary_fwd tanh", 8 * n, lambda: D.unary(L.U_TANH, x))
    timed("unary_bwd tanh", 12 * n, lambda: D.unary_bwd(L.U_TANH, g, x=x, y=y))
    y = timed("unary_fwd sigmoid", 8 * n, lambda: D.unary(L.U_SIGMOID, x))
    timed("unary_bwd sigmoid", 12 * n, lambda: D.unary_bwd(L.U_SIGMOID, g, x=x, y=y))
    y = timed("unary_fwd relu", 8 * n, lambda: D.unary(L.U_RELU, x))
    timed("unary_bwd relu", 12 * n, lambda: D.unary_bwd(L.U_RELU, g, x=x, y=y))
    timed("binary_fwd mul same", 12 * n, lambda: D.binary(L.B_MUL, a, b))
    timed("binary_bwd mul same", 20 * n, lambda: D.binary_bwd(L.B_MUL, g, a, b))
    timed("binary_fwd add bcast", 8 * n, lambda: D.binary(L.B_ADD, a, brow))
    timed("binary_bwd add bcast (column sum of g)", 4 * n, lambda: D.binary_bwd(L.B_ADD, g, a, brow))
    timed("binary_fwd mul bcast", 8 * n, lambda: D.binary(L.B_MUL, a, brow))
    timed("binary_bwd mul bcast (row-fused column sum)", 12 * n, lambda: D.binary_bwd(L.B_MUL, g, a, brow))
    p = timed("softmax_fwd V=1000", 8 * rows_v * V, lambda: D.softmax_lastaxis(logits))
    timed("softmax_bwd V=1000", 12 * rows_v * V, lambda: D.softmax_bwd_lastaxis(gsm, p))
    lc, probs = timed("xent_fwd V=1000", 8 * rows_v * V, lambda: D.xent_fwd(logits, labels, -1, 0.0, 0))
    timed("xent_bwd V=1000 (in place)", 8 * rows_v * V, lambda: D.xent_bwd(probs, labels, lc, 1.0, -1, out=probs))
    timed("colsum (rows,256)->(256)", 4 * n, lambda: D.reduce_mid(g, 1, rows256, 256))
    timed("rowsum (rows,1024)->(rows)", 4 * n, lambda: D.reduce_mid(wide, n // 1024, 1024, 1))
    adam.update(params, grads)
    timed("adam_multi 4 tensors", 28 * n, lambda: adam.update(params, grads))
    sgd.update(params, grads)
    timed("sgd_multi 4 tensors", 12 * n, lambda: sgd.update(params, grads))
    gbuf, cbuf, hbuf = D.empty((nl, 4 * H)), D.empty((nl, H)), D.empty((nl, H))
    timed("lstm_cell_fwd", 44 * nl * H, lambda: L.check(lib.b200ad_lstm_cell_fwd(
        gates_a.ptr, 4 * H, c_prev.ptr, H, gbuf.ptr, 4 * H, cbuf.ptr, H, hbuf.ptr, H, nl, H, D.stream_ptr())))
    timed("embedding_gather D=256", 8 * (n // 256) * 256, lambda: D.embedding_gather(emb_w, emb_idx))
    timed("fill", 4 * n, lambda: D.fill_(x, 0.5))
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6540.5
    for k, v in results.items():
        v["frac_hbm"] = v["GBps"] / peak
    print(json.dumps({"elements": n, "hbm_peak_GBps": peak, "kernels": results}))


if __name__ == "__main__":
    main()
