"""GPU probe of the fp16 hi/lo three-product GEMM (B200AD_GEMM_F16X3, csrc/gemm_h16.cu) against the in-kernel 3xTF32 split
(gemm_tc3_kernel): error vs fp64 (on a sample of rows at the large sizes) and device time per product INCLUDING the operand
split pre-pass, at the config-5 projection shapes, the weight-gradient shape, config 3's M and 8192^3.
    python profiles/gemm_h16_probe.py [quick]
Every case runs under a bounded wait inside the kernel (mbar_wait traps after ~2 s); the first case is tiny on purpose.
"""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
from auto_differentiation_b200 import device as D

rng = np.random.default_rng(0)
results = []


def case(name, M, N, K, ta=False, tb=False, precs=("f16x3", "tf32x3"), reps=10, sample_rows=256, scale_rows=False):
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    if scale_rows:          # operand rows / columns spanning ten decades
        sa = (10.0 ** rng.uniform(-6, 4, size=M)).astype(np.float32)
        sb = (10.0 ** rng.uniform(-6, 4, size=N)).astype(np.float32)
        A = A * (sa[None, :] if ta else sa[:, None])
        B = B * (sb[:, None] if tb else sb[None, :])
    a, b = D.asarray(A), D.asarray(B)
    rows = np.arange(M) if M <= sample_rows else np.sort(rng.choice(M, sample_rows, replace=False))
    Aop = (A.T if ta else A)[rows].astype(np.float64)
    Bop = (B.T if tb else B).astype(np.float64)
    ref = Aop @ Bop
    bound = np.abs(Aop) @ np.abs(Bop) + 1e-300
    rec = {"case": name, "M": M, "N": N, "K": K, "ta": int(ta), "tb": int(tb)}
    for prec in precs:
        try:
            got = D.matmul(a, b, precision=prec, trans_a=ta, trans_b=tb)
            torch.cuda.synchronize()
            g = np.asarray(got, dtype=np.float64)[rows]
            err = np.abs(g - ref)
            for _ in range(3):
                D.matmul(a, b, precision=prec, trans_a=ta, trans_b=tb)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                D.matmul(a, b, precision=prec, trans_a=ta, trans_b=tb)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rec[prec] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "max_err_over_sqrtK": float(err.max() / np.sqrt(K)),
                         "max_err_over_absAabsB": float((err / bound).max())}
            print("%-26s M=%-6d N=%-5d K=%-6d ta=%d tb=%d %-7s %8.3f ms %7.1f TFLOP/s  err/sqrtK=%.2e  err/(|A||B|)=%.2e" % (
                name, M, N, K, ta, tb, prec, ms, rec[prec]["tflops"], rec[prec]["max_err_over_sqrtK"], rec[prec]["max_err_over_absAabsB"]), flush=True)
        except Exception as e:  # noqa: BLE001
            rec[prec] = {"error": str(e)[:300]}
            print("%-26s M=%d N=%d K=%d %s FAILED: %s" % (name, M, N, K, prec, str(e)[:300]), flush=True)
    results.append(rec)


def shared_case(M, N, K, reps=10):
    A = rng.standard_normal((M, K)).astype(np.float32)
    Ws = [D.asarray((rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)) for _ in range(3)]
    a = D.asarray(A)
    rec = {"case": "shared A x3", "M": M, "N": N, "K": K}
    for label, fn in (("shared f16x3", lambda: D.matmul_shared_a(a, Ws, precision="f16x3")),
                      ("3 calls f16x3", lambda: [D.matmul(a, w, precision="f16x3") for w in Ws]),
                      ("3 calls tf32x3", lambda: [D.matmul(a, w, precision="tf32x3") for w in Ws])):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec[label] = {"ms": ms, "tflops": 3 * 2.0 * M * N * K / ms / 1e9}
        print("shared-A x3 %-16s M=%d N=%d K=%d %8.3f ms %7.1f TFLOP/s" % (label, M, N, K, ms, rec[label]["tflops"]), flush=True)
    results.append(rec)


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
case("one pair tile", 256, 256, 128, precs=("f16x3",), reps=2)
case("tails", 300, 200, 100, precs=("f16x3",), reps=2)
case("tails TT", 300, 200, 100, ta=True, tb=True, precs=("f16x3",), reps=2)
case("many tiles", 4096, 1024, 512, reps=5)
case("decades", 1024, 512, 512, reps=2, scale_rows=True)
if not quick:
    case("c5 projection fwd (NN)", 131072, 512, 512)
    case("c5 projection bwd (NT)", 131072, 512, 512, tb=True)
    case("c5 micro-batch 256 (NN)", 65536, 512, 512)
    case("c3 projection (NN)", 8192, 512, 512)
    case("c4 gate product (NN)", 32768, 4096, 256)
    case("c4 vocabulary (NN)", 32768, 1000, 1024)
    case("weight gradient (TN)", 512, 512, 131072, ta=True)
    case("square 4096 (NN)", 4096, 4096, 4096, reps=5)
    case("square 8192 (NN)", 8192, 8192, 8192, reps=3)
    case("square 8192 (NT)", 8192, 8192, 8192, tb=True, reps=3)
    shared_case(131072, 512, 512)
    shared_case(65536, 512, 512)
out = os.path.join(ROOT, "gpurun_out", "gemm_h16_probe.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "w") as f:
    json.dump(results, f, indent=1)
print("wrote", out)
