import os, sys
sys.path.insert(0, "/root/repo")
import torch
from auto_differentiation_b200 import device as D, _lib as L
def bench(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(64, 1024, 1024), (64, 4096, 1024), (64, 1024, 4096)]:
    a = D.DeviceArray(torch.randn((M, K), device="cuda")); b = D.DeviceArray(torch.randn((K, N), device="cuda"))
    bt = D.DeviceArray(torch.randn((N, K), device="cuda"))
    out = D.zeros((M, N))
    for prec in ("tf32x3", "auto", "fp32"):
        t1 = bench(lambda: D.matmul(a, b, out=out, precision=prec))
        t2 = bench(lambda: D.matmul(a, b, out=out, beta=1.0, act=L.ACT_TANH, precision=prec))
        t3 = bench(lambda: D.matmul(a, bt, trans_b=True, out=out, precision=prec))
        print("M=%d N=%d K=%d %-7s NN %.1f us | NN beta+tanh %.1f us | NT %.1f us (launch-to-launch, eager)" % (M, N, K, prec, t1, t2, t3), flush=True)
