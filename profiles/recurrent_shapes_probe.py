"""Dependent-step cost (CUDA-graph chain of 64 steps, us per step) of out = tanh(out + h @ W) for small-M shapes:
exact-fp32 skinny kernels against the tcgen05 path (64-column tiles + split-K)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auto_differentiation_b200 import device as D, _lib as L
from auto_differentiation_b200.graph import GraphedStep

T = 64


def timeit(fn, reps=10):
    g = GraphedStep(fn, warmup=2)
    for _ in range(3):
        g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / T * 1e3


for (M, N, K) in [(64, 256, 256), (64, 512, 512), (64, 1024, 1024), (64, 2048, 2048), (16, 1024, 1024), (64, 1024, 256),
                  (64, 4096, 1024), (64, 1024, 4096)]:
    w = D.DeviceArray(torch.randn((K, N), device="cuda") / 32)
    wt = D.DeviceArray(torch.randn((N, K), device="cuda") / 32)
    h = D.DeviceArray(torch.randn((M, K), device="cuda") * 0.1)
    out = D.zeros((M, N))
    res = []
    for env in ({}, {"B200AD_NO_SKINNY": "1"}):
        os.environ.pop("B200AD_NO_SKINNY", None)
        os.environ.update(env)

        def nn():
            for _ in range(T):
                D.matmul(h, w, out=out, beta=1.0, act=L.ACT_TANH)

        def nt():
            for _ in range(T):
                D.matmul(h, wt, trans_b=True, out=out, beta=1.0, act=L.ACT_TANH)
        res.append((timeit(nn), timeit(nt)))
    print("M=%-3d N=%-5d K=%-5d  default NN %.2f NT %.2f | tcgen05 only NN %.2f NT %.2f us" % (M, N, K, res[0][0], res[0][1], res[1][0], res[1][1]), flush=True)
