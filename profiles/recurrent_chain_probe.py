"""Config-4 recurrence as a CUDA-graph chain: 256 dependent steps hs[t+1] = tanh(hs[t+1] + hs[t] @ Wh) (M=64, H=1024),
replayed from one graph; microseconds per step for each kernel choice, and for a chain of trivial kernels (graph node
overhead)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auto_differentiation_b200 import device as D, _lib as L
from auto_differentiation_b200.graph import GraphedStep

T, N, H = 256, 64, 1024
wh = D.DeviceArray(torch.randn((H, H), device="cuda") / 32)
hs = D.DeviceArray(torch.randn((T + 1, N, H), device="cuda") * 0.1)


def timeit(fn, reps=20):
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


def chain(trans_b=False):
    def fn():
        for t in range(T):
            D.matmul(hs[t], wh, trans_b=trans_b, out=hs[t + 1], beta=1.0, act=L.ACT_TANH)
    return fn


def trivial():
    for t in range(T):
        D.axpby_(hs[t][:1], hs[t + 1][:1], 1.0, 0.0)


print("trivial kernel chain            : %.2f us per node" % timeit(trivial), flush=True)
for label, env in (("skinny cluster (default)", {}), ("skinny, no cluster", {"B200AD_NO_SKINNY_CLUSTER": "1"}),
                   ("tensor-core path", {"B200AD_NO_SKINNY": "1"})):
    for k in ("B200AD_NO_SKINNY_CLUSTER", "B200AD_NO_SKINNY"):
        os.environ.pop(k, None)
    os.environ.update(env)
    print("%-32s: NN %.2f us per step, NT %.2f us per step" % (label, timeit(chain(False)), timeit(chain(True))), flush=True)
