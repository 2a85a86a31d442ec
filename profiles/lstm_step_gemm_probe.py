"""Which orientation of the LSTM recurrent products runs faster on the tcgen05 split-K GEMM: batch rows on M (64 of the 128
tile rows used) or the weight's big dimension on M (operands swapped, result transposed)?  Times 200 launches each inside a
CUDA graph (weights stay L2-resident like in the time loop).  N = 64, H = 1024."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import auto_differentiation_b200 as pkg
from auto_differentiation_b200 import device as D, _lib as L

N, H = 64, 1024
g = torch.Generator(device="cuda").manual_seed(0)
rand = lambda *s: D.DeviceArray(torch.randn(s, generator=g, device="cuda"))
wh, h, dA = rand(H, 4 * H), rand(N, H), rand(N, 4 * H)
P = L.GEMM_TF32X3 if hasattr(L, "GEMM_TF32X3") else 1


def timed(fn, reps=200):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


out_f, out_ft = D.empty((N, 4 * H)), D.empty((4 * H, N))
out_b, out_bt = D.empty((N, H)), D.empty((H, N))
cases = [
    ("fwd  h @ Wh            (M=64,   N=4096, K=1024)", lambda: D.matmul(h, wh, out=out_f, precision=P)),
    ("fwd  Wh^T @ h^T        (M=4096, N=64,   K=1024)", lambda: D.matmul(wh, h, out=out_ft, precision=P, trans_a=True, trans_b=True)),
    ("bwd  dA @ Wh^T         (M=64,   N=1024, K=4096)", lambda: D.matmul(dA, wh, out=out_b, precision=P, trans_b=True)),
    ("bwd  Wh @ dA^T         (M=1024, N=64,   K=4096)", lambda: D.matmul(wh, dA, out=out_bt, precision=P, trans_b=True)),
]
for name, fn in cases:
    print("%-52s %7.2f us per launch pair (product + split-K reduce)" % (name, timed(fn)), flush=True)
ref = np.asarray(h, dtype=np.float64) @ np.asarray(wh, dtype=np.float64)
print("max |fwd - fwd^T^T|", float(np.abs(np.asarray(out_f) - np.asarray(out_ft).T).max()), "vs fp64", float(np.abs(np.asarray(out_ft).T - ref).max()))
