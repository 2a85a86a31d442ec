"""Diagnostic: is the K = 512 projection GEMM bound by its output stores?  Times the 131072 x 512 x 512 problem on the
1-CTA kernel and on the generation-3 pair kernel with and without the global stores (B200AD_TC_PROBE bit 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auto_differentiation_b200 import device as D

def run(M, N, K, ta=False):
    a = D.DeviceArray(torch.randn((K, M) if ta else (M, K), device="cuda"))
    b = D.DeviceArray(torch.randn((K, N), device="cuda"))
    out = D.empty((M, N))
    for _ in range(3):
        D.matmul(a, b, trans_a=ta, out=out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.matmul(a, b, trans_a=ta, out=out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[2]

for pair in ("0", "3"):
    for probe in ("0", "4"):
        os.environ["B200AD_TC_PAIR"] = pair
        os.environ["B200AD_TC_PROBE"] = probe
        t = run(131072, 512, 512)
        print("pair=%s probe=%s  131072x512x512: %.3f ms  %.1f TFLOP/s" % (pair, probe, t, 2 * 131072 * 512 * 512 / t / 1e9), flush=True)
