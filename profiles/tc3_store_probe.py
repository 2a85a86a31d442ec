"""gen3 GEMM epilogue: TMA-store path (default) against the coalesced st.global path (B200AD_TC_NO_TMA_STORE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auto_differentiation_b200 import device as D


def run(M, N, K, ta=False, tb=False):
    a = D.DeviceArray(torch.randn((K, M) if ta else (M, K), device="cuda"))
    b = D.DeviceArray(torch.randn((N, K) if tb else (K, N), device="cuda"))
    out = D.empty((M, N))
    for _ in range(3):
        D.matmul(a, b, trans_a=ta, trans_b=tb, out=out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.matmul(a, b, trans_a=ta, trans_b=tb, out=out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[3]


for no_tma in ("", "1"):
    if no_tma:
        os.environ["B200AD_TC_NO_TMA_STORE"] = "1"
    for name, (M, N, K, ta, tb) in {"C5 projection": (131072, 512, 512, False, False), "C5 dX": (131072, 512, 512, False, True),
                                    "C4 gates": (16384, 4096, 256, False, False), "square 4096": (4096, 4096, 4096, False, True)}.items():
        t = run(M, N, K, ta, tb)
        print("no_tma_store=%-1s %-14s %.3f ms  %.1f TFLOP/s" % (no_tma or "0", name, t, 2.0 * M * N * K / t / 1e9), flush=True)
