"""Role-cycle counters of the generation-3 GEMM (build with `make NVCCFLAGS_EXTRA=-DTC3_PROFILE`): prints, for block 0,
where the MMA warp, the splitter and the epilogue warps spend their cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auto_differentiation_b200 import device as D


def run(M, N, K, ta=False, tb=False):
    a = D.DeviceArray(torch.randn((K, M) if ta else (M, K), device="cuda"))
    b = D.DeviceArray(torch.randn((N, K) if tb else (K, N), device="cuda"))
    out = D.empty((M, N))
    D.matmul(a, b, trans_a=ta, trans_b=tb, out=out)
    torch.cuda.synchronize()
    print("=== %dx%dx%d probe=%s" % (M, N, K, os.environ.get("B200AD_TC_PROBE", "0")), flush=True)
    D.matmul(a, b, trans_a=ta, trans_b=tb, out=out)
    torch.cuda.synchronize()


for probe in sys.argv[1:] or ["0"]:
    os.environ["B200AD_TC_PROBE"] = probe
    run(8192, 8192, 8192, tb=True)
    run(131072, 512, 512)
