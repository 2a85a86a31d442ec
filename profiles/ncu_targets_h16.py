"""Launches the fp16-split GEMM (csrc/gemm_h16.cu) for an `ncu --set full` capture: the shared-A Q / K / V projection of the
config-5 micro-batch (131072 x 512 x 512, three products on one operand split), a config-4 gate product (K = 256), 4096^3.
    python profiles/ncu_targets_h16.py > gpurun_out/plain_h16.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'gemm_h16_kernel|h16_split|h16_rowscale' -o gpurun_out/r2b_h16 python profiles/ncu_targets_h16.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auto_differentiation_b200 import device as D            # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
rand = lambda *s: D.DeviceArray(torch.randn(s, generator=g, device="cuda"))
x = rand(131072, 512)
ws = [rand(512, 512) for _ in range(3)]
D.matmul_shared_a(x, ws, precision="f16x3")
D.matmul(x, ws[0], precision="f16x3", trans_b=True)          # single projection: in-kernel conversion of A (gemm_h16_kernel<true>)
D.matmul(rand(32768, 256), rand(256, 4096), precision="f16x3")
D.matmul(rand(4096, 4096), rand(4096, 4096), precision="f16x3", trans_b=True)
torch.cuda.synchronize()
print("ok")
