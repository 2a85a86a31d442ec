"""Launches each tensor-core kernel family once or twice for an `ncu --set full` capture:
the generation-3 GEMM on a projection shape, the fused attention forward (tf32 split) and backward (fp16 split) passes at
64 sequences, and the persistent tanh-RNN kernels at T = 64.

    python profiles/ncu_targets.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'gemm_tc3|attn_tc_kernel|attn_h16_kernel|rnn_seq_kernel' \
        -o gpurun_out/r2_tensor_kernels python profiles/ncu_targets.py
    ncu -i gpurun_out/r2_tensor_kernels.ncu-rep --page raw --csv > profiles/r2_ncu_tensor_kernels.csv      # here, no GPU needed
"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auto_differentiation_b200 as pkg                      # noqa: E402
from auto_differentiation_b200 import device as D            # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
rand = lambda *s: D.DeviceArray(torch.randn(s, generator=g, device="cuda"))
# 1. projection-shaped GEMM (gemm_tc3_kernel): (65536 x 512) @ (512 x 512), and its split-K weight gradient
x, w = rand(65536, 512), rand(512, 512)
y = D.matmul(x, w)
D.matmul(x, y, trans_a=True)
# 2. attention at 64 sequences x 8 heads x 512 x 512 x 64
B, h, S, dk = 64, 8, 512, 64
bufs = [torch.randn((B, S, h * dk), device="cuda", generator=g) for _ in range(4)]
Q, K, V, dO = (D.DeviceArray(b).reshape(B, S, h, dk).transpose(0, 2, 1, 3) for b in bufs)
O, lse = D.attn_fwd(Q, K, V, 1.0 / math.sqrt(dk))
D.attn_bwd(Q, K, V, O, dO, lse, 1.0 / math.sqrt(dk))
# 3. persistent tanh-RNN, H = 1024, N = 64, T = 64
H, N, T = 1024, 64, 64
wh = rand(H, H)
D.axpby_(wh, wh, 1.0 / math.sqrt(H), 0.0)
hs = D.zeros((T + 1, N, H))
pre = rand(T, N, H)
D.rnn_tanh_seq_fwd_(hs, pre, wh)
D.rnn_tanh_seq_bwd(rand(N, T, H), hs, wh)
torch.cuda.synchronize()
print("ok")
