"""ncu target: the bf16 GEMM variant at 8192^3 (conversion + kernel), three calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auto_differentiation_b200 import device as D
a = D.DeviceArray(torch.randn((8192, 8192), device="cuda"))
b = D.DeviceArray(torch.randn((8192, 8192), device="cuda"))
for _ in range(3):
    D.matmul(a, b, trans_b=True, precision="bf16")
torch.cuda.synchronize()
print("ok")
