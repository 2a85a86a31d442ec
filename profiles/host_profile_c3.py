"""cProfile of the host side of the config-3 layer call (MultiHeadAttention forward + backward, B = 16): where do the
~15 us of Python per launch go?"""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auto_differentiation_b200 import device as D                    # noqa: E402
from auto_differentiation_b200.layers import layers as Ls            # noqa: E402

rng = np.random.default_rng(1237)
B, S, d = 16, 512, 512
ws = [rng.standard_normal((d, d)) / np.sqrt(d) for _ in range(4)]
layer = Ls.MultiHeadAttention(d, d, 8, weights=ws)
g = torch.Generator(device="cuda").manual_seed(1237)
x = D.DeviceArray(torch.randn((B, S, d), generator=g, device="cuda"))
dout = D.DeviceArray(torch.randn((B, S, d), generator=g, device="cuda"))


def step():
    layer.forward(x, x, x)
    layer.backward(dout)


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
