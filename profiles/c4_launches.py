"""Config 4 (tanh RNN LM, H=1024, T=256, N=64): two steps for an `ncu --metrics gpu__time_duration.sum` launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from auto_differentiation_b200.models.model import RnnLM
from auto_differentiation_b200.utils.optimizer import Adam

cell = sys.argv[1] if len(sys.argv) > 1 else "rnn"
rng = np.random.default_rng(1238)
N, T, V = 64, 256, 1000
xs = rng.integers(0, V, (N, T))
ts = np.roll(xs, -1, axis=1)
m, opt = RnnLM(vocab=V, wordvec=256, hidden=1024, cell=cell), Adam(lr=1e-3)
for _ in range(2):
    loss = m.forward(xs, ts)
    m.backward()
    opt.update(m.weights, m.grads)
torch.cuda.synchronize()
print("loss", float(loss))
