import sys, cProfile, pstats, io
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from auto_differentiation_b200.models.model import TapeMLP
from auto_differentiation_b200.utils.optimizer import Adam
rng = np.random.default_rng(0)
m = TapeMLP(d_in=784, d_hidden=256, d_out=10)
opt = Adam(lr=1e-3)
x = rng.standard_normal((64, 784)); t = rng.integers(0, 10, 64)
def step():
    l = m.forward(x, t); m.backward(); opt.update(m.weights, m.grads)
for _ in range(20): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
