"""Times the persistent tanh-RNN kernels at the config-4 shape (H=1024, N=64, T=256): us per dependent phase, forward and
backward; with B200AD_RNN_SEQ_PROF=1 the library also prints per-segment cycle counters; B200AD_RNN_SEQ_NSPLIT=1 forces
the one-cluster-per-column-slice decomposition (64 CTAs instead of 128)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auto_differentiation_b200 as pkg                      # noqa: E402
from auto_differentiation_b200 import device as D            # noqa: E402

H, N, T = int(os.environ.get("RS_H", 1024)), int(os.environ.get("RS_N", 64)), int(os.environ.get("RS_T", 256))
rng = np.random.default_rng(0)
wh = pkg.asarray(rng.standard_normal((H, H)) / np.sqrt(H))
init = pkg.asarray(np.concatenate([np.zeros((1, N, H)), rng.standard_normal((T, N, H))], axis=0))
hs = D.empty((T + 1, N, H))
dhs = pkg.asarray(rng.standard_normal((N, T, H)) * 0.01)
print("supported:", D.rnn_seq_supported(N, H))
pre = init[1:].copy()
for name, fn in (("fwd", lambda: D.rnn_tanh_seq_fwd_(hs, pre, wh)), ("bwd", lambda: D.rnn_tanh_seq_bwd(dhs, hs, wh))):
    D.copy_into_(init, hs)
    fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        if name == "fwd":
            D.copy_into_(init, hs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    print("%s: %.3f ms per launch, %.2f us per phase (T=%d, N=%d, H=%d)" % (name, ms, ms * 1e3 / (T + (name == "bwd")), T, N, H))
