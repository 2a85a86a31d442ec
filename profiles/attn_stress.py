"""Stress of the fused attention passes: the same inputs through forward + backward many times must give bit-identical
results every time (the two MMA-issuing warps, the aliased score buffers and the spin-wait protocol leave no room for a
schedule-dependent result), at several shapes incl. ragged ones, interleaved with other work on the device."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auto_differentiation_b200 import device as D            # noqa: E402

g = torch.Generator(device="cuda").manual_seed(7)
bad = 0
for (B, h, Sq, Skv, dk, reps) in [(64, 8, 512, 512, 64, 60), (9, 4, 333, 200, 64, 150), (40, 8, 128, 128, 32, 150), (3, 2, 70, 700, 64, 150)]:
    bufs = [torch.randn((B, S, h * dk), device="cuda", generator=g) for S in (Sq, Skv, Skv, Sq)]
    Q, K, V, dO = (D.DeviceArray(b).reshape(B, b.shape[1], h, dk).transpose(0, 2, 1, 3) for b in bufs)
    scale = 1.0 / math.sqrt(dk)
    ref = None
    noise = torch.randn((4096, 4096), device="cuda", generator=g)
    for i in range(reps):
        if i % 3 == 1:
            noise = noise @ noise.T * 1e-4          # unrelated work before the passes: different cache / clock state
        O, lse = D.attn_fwd(Q, K, V, scale)
        dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
        cur = [t.t.clone() for t in (O, lse, dQ, dK, dV)]
        if ref is None:
            ref = cur
            continue
        for name, a, b_ in zip(("O", "lse", "dQ", "dK", "dV"), ref, cur):
            if not torch.equal(a, b_):
                bad += 1
                print("MISMATCH shape", (B, h, Sq, Skv, dk), "iteration", i, name, float((a - b_).abs().max()), flush=True)
    torch.cuda.synchronize()
    print("shape", (B, h, Sq, Skv, dk), "x", reps, "ok" if bad == 0 else "BAD", flush=True)
print("attention stress:", "bit-identical on every repetition" if bad == 0 else "%d mismatches" % bad)
sys.exit(1 if bad else 0)
