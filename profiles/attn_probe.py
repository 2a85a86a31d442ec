"""Correctness + timing probe of the fused attention passes (b200ad_attn_fwd / b200ad_attn_bwd) on one B200.
    python profiles/attn_probe.py [quick]
Prints max errors vs fp64 NumPy for a ladder of shapes, then times the config-5 micro-batch shape
(256 sequences x 8 heads x 512 x 512 x 64) with CUDA events and reports algorithmic TFLOP/s
(fwd 4*S*S*dk flops per head, bwd 10*S*S*dk)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from auto_differentiation_b200 import device as D
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    timing_only = len(sys.argv) > 1 and sys.argv[1] in ("timing", "fwdonly")
    fwd_only = len(sys.argv) > 1 and sys.argv[1] == "fwdonly"

    def heads(a):
        B, h, S, dk = a.shape
        merged = D.asarray(np.ascontiguousarray(a.transpose(0, 2, 1, 3)).reshape(B, S, h * dk))
        return merged.reshape(B, S, h, dk).transpose(0, 2, 1, 3)

    def ref(q, k, v, do, scale):
        q, k, v, do = (a.astype(np.float64) for a in (q, k, v, do))
        s = scale * (q @ k.transpose(0, 1, 3, 2))
        e = np.exp(s - s.max(-1, keepdims=True))
        p = e / e.sum(-1, keepdims=True)
        o = p @ v
        dv = p.transpose(0, 1, 3, 2) @ do
        dp = do @ v.transpose(0, 1, 3, 2)
        ds = p * (dp - (dp * p).sum(-1, keepdims=True)) * scale
        return o, ds @ k, ds.transpose(0, 1, 3, 2) @ q, dv

    dk = 64
    scale = 1.0 / math.sqrt(dk)
    for (B, h, Sq, Skv) in [] if timing_only else [(1, 1, 128, 64), (1, 1, 128, 128), (1, 2, 256, 256), (2, 3, 100, 200), (4, 8, 512, 512),
                            (40, 8, 128, 192)]:
        rng = np.random.default_rng(Sq + Skv)
        q = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
        k = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
        v = rng.standard_normal((B, h, Skv, dk)).astype(np.float32)
        do = rng.standard_normal((B, h, Sq, dk)).astype(np.float32)
        Q, K, V, dO = (heads(a) for a in (q, k, v, do))
        O, lse = D.attn_fwd(Q, K, V, scale)
        torch.cuda.synchronize()
        o_ref, dq_ref, dk_ref, dv_ref = ref(q, k, v, do, scale)
        eo = np.abs(np.asarray(O) - o_ref).max() / max(1.0, np.abs(o_ref).max())
        print("B=%d h=%d Sq=%d Skv=%d  fwd err %.3e" % (B, h, Sq, Skv, eo), flush=True)
        dQ, dK, dV = D.attn_bwd(Q, K, V, O, dO, lse, scale)
        torch.cuda.synchronize()
        for name, g, r in (("dV", dV, dv_ref), ("dQ", dQ, dq_ref), ("dK", dK, dk_ref)):
            e = np.abs(np.asarray(g) - r).max() / max(1.0, np.abs(r).max())
            print("    %s err %.3e" % (name, e), flush=True)
    if quick:
        return
    # timing at the config-5 micro-batch shape
    B, h, S = 256, 8, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    bufs = [torch.randn((B, S, h * dk), device="cuda", generator=g) for _ in range(4)]
    Q, K, V, dO = (D.DeviceArray(b).reshape(B, S, h, dk).transpose(0, 2, 1, 3) for b in bufs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    O, lse = D.attn_fwd(Q, K, V, scale)
    if not fwd_only:
        D.attn_bwd(Q, K, V, O, dO, lse, scale)
    torch.cuda.synchronize()
    iters = 5
    for name, fn, flops in (("fwd", lambda: D.attn_fwd(Q, K, V, scale), 4.0 * B * h * S * S * dk),
                            ("bwd", lambda: D.attn_bwd(Q, K, V, O, dO, lse, scale), 10.0 * B * h * S * S * dk)):
        if fwd_only and name == "bwd":
            continue
        ms = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = float(np.median(ms))
        print("config-5 micro-batch %s: %.3f ms  %.1f algorithmic TFLOP/s" % (name, t, flops / t / 1e9), flush=True)


if __name__ == "__main__":
    main()
