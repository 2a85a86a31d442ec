"""Diagnostic: where the fp16-split GEMM (csrc/gemm_h16.cu) spends its time.  Runs the config-5 projection and 8192^3 under the
B200AD_H16_PROBE bits (1 = no output stores, 4 = no operand split: kernel alone, 8 = MMA-warp wait counters, 16 = plane
path instead of the in-kernel conversion of A), each setting in
its own process (the library reads the variable once).
    python profiles/gemm_h16_prof.py            # all settings
    python profiles/gemm_h16_prof.py run        # one setting (B200AD_H16_PROBE from the environment)
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    from auto_differentiation_b200 import device as D
    probe = os.environ.get("B200AD_H16_PROBE", "0")
    for (M, N, K, tb) in ((131072, 512, 512, False), (131072, 512, 512, True), (8192, 8192, 8192, True)):
        a = D.DeviceArray(torch.randn((M, K), device="cuda"))
        b = D.DeviceArray(torch.randn((N, K) if tb else (K, N), device="cuda"))
        out = D.empty((M, N))
        for _ in range(3):
            D.matmul(a, b, trans_b=tb, out=out, precision="f16x3")
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); D.matmul(a, b, trans_b=tb, out=out, precision="f16x3"); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        print("probe=%s  %dx%dx%d tb=%d: %.3f ms  %.1f TFLOP/s" % (probe, M, N, K, tb, ms[2], 2.0 * M * N * K / ms[2] / 1e9), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "run":
    run()
else:
    for probe in ("0", "16", "4", "5", "12"):
        env = dict(os.environ, B200AD_H16_PROBE=probe)
        subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=env, timeout=300)
