"""GPU probe of the tcgen05 GEMM: correctness vs fp64 and timing, case by case.
    python profiles/gemm_tc_probe.py            # runs every group in its own subprocess (a trap only kills its group)
    python profiles/gemm_tc_probe.py GROUP      # one group in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["basic", "layouts", "x3", "batched", "epilogue", "perf", "chunks", "ceiling", "atmem", "mmarate", "bf16"]


def run_group(group):
    import numpy as np
    import torch
    from auto_differentiation_b200 import _lib as L, device as D
    rng = np.random.default_rng(0)

    def case(name, M, N, K, ta=False, tb=False, prec="tf32", batch=None, alpha=1.0, beta=0.0, bias=False, act=0,
             time_it=False):
        L.load().b200ad_gemm_reload_env()      # the library caches its B200AD_TC_* knobs; the groups below flip them
        try:
            bs = batch or ()
            A = rng.standard_normal(bs + ((K, M) if ta else (M, K)))
            B = rng.standard_normal(bs + ((N, K) if tb else (K, N)))
            a, b = D.asarray(A), D.asarray(B)
            A32, B32 = A.astype(np.float32).astype(np.float64), B.astype(np.float32).astype(np.float64)
            Aop = np.swapaxes(A32, -1, -2) if ta else A32
            Bop = np.swapaxes(B32, -1, -2) if tb else B32
            ref = alpha * (Aop @ Bop)
            c0 = rng.standard_normal(ref.shape)
            out = D.asarray(c0) if beta != 0 else None
            bv = rng.standard_normal(N) if bias else None
            if beta != 0:
                ref = ref + beta * c0.astype(np.float32)
            if bias:
                ref = ref + bv
            if act == 1:
                ref = np.tanh(ref)
            elif act == 3:
                ref = np.maximum(ref, 0)
            got = D.matmul(a, b, out=out, alpha=alpha, beta=beta, bias=D.asarray(bv) if bias else None, act=act,
                           precision=prec, trans_a=ta, trans_b=tb)
            torch.cuda.synchronize()
            g = np.asarray(got, dtype=np.float64)
            scale = np.sqrt(K) * 1.0
            err = np.abs(g - ref).max() / scale          # error relative to the |a||b| sqrt(K) magnitude
            msg = "%-34s M=%-6d N=%-6d K=%-7d ta=%d tb=%d %-7s err/sqrtK=%.3e" % (name, M, N, K, ta, tb, prec, err)
            if time_it:
                for _ in range(3):
                    D.matmul(a, b, precision=prec, trans_a=ta, trans_b=tb)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record()
                for _ in range(reps):
                    D.matmul(a, b, precision=prec, trans_a=ta, trans_b=tb)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                nb = int(np.prod(bs)) if bs else 1
                msg += "  %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K * nb / ms / 1e9)
            print(msg, flush=True)
        except Exception as e:  # noqa: BLE001
            print("%-34s M=%d N=%d K=%d ta=%d tb=%d %s FAILED: %s" % (name, M, N, K, ta, tb, prec, str(e)[:200]), flush=True)

    if group == "basic":
        case("one tile one kblock", 128, 256, 32, tb=True)
        case("one tile 2 kblocks", 128, 256, 64, tb=True)
        case("one tile K=1024", 128, 256, 1024, tb=True)
        case("BN=128", 128, 128, 256, tb=True)
        case("BN=64", 128, 64, 256, tb=True)
        case("4x2 tiles", 512, 512, 512, tb=True)
        case("tails", 200, 300, 100, tb=True)
        case("many tiles (persistent loop)", 4096, 2048, 256, tb=True)
        case("fp32 simt reference", 200, 300, 100, tb=True, prec="fp32")
    elif group == "layouts":
        for ta in (False, True):
            for tb in (False, True):
                case("layout", 256, 512, 256, ta=ta, tb=tb)
                case("layout tails", 200, 300, 100, ta=ta, tb=tb)
    elif group == "x3":
        for ta in (False, True):
            for tb in (False, True):
                case("x3 layout", 256, 512, 256, ta=ta, tb=tb, prec="tf32x3")
        case("x3 tails", 200, 300, 1000, prec="tf32x3")
        case("x3 K=784 (C1)", 64, 256, 784, prec="tf32x3")
        case("x3 deep K", 512, 512, 16384, ta=True, prec="tf32x3")
        os.environ["B200AD_TC_MASK_HI"] = "0"
        case("x3 without hi masking", 256, 512, 256, tb=True, prec="tf32x3")
        os.environ["B200AD_TC_MASK_HI"] = "1"
    elif group == "batched":
        case("batched 2x3", 128, 256, 64, tb=True, batch=(2, 3))
        case("batched attention QK^T", 512, 512, 64, tb=True, batch=(2, 8), prec="tf32x3")
        case("batched attention PV", 512, 64, 512, batch=(2, 8), prec="tf32x3")
        case("batched TN", 64, 512, 512, ta=True, batch=(2, 8), prec="tf32x3")
        case("split-K dW", 512, 512, 32768, ta=True, prec="tf32x3")
        case("split-K dW tf32", 512, 512, 32768, ta=True)
        case("split-K tails", 200, 72, 5000, ta=True, prec="tf32x3")
    elif group == "epilogue":
        case("alpha", 256, 256, 128, prec="tf32x3", alpha=0.125)
        case("alpha beta", 256, 256, 128, prec="tf32x3", alpha=0.5, beta=2.0)
        case("bias tanh", 256, 256, 128, prec="tf32x3", bias=True, act=1)
        case("bias relu tails", 200, 100, 96, prec="tf32x3", bias=True, act=3)
        case("split-K beta bias", 256, 256, 8192, ta=True, prec="tf32x3", beta=1.0, bias=True)
    elif group == "atmem":
        os.environ["B200AD_TC_BN"] = "64"
        for at in ("1", "0"):
            os.environ["B200AD_TC_ATMEM"] = at
            print("-- BN=64 forced, A-from-TMEM = %s" % at, flush=True)
            for ta in (False, True):
                for tb in (False, True):
                    case("layout", 256, 192, 256, ta=ta, tb=tb, prec="tf32x3")
            case("tails", 200, 60, 1000, prec="tf32x3")
            case("tails TN", 200, 60, 1000, ta=True, prec="tf32x3")
            case("C5 PV", 512, 64, 512, batch=(256, 8), prec="tf32x3", time_it=True)
            case("C5 dV (TN)", 512, 64, 512, ta=True, batch=(256, 8), prec="tf32x3", time_it=True)
            case("square 4096 BN64", 4096, 4096, 4096, tb=True, prec="tf32x3", time_it=True)
        os.environ.pop("B200AD_TC_BN"); os.environ.pop("B200AD_TC_ATMEM")
    elif group == "bf16":
        # bf16 variant: error vs fp64 is dominated by the bf16 rounding of the inputs (~2^-9 relative per operand)
        for ta in (False, True):
            for tb in (False, True):
                case("bf16 layout", 256, 512, 256, ta=ta, tb=tb, prec="bf16")
        case("bf16 tails", 200, 300, 1000, prec="bf16")
        case("bf16 K=784 (C1)", 64, 256, 784, prec="bf16")
        case("bf16 batched", 512, 512, 64, tb=True, batch=(4, 8), prec="bf16")
        case("bf16 split-K dW", 512, 512, 32768, ta=True, prec="bf16")
        case("bf16 bias tanh", 256, 256, 128, prec="bf16", bias=True, act=1)
        case("bf16 beta", 256, 256, 512, prec="bf16", beta=1.0)
        for shape in ((4096, 4096, 4096), (8192, 8192, 8192)):
            case("bf16 square (incl. conversion)", *shape, tb=True, prec="bf16", time_it=True)
        case("bf16 C5 projection (incl. conversion)", 131072, 512, 512, prec="bf16", time_it=True)
        case("bf16 C5 dW (incl. conversion)", 512, 512, 131072, ta=True, prec="bf16", time_it=True)
        os.environ["B200AD_BF16_SKIP_CVT"] = "1"
        print("-- tensor-core kernel alone (conversion skipped: operands are whatever the workspace holds)", flush=True)
        for shape in ((4096, 4096, 4096), (8192, 8192, 8192)):
            case("bf16 square kernel-only", *shape, tb=True, prec="bf16", time_it=True)
        case("bf16 C5 projection kernel-only", 131072, 512, 512, prec="bf16", time_it=True)
        case("bf16 C5 dW kernel-only", 512, 512, 131072, ta=True, prec="bf16", time_it=True)
        os.environ.pop("B200AD_BF16_SKIP_CVT")
    elif group == "mmarate":
        # pure MMA + epilogue (probe 3: no TMA, no split; results are garbage): how does the tcgen05.mma rate depend on N?
        # per-CTA MMA count for 8192^3 x3: tiles(64 x 8192/BN) / 148 CTAs x 256 k-blocks x 12 MMAs
        os.environ["B200AD_TC_CHUNK_KB"] = "100000"
        os.environ["B200AD_TC_PROBE"] = "3"
        for bn in ("64", "128", "256"):
            os.environ["B200AD_TC_BN"] = bn
            for at in ("0", "1"):
                os.environ["B200AD_TC_ATMEM"] = at
                case("BN=%s atmem=%s x3" % (bn, at), 8192, 8192, 8192, tb=True, prec="tf32x3", time_it=True)
            case("BN=%s tf32 1-pass" % bn, 8192, 8192, 8192, tb=True, prec="tf32", time_it=True)
        for k in ("B200AD_TC_CHUNK_KB", "B200AD_TC_PROBE", "B200AD_TC_BN", "B200AD_TC_ATMEM"):
            os.environ.pop(k)
    elif group == "ceiling":
        os.environ["B200AD_TC_CHUNK_KB"] = "100000"
        print("-- probe bits: 1 = no TMA, 2 = no split, 3 = pure MMA + epilogue (results are garbage, only the rate matters)", flush=True)
        for pair in ("1", "0"):
            os.environ["B200AD_TC_PAIR"] = pair
            for probe in ("0", "1", "2", "3"):
                os.environ["B200AD_TC_PROBE"] = probe
                case("pair=%s probe=%s x3" % (pair, probe), 8192, 8192, 8192, tb=True, prec="tf32x3", time_it=True)
        os.environ["B200AD_TC_PAIR"] = "1"
        for probe in ("0", "1", "2", "3"):
            os.environ["B200AD_TC_PROBE"] = probe
            case("pair=1 probe=%s projection" % probe, 131072, 512, 512, prec="tf32x3", time_it=True)
        os.environ.pop("B200AD_TC_PROBE"); os.environ.pop("B200AD_TC_CHUNK_KB"); os.environ.pop("B200AD_TC_PAIR")
    elif group == "chunks":
        for ck in ("2", "4", "8", "16"):
            os.environ["B200AD_TC_CHUNK_KB"] = ck
            print("-- chunk_kb = %s" % ck, flush=True)
            case("K=256", 512, 512, 256, tb=True, prec="tf32x3")
            case("K=784", 64, 256, 784, prec="tf32x3")
            case("square 4096", 4096, 4096, 4096, tb=True, prec="tf32x3", time_it=True)
            case("C5 projection", 131072, 512, 512, prec="tf32x3", time_it=True)
            case("C5 dW", 512, 512, 131072, ta=True, prec="tf32x3", time_it=True)
        os.environ.pop("B200AD_TC_CHUNK_KB")
        for bn in ("128",):
            os.environ["B200AD_TC_BN"] = bn
            print("-- BN = %s" % bn, flush=True)
            case("square 4096", 4096, 4096, 4096, tb=True, prec="tf32x3", time_it=True)
            case("C5 projection", 131072, 512, 512, prec="tf32x3", time_it=True)
            case("C5 QK^T", 512, 512, 64, tb=True, batch=(256, 8), prec="tf32x3", time_it=True)
        os.environ.pop("B200AD_TC_BN")
    elif group == "perf":
        for prec in ("tf32", "tf32x3", "fp32"):
            case("square 4096", 4096, 4096, 4096, tb=True, prec=prec, time_it=True)
        case("square 8192", 8192, 8192, 8192, tb=True, prec="tf32", time_it=True)
        case("square 8192", 8192, 8192, 8192, tb=True, prec="tf32x3", time_it=True)
        case("square 8192 NN", 8192, 8192, 8192, prec="tf32x3", time_it=True)
        case("C5 projection", 131072, 512, 512, prec="tf32x3", time_it=True)
        case("C5 dW", 512, 512, 131072, ta=True, prec="tf32x3", time_it=True)
        case("C5 dX", 131072, 512, 512, tb=True, prec="tf32x3", time_it=True)
        case("C5 QK^T", 512, 512, 64, tb=True, batch=(256, 8), prec="tf32x3", time_it=True)
        case("C5 PV", 512, 64, 512, batch=(256, 8), prec="tf32x3", time_it=True)
        case("C5 dP^T..", 512, 64, 512, ta=True, batch=(256, 8), prec="tf32x3", time_it=True)
        case("C4 gates", 16384, 4096, 256, prec="tf32x3", time_it=True)
        case("C4 recurrent step", 64, 1024, 1024, prec="tf32x3", time_it=True)
        case("C4 affine", 16384, 1000, 1024, prec="tf32x3", time_it=True)


def run_chunks():
    pass


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_group(sys.argv[1])
    else:
        for g in GROUPS:
            print("==== group %s ====" % g, flush=True)
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), g], timeout=300)
                print("---- group %s exit %d (%.1fs)" % (g, r.returncode, time.time() - t0), flush=True)
            except subprocess.TimeoutExpired:
                print("---- group %s TIMEOUT" % g, flush=True)
