import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    """'cuda': the real libb200ad.so on a GPU (the parity tests proper, `-m gpu`).
    'emu':  tests/abi_emulator.py behind the same ctypes surface, to exercise the host-side logic on a
            CPU-only machine (`-m "not gpu"`)."""
    from auto_differentiation_b200 import _lib, device
    if request.param == "emu":
        from abi_emulator import Emulator
        monkeypatch.setattr(_lib, "_lib", Emulator())
        monkeypatch.setattr(device, "_DEVICE", "cpu")
        monkeypatch.setattr(device, "require_cuda", lambda: None)
        monkeypatch.setattr(device, "stream_ptr", lambda: 0)
        yield "emu"
    else:
        import torch
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        _lib.load()
        yield "cuda"


def assert_close(got, ref, rtol=1e-5, atol=1e-6, what=""):
    """fp32 result vs fp64 reference: |got-ref| <= atol*max(1,max|ref|) + rtol*|ref| elementwise
    (north_star tolerance: rtol 1e-5 / atol 1e-6, the absolute part scaled to the tensor's magnitude)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s != %s" % (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    scale = max(1.0, float(np.max(np.abs(ref))))
    err = np.abs(got - ref)
    tol = atol * scale + rtol * np.abs(ref)
    bad = ~(err <= tol)   # NaN/inf anywhere is a failure
    assert not bad.any(), "%s: %d/%d elements out of tolerance, max err %.3e (tol %.3e) at %s" % (
        what, int(bad.sum()), ref.size, float(err.max()), float(tol.reshape(-1)[int(np.argmax(err))]),
        np.unravel_index(int(np.argmax(err)), ref.shape))
