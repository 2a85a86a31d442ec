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


# ---- achieved-error ledger: every assert_close records (test, what, max error, tolerance there, worst err/tol) -------
_LEDGER = []
_CURRENT = [""]


@pytest.fixture(autouse=True)
def _ledger_test_name(request):
    _CURRENT[0] = request.node.nodeid
    yield


def record_parity(what, max_err, tol_at_max, worst_ratio, rtol, atol, n):
    _LEDGER.append({"test": _CURRENT[0], "what": what, "max_err": float(max_err), "tol_at_max_err": float(tol_at_max),
                    "worst_err_over_tol": float(worst_ratio), "rtol": rtol, "atol": atol, "elements": int(n)})


def pytest_terminal_summary(terminalreporter):
    """Prints the margin of every parity assertion (achieved error / tolerance), worst first, and optionally dumps
    the full ledger as JSON (B200AD_PARITY_REPORT=<path>)."""
    if not _LEDGER:
        return
    worst = {}
    for e in _LEDGER:
        key = (e["test"], e["rtol"], e["atol"])
        if key not in worst or e["worst_err_over_tol"] > worst[key]["worst_err_over_tol"]:
            worst[key] = e
    rows = sorted(worst.values(), key=lambda e: -e["worst_err_over_tol"])
    tr = terminalreporter
    tr.write_line("")
    tr.write_line("parity margins: %d assertions in %d tests; worst err/tol per test (1.0 = at the tolerance):"
                  % (len(_LEDGER), len({e["test"] for e in _LEDGER})))
    for e in rows[:int(os.environ.get("B200AD_PARITY_ROWS", "25"))]:
        tr.write_line("  %5.3f  max_err %.2e (tol %.2e, rtol %g atol %g)  %s :: %s"
                      % (e["worst_err_over_tol"], e["max_err"], e["tol_at_max_err"], e["rtol"], e["atol"],
                         e["test"].split("/")[-1], e["what"]))
    path = os.environ.get("B200AD_PARITY_REPORT")
    if path:
        import json
        with open(path, "w") as f:
            json.dump(_LEDGER, f, indent=0)


def assert_close(got, ref, rtol=1e-5, atol=1e-6, what=""):
    """fp32 result vs fp64 reference: |got-ref| <= atol*max(1,max|ref|) + rtol*|ref| elementwise
    (north_star tolerance: rtol 1e-5 / atol 1e-6, the absolute part scaled to the tensor's magnitude).
    The achieved error and its margin to the tolerance go to the ledger printed at the end of the run."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s != %s" % (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    scale = max(1.0, float(np.max(np.abs(ref))))
    err = np.abs(got - ref)
    tol = atol * scale + rtol * np.abs(ref)
    bad = ~(err <= tol)   # NaN/inf anywhere is a failure
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = np.where(np.isfinite(err), err / tol, np.inf)
    imax = int(np.argmax(np.where(np.isfinite(err), err, np.inf)))
    record_parity(what, err.reshape(-1)[imax], tol.reshape(-1)[imax], float(np.max(ratio)), rtol, atol, ref.size)
    assert not bad.any(), "%s: %d/%d elements out of tolerance, max err %.3e (tol %.3e) at %s" % (
        what, int(bad.sum()), ref.size, float(err.max()), float(tol.reshape(-1)[int(np.argmax(err))]),
        np.unravel_index(int(np.argmax(err)), ref.shape))
