"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference in place from /root/reference.

Only usable in the build container (``/root/reference`` does not exist on the GPU box).
It is used by ``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/`` and by the ``-m "not gpu"`` tests that pin ``oracle/np_oracle.py``
against the live reference when it is present.  Nothing under ``oracle/`` may be imported
by the product package (``auto_differentiation_b200``).

Recipe follows SURVEY.md Appendix A:
  * ``auto_grad/tgrad.py`` imports ``server``, ``graphviz``, ``matplotlib.pyplot``
    (tgrad.py:5,7,11) — only used by visualisation code → inert stub modules.
  * it must be registered as module ``ad`` (layers/layers.py:11 ``import ad``).
  * ``layers/layers.py:6`` binds ``cupy as np`` → alias cupy to numpy; ``tensorboardX`` stub (:10).
  * ``utils/optimizer.py`` / ``utils/Train.py`` import with /root/reference on sys.path.
No reference source is copied; the files are executed where they lie.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200AD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "auto_grad", "tgrad.py"))


_cache = {}


def _stub(name):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
    return sys.modules[name]


def load_tape():
    """Return the reference tape module (auto_grad/tgrad.py) registered as ``ad``."""
    if "ad" in _cache:
        return _cache["ad"]
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    have_mpl = True
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        have_mpl = False
    for n in ("server", "graphviz"):
        _stub(n)
    if not have_mpl:
        m = _stub("matplotlib")
        p = _stub("matplotlib.pyplot")
        m.pyplot = p
    spec = importlib.util.spec_from_file_location(
        "ad", os.path.join(REFERENCE_ROOT, "auto_grad", "tgrad.py"))
    ad = importlib.util.module_from_spec(spec)
    sys.modules["ad"] = ad
    spec.loader.exec_module(ad)
    _cache["ad"] = ad
    return ad


def load_layers():
    """Return the reference layers module (layers/layers.py) running on NumPy."""
    if "layers" in _cache:
        return _cache["layers"]
    import numpy
    load_tape()
    prev_cupy = sys.modules.get("cupy")
    sys.modules["cupy"] = numpy
    tb = _stub("tensorboardX")
    if not hasattr(tb, "SummaryWriter"):
        tb.SummaryWriter = object
    state = numpy.random.get_state()  # layers.py:15 reseeds the global RNG at import
    try:
        spec = importlib.util.spec_from_file_location(
            "ref_layers", os.path.join(REFERENCE_ROOT, "layers", "layers.py"))
        L = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(L)
    finally:
        numpy.random.set_state(state)
        if prev_cupy is None:
            sys.modules.pop("cupy", None)
        else:
            sys.modules["cupy"] = prev_cupy
    _cache["layers"] = L
    return L


def load_utils():
    """Return (optimizer module, Train module, activation_func module) of the reference."""
    if "utils" in _cache:
        return _cache["utils"]
    import numpy
    load_tape()
    out = []
    prev_cupy = sys.modules.get("cupy")
    sys.modules["cupy"] = numpy
    try:
        for fname, modname in (("optimizer.py", "ref_utils_optimizer"),
                               ("useful_func.py", "ref_utils_useful_func"),
                               ("activation_func.py", "ref_utils_activation_func")):
            spec = importlib.util.spec_from_file_location(
                modname, os.path.join(REFERENCE_ROOT, "utils", fname))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            out.append(m)
    finally:
        if prev_cupy is None:
            sys.modules.pop("cupy", None)
        else:
            sys.modules["cupy"] = prev_cupy
    _cache["utils"] = tuple(out)
    return _cache["utils"]
