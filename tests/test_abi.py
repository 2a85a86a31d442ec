"""The C-ABI library loads and exports exactly the symbols include/b200ad.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200ad.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ad_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from auto_differentiation_b200 import _lib
    names = _declared()
    assert len(names) >= 35
    assert os.path.isfile(_lib.LIB_PATH), "build with `python -m auto_differentiation_b200.build`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libb200ad.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes prototypes and header disagree"
    assert _lib.load().b200ad_abi_version() == 1


def test_library_is_sm100a_and_has_no_torch_dependency():
    from auto_differentiation_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        arch = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
        assert "sm_100a" in arch


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import auto_differentiation_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.asarray(np.ones(3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "auto_differentiation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("np_oracle", "oracle") or "import" not in "".join(
                    l for l in src.splitlines() if "oracle" in l), "%s references the oracle" % f
