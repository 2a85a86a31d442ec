"""Builds libb200ad.so in-tree with nvcc for sm_100a (cross-compiles without a GPU):
    python -m auto_differentiation_b200.build
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, jobs=None):
    jobs = jobs or max(1, (os.cpu_count() or 2))
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j%d" % jobs]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libb200ad.so failed (see output above)")
    return os.path.join(HERE, "libb200ad.so")


if __name__ == "__main__":
    print(build(verbose=True))
