"""In-tree build of the CUDA library (nvcc, sm_100a only).  No GPU is needed to compile."""
from __future__ import annotations

import glob
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib", "libllcomp_b200.so")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(PKG, "csrc", "*")) + [os.path.join(ROOT, "include", "llcomp_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, tools: bool = True) -> str:
    """Compile llcomp_b200/lib/libllcomp_b200.so (and the llcompc/llcompd tools) if out of date."""
    if force or _stale():
        if not os.path.exists("/usr/local/cuda/bin/nvcc"):
            raise RuntimeError("nvcc not found and llcomp_b200/lib/libllcomp_b200.so is missing or stale")
        r = subprocess.run(["make", "-C", ROOT, LIB[len(ROOT) + 1:]], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libllcomp_b200.so failed:\n" + r.stdout + r.stderr)
    if tools and os.path.exists(os.path.join(PKG, "host", "Makefile")):
        r = subprocess.run(["make", "-C", os.path.join(PKG, "host")], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building llcompc/llcompd failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
