"""The front end's per-sample arithmetic (llcomp_b200/csrc/sample.cuh) compiled for the host: the multiply-add form of
the record packing is checked exhaustively against the plain packing, and code_sample against the oracle's front end on
random neighbourhoods.  The product only runs this code on the device; tests/host/sample_host.cpp is a test harness."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "host", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsample_host.so")
    src = os.path.join(HERE, "host", "sample_host.cpp")
    hdr = os.path.join(HERE, "..", "llcomp_b200", "csrc", "sample.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-w", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                        "-o", so, src], check=True)
    L = C.CDLL(so)
    L.sample_check_record_of.restype = C.c_longlong
    L.sample_code.restype = C.c_uint32
    L.sample_code.argtypes = [C.c_int] * 7
    return L


def test_record_packing_equals_plain_packing_everywhere(lib):
    """record_of(hash, diff) == pack(|hash|, fold(diff)) for all 15,851 x 2,047 pairs (llcomp.hpp:431-436)."""
    assert lib.sample_check_record_of() == 0


def test_code_sample_equals_oracle_on_random_neighbourhoods(lib):
    """An interior pixel of a 4x3 gray image has all six neighbours (llcomp.hpp:417-422): its record from code_sample must
    be the oracle's."""
    rng = np.random.default_rng(11)
    for trial in range(400):
        spread = int(rng.choice([2, 8, 40, 255]))
        base = int(rng.integers(0, 256 - spread)) if spread < 255 else 0
        img = (base + rng.integers(0, spread + 1, size=(3, 4, 1))).astype(np.uint8)
        sym = oracle.frontend(img).reshape(3, 4)
        p = img[:, :, 0].astype(int)
        got = lib.sample_code(p[2][2], p[2][1], p[2][0], p[1][1], p[1][2], p[1][3], p[0][2])
        assert got == int(sym[2][2]) & 0xFFFFFFFF, (trial, p.tolist())
