"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol the
header declares, and its host-only helpers (geometry, header parsing, tables) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT


@pytest.fixture(scope="module")
def L():
    from llcomp_b200.build import build
    build(tools=False)
    from llcomp_b200 import _capi
    return _capi.lib()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "llcomp_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(llcomp_b200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(L):
    from llcomp_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/llcomp_b200.h but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_capi.SIGNATURES) == names


def test_status_strings_match_reference_exceptions(L):
    assert L.llcomp_b200_status_string(1) == b"Invalid magic number"     # llcomp.hpp:466
    assert L.llcomp_b200_status_string(2) == b"Invalid exponent"         # llcomp.hpp:233
    assert L.llcomp_b200_status_string(0) == b"ok"


def test_model_tables_equal_oracle(L):
    O = oracle.lib()
    for s in range(128):
        e = L.llcomp_b200_debug_table(s)
        assert e & 0xFF == O.llo_state_probability(s)
        assert (e >> 8) & 0xFF == O.llo_next_state_mps(s)
        assert (e >> 16) & 0xFF == O.llo_next_state_lps(s)


def test_geometry_helpers(L):
    from llcomp_b200 import Geometry
    g = Geometry(4096, 4096, 3, 512, 512, 1)
    assert L.llcomp_b200_slice_count(C.byref(g)) == 64
    assert L.llcomp_b200_sample_count(C.byref(g)) == 4096 * 4096 * 3
    assert L.llcomp_b200_payload_capacity(C.byref(g)) == 2 * 4096 * 4096 * 3 + 384 * 64
    g = Geometry(1024, 1024, 3, 0, 0, 1024)
    assert L.llcomp_b200_slice_count(C.byref(g)) == 1024
    g = Geometry(600, 500, 3, 256, 128, 2)
    assert L.llcomp_b200_slice_count(C.byref(g)) == 3 * 4 * 2
    for bad in (Geometry(0, 4, 3, 0, 0, 1), Geometry(4, 4, 0, 0, 0, 1), Geometry(4, 4, 256, 0, 0, 1),
                Geometry(4, 4, 3, -1, 0, 1), Geometry(4, 4, 3, 0, 0, 0)):
        assert L.llcomp_b200_slice_count(C.byref(bad)) == 0


def test_peek_parses_both_headers(L):
    v = [C.c_int() for _ in range(5)]
    s = np.frombuffer(oracle.compress(oracle.generate(20, 10, 3, 2, 1)), np.uint8)
    assert L.llcomp_b200_peek(s.ctypes.data, s.size, *[C.byref(x) for x in v]) == 0
    assert [x.value for x in v] == [20, 10, 3, 20, 10]
    bad = np.array([0x77, 3, 1, 0, 1, 0, 0], np.uint8)
    assert L.llcomp_b200_peek(bad.ctypes.data, bad.size, *[C.byref(x) for x in v]) == 1
    short = np.array([0x79, 3, 1], np.uint8)
    assert L.llcomp_b200_peek(short.ctypes.data, short.size, *[C.byref(x) for x in v]) == 7
    import struct
    hdr = bytes([0xB2, 1, 3, 0]) + struct.pack("<5I", 600, 500, 256, 128, 12) + struct.pack("<12I", *([5] * 12))
    h = np.frombuffer(hdr, np.uint8)
    assert L.llcomp_b200_peek(h.ctypes.data, h.size, *[C.byref(x) for x in v]) == 0
    assert [x.value for x in v] == [600, 500, 3, 256, 128]
    hdr2 = bytes([0xB2, 1, 3, 0]) + struct.pack("<5I", 600, 500, 256, 128, 11) + struct.pack("<11I", *([5] * 11))
    h2 = np.frombuffer(hdr2, np.uint8)
    assert L.llcomp_b200_peek(h2.ctypes.data, h2.size, *[C.byref(x) for x in v]) == 3   # slice count mismatch


def test_no_cpu_fallback_without_a_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import llcomp_b200
    with pytest.raises(llcomp_b200.LlcompError):
        llcomp_b200.Codec(0)
    with pytest.raises(llcomp_b200.LlcompError):
        llcomp_b200.compressImage(np.zeros(12, np.uint8), 2, 2, 3)


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "llcomp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                with open(os.path.join(dirpath, f), errors="ignore") as fh:
                    t = fh.read()
                if re.search(r"import oracle|from oracle|llcomp_oracle|libllcomp_ref|-I\s*/root/reference|"
                             r"#include\s*[<\"][^>\"]*reference", t):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_unmodified_reference_tools_compile_against_the_host_header():
    """The drop-in claim of llcomp_b200/host/llcomp.hpp: the reference's own llcompc.cpp (:33) and llcompd.cpp (:26,
    structured binding of RawImage) compile and link UNCHANGED against it.  The sources are read from /root/reference
    at build time (nothing is copied); stb_image*.h, which the reference does not vendor, are the stubs in tests/stubs."""
    import os
    import subprocess
    if not os.path.exists("/root/reference/llcompc.cpp"):
        pytest.skip("/root/reference is not on this machine")
    host = os.path.join(ROOT, "llcomp_b200", "host")
    from llcomp_b200.build import build
    build()
    r = subprocess.run(["make", "-C", host, "-B", "ref_cli"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for n in ("llcompc", "llcompd"):
        assert os.access(os.path.join(host, "_ref_cli", n), os.X_OK)
