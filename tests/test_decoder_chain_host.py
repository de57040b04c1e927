"""The decoder's slice chain (llcomp_b200/csrc/decoder_chain.cuh) compiled for the host and compared with the oracle:
the same source runs on the device in k_slice_decoder_chain, so its logic (the scaled range/low arithmetic, the
two-part context preparation, row forwarding, ring refills, borders) is checked here without a GPU.  The product never
runs this build: tests/host/chain_host.cpp is a test harness."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def chain():
    out = os.path.join(HERE, "host", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libchain_host.so")
    src = os.path.join(HERE, "host", "chain_host.cpp")
    hdr = os.path.join(HERE, "..", "llcomp_b200", "csrc", "decoder_chain.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                        "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.chain_decode_tile_v.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int]

    def dec(payload: bytes, w: int, h: int, c: int):
        """Both forms of the decision code (Chain's kV 0 and 4) must agree; returns what they return."""
        buf = np.frombuffer(payload, dtype=np.uint8).copy() if len(payload) else np.zeros(1, np.uint8)
        res = []
        for variant in (0, 4):
            out_px = np.zeros((h, w, c), np.uint8)
            rc = lib.chain_decode_tile_v(buf.ctypes.data, len(payload), w, h, c, out_px.ctypes.data, w * c, variant)
            res.append((rc, out_px))
        assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all(), "the two forms of the chain disagree"
        return res[0]
    return dec


CASES = [(64, 48, 3, 4), (1, 1, 3, 4), (1, 9, 3, 0), (9, 1, 3, 8), (2, 2, 1, 4), (33, 17, 1, 16), (40, 30, 4, 4),
         (17, 40, 2, 32), (256, 64, 3, 0), (128, 128, 3, -1), (700, 5, 3, 4), (96, 96, 1, -1), (300, 20, 4, 64),
         (1024, 6, 3, 128), (3, 3, 2, 255), (512, 512, 3, 4)]


@pytest.mark.parametrize("w,h,c,noise", CASES)
def test_chain_decodes_oracle_payloads(chain, w, h, c, noise):
    img = oracle.generate(w, h, c, noise, 77 + w + h)
    payload = oracle.encode_tile(img, 0, 0, w, h)
    rc, out = chain(payload, w, h, c)
    assert rc == 0 and (out == img).all()


@pytest.mark.parametrize("w,h,c,noise", CASES[:13])
def test_chain_follows_the_oracle_on_damaged_streams(chain, w, h, c, noise):
    """Truncated, padded and random payloads decode to whatever the reference algorithm makes of them (zero fill past
    the end, llcomp.hpp:475-479; int16 wrap of wild samples; "Invalid exponent", :232)."""
    rng = np.random.default_rng(w * 131 + h)
    img = oracle.generate(w, h, c, noise, 5 + w)
    p = oracle.encode_tile(img, 0, 0, w, h)
    variants = [p[: len(p) // 2], p[:1], b"", p + bytes(rng.integers(0, 256, 40, dtype=np.uint8)),
                bytes(rng.integers(0, 256, max(8, len(p)), dtype=np.uint8)), b"\xff" * max(8, len(p)), b"\x00" * 16]
    for q in variants:
        try:
            want, wrc = oracle.decode_tile(q, w, h, c), 0
        except oracle.OracleError:
            want, wrc = None, 2
        rc, out = chain(q, w, h, c)
        assert rc == wrc
        if want is not None:
            assert (out == want).all()


def test_chain_reports_a_runaway_exponent(chain):
    """A nonzero flag followed by all ones: the exponent runs past 31 -> "Invalid exponent" (llcomp.hpp:232-233); the
    chain flags it and stops at the next pixel.  Same payload as tests/test_gpu_parity.py::test_error_behaviour."""
    bad = bytes([2, 250]) + b"\xff" * 60
    hits = 0
    for c, w, h in ((1, 4, 4), (1, 9, 2), (2, 4, 4), (3, 8, 5), (4, 3, 3)):
        try:
            oracle.decode_tile(bad, w, h, c)
            want = 0
        except oracle.OracleError as e:
            assert e.code == 2
            want, hits = 2, hits + 1
        rc, _ = chain(bad, w, h, c)
        assert rc == want, (c, w, h)
    assert hits >= 1
