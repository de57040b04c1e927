"""ctypes doorway to the CPU checkers in oracle/.

TEST INFRASTRUCTURE ONLY.  Import this from tests/, from
``__graft_entry__.smoke()`` and from the cpu_baseline / ``--impl reference``
legs of ``bench.py`` -- never from ``llcomp_b200``.

Two libraries live here:

* ``libllcomp_oracle.so`` -- plain-C restatement (``llcomp_oracle.c``) of
  ``/root/reference/llcomp.hpp``; defined for every input.
* ``_ref/libllcomp_ref.so`` -- the unmodified reference header behind a C shim
  (``ref_shim.cpp``).  Undefined (heap overflow / OOB) when the stream is longer
  than the raw image or when decoding ``channels < 3``; ``ref_*`` helpers refuse
  those inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libllcomp_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libllcomp_ref.so")

STATUS = {0: "ok", 1: "Invalid magic number", 2: "Invalid exponent", 3: "nomem", 4: "bad argument"}

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> None:
    """Compile the checkers (gcc only; no GPU needed)."""
    if force or not os.path.exists(_ORACLE_SO) or (
        os.path.exists("/root/reference/llcomp.hpp") and not os.path.exists(_REF_SO)
    ):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_ORACLE_SO)
        L.llo_frontend_tile.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.llo_encode_tile.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.llo_encode_symbols.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.llo_decode_tile.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.llo_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(_u8p), C.POINTER(C.c_size_t)]
        L.llo_decompress.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_u8p),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.llo_count_bins.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.llo_count_bins.restype = C.c_uint64
        L.llo_free.argtypes = [C.c_void_p]
        L.llo_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.llo_fnv1a64.restype = C.c_uint64
        L.llo_generate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.llo_binarize.argtypes = [C.c_int, C.c_void_p]
        L.llo_compress_batch_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.llo_compress_batch_mt.restype = C.c_uint64
        L.llo_decompress_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.llo_decompress_batch_mt.restype = C.c_uint64
        _lib = L
    return _lib


def have_ref() -> bool:
    build()
    return os.path.exists(_REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libllcomp_ref.so not built (no /root/reference here)")
        R = C.CDLL(_REF_SO)
        R.ref_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        R.ref_compress.restype = C.c_size_t
        R.ref_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        R.ref_binarize.argtypes = [C.c_int, C.c_void_p]
        R.ref_compress_batch_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        R.ref_compress_batch_mt.restype = C.c_uint64
        R.ref_decompress_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        R.ref_decompress_batch_mt.restype = C.c_uint64
        _ref = R
    return _ref


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(STATUS.get(code, f"status {code}"))
        self.code = code


def _img(px) -> np.ndarray:
    a = np.ascontiguousarray(px, dtype=np.uint8)
    if a.ndim != 3:
        raise ValueError("expected an HxWxC uint8 array")
    return a


def _take(ptr, n) -> bytes:
    try:
        return C.string_at(ptr, n)
    finally:
        lib().llo_free(ptr)


# ----- restatement ---------------------------------------------------------
def generate(w: int, h: int, c: int, noise: int, seed: int = 1234) -> np.ndarray:
    """Synthetic generator G of SURVEY.md appendix C; noise<0 = uniform random bytes."""
    a = np.empty((h, w, c), dtype=np.uint8)
    lib().llo_generate(a.ctypes.data, w, h, c, noise, seed)
    return a


def fnv1a64(data: bytes) -> int:
    b = np.frombuffer(data, dtype=np.uint8)
    return int(lib().llo_fnv1a64(b.ctypes.data, b.size))


def tile_view(img: np.ndarray, x0: int, y0: int, tw: int, th: int):
    """(pointer, pitch) of a tile inside a contiguous HxWxC image."""
    h, w, c = img.shape
    return img.ctypes.data + (y0 * w + x0) * c, w * c


def frontend(px, x0=0, y0=0, tw=None, th=None) -> np.ndarray:
    a = _img(px)
    h, w, c = a.shape
    tw = w - x0 if tw is None else tw
    th = h - y0 if th is None else th
    out = np.empty(tw * th * c, dtype=np.uint32)
    p, pitch = tile_view(a, x0, y0, tw, th)
    rc = lib().llo_frontend_tile(p, pitch, tw, th, c, out.ctypes.data)
    if rc:
        raise OracleError(rc)
    return out


def encode_tile(px, x0=0, y0=0, tw=None, th=None) -> bytes:
    """Headerless payload == reference compressImage(tile)[6:]."""
    a = _img(px)
    h, w, c = a.shape
    tw = w - x0 if tw is None else tw
    th = h - y0 if th is None else th
    p, pitch = tile_view(a, x0, y0, tw, th)
    out, n = _u8p(), C.c_size_t()
    rc = lib().llo_encode_tile(p, pitch, tw, th, c, C.byref(out), C.byref(n))
    if rc:
        raise OracleError(rc)
    return _take(out, n.value)


def encode_symbols(sym: np.ndarray) -> bytes:
    s = np.ascontiguousarray(sym, dtype=np.uint32)
    out, n = _u8p(), C.c_size_t()
    rc = lib().llo_encode_symbols(s.ctypes.data, s.size, C.byref(out), C.byref(n))
    if rc:
        raise OracleError(rc)
    return _take(out, n.value)


def decode_tile(payload: bytes, w: int, h: int, c: int) -> np.ndarray:
    buf = np.frombuffer(payload, dtype=np.uint8) if len(payload) else np.zeros(1, np.uint8)
    out = np.empty((h, w, c), dtype=np.uint8)
    rc = lib().llo_decode_tile(buf.ctypes.data, len(payload), w, h, c, out.ctypes.data, w * c)
    if rc:
        raise OracleError(rc)
    return out


def compress(px) -> bytes:
    a = _img(px)
    h, w, c = a.shape
    out, n = _u8p(), C.c_size_t()
    rc = lib().llo_compress(a.ctypes.data, w, h, c, C.byref(out), C.byref(n))
    if rc:
        raise OracleError(rc)
    return _take(out, n.value)


def decompress(stream: bytes) -> np.ndarray:
    buf = np.frombuffer(stream, dtype=np.uint8)
    out = _u8p()
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    rc = lib().llo_decompress(buf.ctypes.data, len(stream), C.byref(out), C.byref(w), C.byref(h), C.byref(c))
    if rc:
        raise OracleError(rc)
    n = w.value * h.value * c.value
    raw = _take(out, n)
    return np.frombuffer(raw, dtype=np.uint8).reshape(h.value, w.value, c.value).copy()


def count_bins(px) -> int:
    a = _img(px)
    h, w, c = a.shape
    return int(lib().llo_count_bins(a.ctypes.data, w * c, w, h, c))


def binarize(diff: int) -> list[tuple[int, int]]:
    b = (C.c_uint8 * 40)()
    n = lib().llo_binarize(diff, b)
    return [(b[i] >> 1, b[i] & 1) for i in range(n)]


# ----- unmodified reference (guarded) --------------------------------------
def ref_compress(px) -> bytes:
    """llcomp::compressImage of the unmodified header.  Refuses inputs on which the
    reference overflows its fixed output buffer (defect D1)."""
    a = _img(px)
    h, w, c = a.shape
    if len(compress(a)) > a.size:
        raise ValueError("reference undefined here: stream longer than the raw image (D1)")
    if w > 0xFFFF or h > 0xFFFF:
        raise ValueError("reference truncates dimensions > 65535 (D3)")
    out = np.empty(a.size + 16, dtype=np.uint8)
    n = ref().ref_compress(a.ctypes.data, w, h, c, out.ctypes.data, out.size)
    return out[:n].tobytes()


def ref_decompress(stream: bytes) -> np.ndarray:
    """llcomp::decompressImage of the unmodified header (channels >= 3 only, D2)."""
    if len(stream) < 6:
        raise ValueError("short stream")
    c, w, h = stream[1], stream[2] | stream[3] << 8, stream[4] | stream[5] << 8
    if stream[0] == 0x79 and c < 3:
        raise ValueError("reference undefined here: decode with channels < 3 (D2)")
    buf = np.frombuffer(stream, dtype=np.uint8)
    out = np.empty(max(1, w * h * c), dtype=np.uint8)
    ww, hh, cc = C.c_int(), C.c_int(), C.c_int()
    rc = ref().ref_decompress(buf.ctypes.data, len(stream), out.ctypes.data, out.size,
                              C.byref(ww), C.byref(hh), C.byref(cc))
    if rc:
        raise OracleError(rc)
    return out[: w * h * c].reshape(h, w, c).copy()
