"""ctypes binding of include/llcomp_b200.h.  Fails loudly when the CUDA library is missing:
there is no CPU fallback behind this module."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB

OK, ERR_BAD_MAGIC, ERR_BAD_EXPONENT, ERR_BAD_ARG, ERR_NOMEM, ERR_OVERFLOW, ERR_CUDA, ERR_TRUNCATED = range(8)
N_STAGES = 6


class Geometry(C.Structure):
    """llcomp_geometry of include/llcomp_b200.h."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("n_images", C.c_int32)]

    def __repr__(self):
        return (f"Geometry({self.n_images}x{self.width}x{self.height}x{self.channels}, "
                f"tile {self.tile_w}x{self.tile_h})")


_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

# name -> (restype, argtypes); also the list the symbol-export test walks.
SIGNATURES = {
    "llcomp_b200_abi_version": (C.c_int, []),
    "llcomp_b200_status_string": (C.c_char_p, [C.c_int]),
    "llcomp_b200_last_error": (C.c_char_p, [_vp]),
    "llcomp_b200_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "llcomp_b200_ctx_destroy": (None, [_vp]),
    "llcomp_b200_slice_count": (C.c_uint64, [C.POINTER(Geometry)]),
    "llcomp_b200_sample_count": (C.c_uint64, [C.POINTER(Geometry)]),
    "llcomp_b200_payload_capacity": (C.c_uint64, [C.POINTER(Geometry)]),
    "llcomp_b200_stream_bound": (C.c_uint64, [C.POINTER(Geometry)]),
    "llcomp_b200_encode": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(_u8p), C.POINTER(C.c_size_t)]),
    "llcomp_b200_decode": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(_u8p),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "llcomp_b200_peek": (C.c_int, [_vp, C.c_size_t] + [C.POINTER(C.c_int)] * 5),
    "llcomp_b200_encode_batch": (C.c_int, [_vp, _vp, C.POINTER(Geometry), _vp, C.c_uint64, _vp]),
    "llcomp_b200_decode_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_uint64, C.POINTER(Geometry)]),
    "llcomp_b200_free": (None, [_vp]),
    "llcomp_b200_host_alloc": (_vp, [C.c_size_t]),
    "llcomp_b200_host_free": (None, [_vp]),
    "llcomp_b200_encode_device": (C.c_int, [_vp, _vp, C.POINTER(Geometry), _vp, C.c_uint64, _vp, _vp]),
    "llcomp_b200_decode_device": (C.c_int, [_vp, _vp, _vp, C.POINTER(Geometry), _vp, _vp]),
    "llcomp_b200_finish": (C.c_int, [_vp, _vp]),
    "llcomp_b200_frontend_device": (C.c_int, [_vp, _vp, C.POINTER(Geometry), _vp, _vp]),
    "llcomp_b200_launch_count": (C.c_uint64, [_vp]),
    "llcomp_b200_set_profiling": (None, [_vp, C.c_int]),
    "llcomp_b200_stage_times": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "llcomp_b200_stage_name": (C.c_char_p, [C.c_int]),
    "llcomp_b200_debug_table": (C.c_uint32, [C.c_int]),
    "llcomp_b200_set_queue_budget": (None, [_vp, C.c_uint64]),
    "llcomp_b200_last_bin_count": (C.c_uint64, [_vp]),
    "llcomp_b200_set_record_budget": (None, [_vp, C.c_uint64]),
    "llcomp_b200_last_encode_from_pixels": (C.c_int, [_vp]),
    "llcomp_b200_reload_switches": (None, []),
    "llcomp_b200_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "llcomp_b200_multi_destroy": (None, [_vp]),
    "llcomp_b200_multi_device_count": (C.c_int, [_vp]),
    "llcomp_b200_multi_ctx": (_vp, [_vp, C.c_int]),
    "llcomp_b200_multi_encode_batch": (C.c_int, [_vp, _vp, C.POINTER(Geometry), _vp, C.c_uint64, _vp]),
    "llcomp_b200_multi_decode_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_uint64, C.POINTER(Geometry)]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(
                f"{LIB} is missing: build it with `python -m llcomp_b200.build` (nvcc, sm_100a). "
                "llcomp_b200 has no CPU fallback.")
        # LLCOMP_B200_LIB: another build of the same library (kernel variants under test), never a different back end
        L = C.CDLL(os.environ.get("LLCOMP_B200_LIB") or LIB)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.llcomp_b200_abi_version() != 1:
            raise RuntimeError("libllcomp_b200.so ABI version mismatch")
        _lib = L
    return _lib


def status_string(code: int) -> str:
    return lib().llcomp_b200_status_string(code).decode()
