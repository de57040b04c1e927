"""Host-side mirror of the reference interface, on top of the C ABI.

``compressImage`` / ``decompressImage`` / ``RawImage`` keep the names, argument meaning and error
behaviour of /root/reference/llcomp.hpp:358, :461, :454-459 so that parity tests read like calls of
the reference.  ``Codec`` adds what the reference does not have: slices, batches, device-resident
buffers.  torch appears here only as the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import Geometry

ext = ".llcomp"             # llcomp.hpp:18
revision = 2                # llcomp.hpp:19
magic_revision = 0x77 + 2   # llcomp.hpp:20


class LlcompError(RuntimeError):
    """Carries the C-ABI status; str() is the reference's exception text where one exists
    ("Invalid magic number" llcomp.hpp:466, "Invalid exponent" llcomp.hpp:233)."""

    def __init__(self, code: int, detail: str = ""):
        msg = _capi.status_string(code)
        super().__init__(msg + (f" ({detail})" if detail else ""))
        self.code = code


@dataclass
class RawImage:
    """llcomp::RawImage (llcomp.hpp:454-459); field order matters to structured-binding callers."""
    pixels: np.ndarray   # uint8, HxWxC
    width: int
    height: int
    channels: int

    def __iter__(self):
        return iter((self.pixels, self.width, self.height, self.channels))


class Codec:
    """One per GPU: owns the library context (scratch buffers, tables) for that device."""

    def __init__(self, device: int = 0):
        self._L = _capi.lib()
        h = C.c_void_p()
        rc = self._L.llcomp_b200_ctx_create(device, C.byref(h))
        if rc:
            raise LlcompError(rc, "no usable CUDA device; llcomp_b200 has no CPU fallback")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.llcomp_b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            detail = self._L.llcomp_b200_last_error(self._h).decode() if rc == _capi.ERR_CUDA else ""
            raise LlcompError(rc, detail)

    # ---- geometry -------------------------------------------------------------------------
    @staticmethod
    def geometry(width: int, height: int, channels: int, tile_w: int = 0, tile_h: int = 0,
                 n_images: int = 1) -> Geometry:
        return Geometry(width, height, channels, tile_w, tile_h, n_images)

    def slice_count(self, g: Geometry) -> int:
        return int(self._L.llcomp_b200_slice_count(C.byref(g)))

    def payload_capacity(self, g: Geometry) -> int:
        return int(self._L.llcomp_b200_payload_capacity(C.byref(g)))

    def stream_bound(self, g: Geometry) -> int:
        return int(self._L.llcomp_b200_stream_bound(C.byref(g)))

    # ---- host buffers: the reference's two calls --------------------------------------------
    def compress(self, rgb, width: int, height: int, channels: int, tile_w: int = 0, tile_h: int = 0) -> bytes:
        """llcomp::compressImage (llcomp.hpp:358).  tile_w = tile_h = 0 (one slice) gives the reference's
        own byte stream; a tile grid gives the sliced container."""
        a = np.ascontiguousarray(np.asarray(rgb, dtype=np.uint8).reshape(-1))
        if a.size != width * height * channels:               # assert at llcomp.hpp:361
            raise ValueError("rgb.size() != width*height*channels")
        out = C.POINTER(C.c_uint8)()
        n = C.c_size_t()
        self._check(self._L.llcomp_b200_encode(self._h, a.ctypes.data, width, height, channels, tile_w, tile_h,
                                               C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            self._L.llcomp_b200_free(out)

    def decompress(self, data: bytes) -> RawImage:
        """llcomp::decompressImage (llcomp.hpp:461).  Accepts reference streams and sliced containers."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        out = C.POINTER(C.c_uint8)()
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self._L.llcomp_b200_decode(self._h, buf.ctypes.data, buf.size, C.byref(out),
                                               C.byref(w), C.byref(h), C.byref(c)))
        try:
            n = w.value * h.value * c.value
            px = np.frombuffer(C.string_at(out, n), dtype=np.uint8).reshape(h.value, w.value, c.value).copy()
        finally:
            self._L.llcomp_b200_free(out)
        return RawImage(px, w.value, h.value, c.value)

    def peek(self, data: bytes):
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        v = [C.c_int() for _ in range(5)]
        self._check(self._L.llcomp_b200_peek(buf.ctypes.data, buf.size, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)   # width, height, channels, tile_w, tile_h

    def compress_batch(self, images: np.ndarray, tile_w: int = 0, tile_h: int = 0,
                       out: Optional[np.ndarray] = None):
        """images: [N,H,W,C] uint8 (host; pinned memory avoids a staging copy).  Returns (buffer, offsets):
        stream k is buffer[offsets[k]:offsets[k+1]], each a complete stream."""
        a = np.ascontiguousarray(images, dtype=np.uint8)
        n, h, w, c = a.shape
        g = self.geometry(w, h, c, tile_w, tile_h, n)
        if out is None:
            out = np.empty(self.stream_bound(g), dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        self._check(self._L.llcomp_b200_encode_batch(self._h, a.ctypes.data, C.byref(g), out.ctypes.data, out.size,
                                                     offsets.ctypes.data))
        return out, offsets

    def decompress_batch(self, buffer: np.ndarray, offsets: Sequence[int], out: Optional[np.ndarray] = None):
        buf = np.ascontiguousarray(buffer, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = off.size - 1
        first = buf[int(off[0]):int(off[1])]                       # the whole first stream: a big slice table is part of it
        v = [C.c_int() for _ in range(5)]
        self._check(self._L.llcomp_b200_peek(first.ctypes.data, first.size, *[C.byref(x) for x in v]))
        w, h, c = v[0].value, v[1].value, v[2].value
        if out is None:
            out = np.empty((n, h, w, c), dtype=np.uint8)
        g = Geometry()
        self._check(self._L.llcomp_b200_decode_batch(self._h, buf.ctypes.data, off.ctypes.data, n, out.ctypes.data,
                                                     out.size, C.byref(g)))
        return out

    # ---- raw host pointers (bench e2e with pinned torch tensors) ----------------------------
    def encode_batch_ptr(self, pixels_ptr: int, g: Geometry, out_ptr: int, out_cap: int, offsets_ptr: int):
        self._check(self._L.llcomp_b200_encode_batch(self._h, pixels_ptr, C.byref(g), out_ptr, out_cap, offsets_ptr))

    def decode_batch_ptr(self, streams_ptr: int, offsets_ptr: int, n_images: int, out_ptr: int, out_cap: int):
        g = Geometry()
        self._check(self._L.llcomp_b200_decode_batch(self._h, streams_ptr, offsets_ptr, n_images, out_ptr, out_cap,
                                                     C.byref(g)))
        return g

    # ---- device-resident buffers (torch tensors on this codec's device) ---------------------
    @staticmethod
    def _stream_handle() -> int:
        import torch
        return int(torch.cuda.current_stream().cuda_stream)

    def frontend_device(self, pixels, g: Geometry, symbols=None):
        """K1 alone: uint8 CUDA tensor -> int32 CUDA tensor of packed (hash<<11 | diff&0x7FF) records."""
        import torch
        n = int(self._L.llcomp_b200_sample_count(C.byref(g)))
        assert pixels.is_cuda and pixels.dtype == torch.uint8 and pixels.is_contiguous() and pixels.numel() == n
        if symbols is None:
            symbols = torch.empty(n, dtype=torch.int32, device=pixels.device)
        self._check(self._L.llcomp_b200_frontend_device(self._h, pixels.data_ptr(), C.byref(g), symbols.data_ptr(),
                                                        self._stream_handle()))
        return symbols

    def encode_device(self, pixels, g: Geometry, payload=None, offsets=None):
        """Asynchronous on torch's current stream.  Returns (payload uint8[capacity], offsets int64[n_slices+1]);
        call finish() before trusting them."""
        import torch
        n = int(self._L.llcomp_b200_sample_count(C.byref(g)))
        assert pixels.is_cuda and pixels.dtype == torch.uint8 and pixels.is_contiguous() and pixels.numel() == n
        if payload is None:
            payload = torch.empty(self.payload_capacity(g), dtype=torch.uint8, device=pixels.device)
        if offsets is None:
            offsets = torch.empty(self.slice_count(g) + 1, dtype=torch.int64, device=pixels.device)
        self._check(self._L.llcomp_b200_encode_device(self._h, pixels.data_ptr(), C.byref(g), payload.data_ptr(),
                                                      payload.numel(), offsets.data_ptr(), self._stream_handle()))
        return payload, offsets

    def decode_device(self, payload, offsets, g: Geometry, pixels=None):
        import torch
        n = int(self._L.llcomp_b200_sample_count(C.byref(g)))
        if pixels is None:
            pixels = torch.empty(n, dtype=torch.uint8, device=payload.device)
        assert offsets.numel() == self.slice_count(g) + 1 and offsets.dtype == torch.int64
        self._check(self._L.llcomp_b200_decode_device(self._h, payload.data_ptr(), offsets.data_ptr(), C.byref(g),
                                                      pixels.data_ptr(), self._stream_handle()))
        return pixels

    def finish(self):
        """Synchronises torch's current stream; raises the first device-side error (overflow, invalid exponent)."""
        self._check(self._L.llcomp_b200_finish(self._h, self._stream_handle()))

    # ---- instrumentation ----------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._L.llcomp_b200_launch_count(self._h))

    def reload_switches(self):
        """Re-read the LLCOMP_* test switches from the environment (they are sampled when a context is created)."""
        self._L.llcomp_b200_reload_switches()

    def set_profiling(self, on: bool):
        self._L.llcomp_b200_set_profiling(self._h, int(on))

    def last_bin_count(self) -> int:
        return int(self._L.llcomp_b200_last_bin_count(self._h))

    def set_queue_budget(self, nbytes: int):
        """HBM the encoder's bin queue may take (default 40 % of the device); smaller forces more launch groups."""
        self._L.llcomp_b200_set_queue_budget(self._h, int(nbytes))

    def set_record_budget(self, nbytes: int):
        """HBM the front end's record array may take (default a third of the device); beyond it -- or with 0 -- the
        coder computes its records from the pixels and no array exists."""
        self._L.llcomp_b200_set_record_budget(self._h, int(nbytes))

    def last_encode_from_pixels(self) -> bool:
        return bool(self._L.llcomp_b200_last_encode_from_pixels(self._h))

    def stage_times(self) -> dict:
        ms = (C.c_float * _capi.N_STAGES)()
        self._check(self._L.llcomp_b200_stage_times(self._h, ms))
        return {self._L.llcomp_b200_stage_name(i).decode(): float(ms[i]) for i in range(_capi.N_STAGES)}


class MultiCodec:
    """Several GPUs of one box behind one call (llcomp_b200_multi_*): a batch is dealt to the devices image-wise, a
    single tiled image by bands of tile rows; one context and one host thread per device, no exchange between the
    devices.  Streams are byte-identical to a single-device Codec's."""

    def __init__(self, devices: Sequence[int]):
        self._L = _capi.lib()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self._L.llcomp_b200_multi_create(devs, len(devices), C.byref(h))
        if rc:
            raise LlcompError(rc, "no usable CUDA device; llcomp_b200 has no CPU fallback")
        self._h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None):
            self._L.llcomp_b200_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            detail = ""
            if rc == _capi.ERR_CUDA:
                detail = "; ".join(self._L.llcomp_b200_last_error(self._L.llcomp_b200_multi_ctx(self._h, k)).decode()
                                   for k in range(len(self.devices)))
            raise LlcompError(rc, detail)

    def launch_count(self) -> int:
        return sum(int(self._L.llcomp_b200_launch_count(self._L.llcomp_b200_multi_ctx(self._h, k)))
                   for k in range(len(self.devices)))

    def compress_batch(self, images: np.ndarray, tile_w: int = 0, tile_h: int = 0, out: Optional[np.ndarray] = None):
        a = np.ascontiguousarray(images, dtype=np.uint8)
        n, h, w, c = a.shape
        g = Geometry(w, h, c, tile_w, tile_h, n)
        if out is None:
            out = np.empty(int(self._L.llcomp_b200_stream_bound(C.byref(g))), dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        self._check(self._L.llcomp_b200_multi_encode_batch(self._h, a.ctypes.data, C.byref(g), out.ctypes.data, out.size,
                                                           offsets.ctypes.data))
        return out, offsets

    def decompress_batch(self, buffer: np.ndarray, offsets: Sequence[int], out: Optional[np.ndarray] = None):
        buf = np.ascontiguousarray(buffer, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = off.size - 1
        v = [C.c_int() for _ in range(5)]
        first = buf[int(off[0]):int(off[1])]
        self._check(self._L.llcomp_b200_peek(first.ctypes.data, first.size, *[C.byref(x) for x in v]))
        w, h, c = v[0].value, v[1].value, v[2].value
        if out is None:
            out = np.empty((n, h, w, c), dtype=np.uint8)
        g = Geometry()
        self._check(self._L.llcomp_b200_multi_decode_batch(self._h, buf.ctypes.data, off.ctypes.data, n, out.ctypes.data,
                                                           out.size, C.byref(g)))
        return out

    def compress(self, rgb, width: int, height: int, channels: int, tile_w: int = 0, tile_h: int = 0) -> bytes:
        """One image; with a tile grid of two or more tile rows the bands go to different devices."""
        a = np.ascontiguousarray(np.asarray(rgb, dtype=np.uint8).reshape(1, height, width, channels))
        out, off = self.compress_batch(a, tile_w, tile_h)
        return out[:int(off[1])].tobytes()

    def decompress(self, data: bytes) -> RawImage:
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        px = self.decompress_batch(buf, [0, buf.size])
        _, h, w, c = px.shape
        return RawImage(px[0], w, h, c)

    def encode_batch_ptr(self, pixels_ptr: int, g: Geometry, out_ptr: int, out_cap: int, offsets_ptr: int):
        self._check(self._L.llcomp_b200_multi_encode_batch(self._h, pixels_ptr, C.byref(g), out_ptr, out_cap, offsets_ptr))

    def decode_batch_ptr(self, streams_ptr: int, offsets_ptr: int, n_images: int, out_ptr: int, out_cap: int):
        g = Geometry()
        self._check(self._L.llcomp_b200_multi_decode_batch(self._h, streams_ptr, offsets_ptr, n_images, out_ptr, out_cap,
                                                           C.byref(g)))
        return g


_default: dict[int, Codec] = {}


def default_codec(device: int = 0) -> Codec:
    if device not in _default:
        _default[device] = Codec(device)
    return _default[device]


def compressImage(rgb, width: int, height: int, channels: int) -> bytes:
    """Drop-in for llcomp::compressImage (llcomp.hpp:358): same arguments, byte-identical stream."""
    return default_codec().compress(rgb, width, height, channels)


def decompressImage(data: bytes) -> RawImage:
    """Drop-in for llcomp::decompressImage (llcomp.hpp:461)."""
    return default_codec().decompress(data)
