"""llcomp_b200 -- B200-native (sm_100a) implementation of llcomp's encode/decode hot path.

Only what the path needs lives here:
  csrc/     hand-written CUDA kernels and the extern "C" layer (include/llcomp_b200.h)
  host/     C++ mirror of the reference interface (llcomp.hpp) and the llcompc / llcompd tools
  codec.py  Python mirror of the same interface, used by tests and bench.py
"""
from .codec import (Codec, Geometry, LlcompError, MultiCodec, RawImage, compressImage, decompressImage, default_codec, ext,
                    magic_revision, revision)

__all__ = ["Codec", "Geometry", "LlcompError", "MultiCodec", "RawImage", "compressImage", "decompressImage", "default_codec",
           "ext", "magic_revision", "revision"]
