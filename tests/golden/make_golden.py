#!/usr/bin/env python
"""Regenerates tests/golden/streams.json and tests/golden/small_fixtures.json.

Run in the dev container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py

Every entry records which implementation produced it:
  "ref"          the UNMODIFIED reference header (oracle/_ref, llcomp.hpp:358/:461)
  "restatement"  oracle/llcomp_oracle.c, only where the reference is undefined
                 (stream longer than raw: D1; channels<3 decode: D2)
Entries produced by "ref" are what pin parity; the script also asserts that the
restatement agrees byte-for-byte wherever the reference is defined.

Note: SURVEY.md appendix B lists FNV-1a64 values that could not be reproduced
with the stated basis/prime (the stream LENGTHS all agree); the hashes stored
here are recomputed from the unmodified reference and checked with two
independent FNV implementations (C in the oracle, pure Python below).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fnv_py(b: bytes) -> int:
    h = 0xCBF29CE484222325
    for x in b:
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def stream_of(img):
    """(stream, source) using the reference wherever it is defined."""
    s = oracle.compress(img)
    if len(s) <= img.size and img.shape[0] <= 0xFFFF and img.shape[1] <= 0xFFFF:
        r = oracle.ref_compress(img)
        assert r == s, "restatement disagrees with the reference header"
        if img.shape[2] >= 3:
            assert (oracle.ref_decompress(r) == img).all()
        return r, "ref"
    return s, "restatement"


def main():
    out = {"generator": "G(W,H,C,n,seed) of SURVEY.md appendix C; n<0 = rng()&0xFF", "whole": [], "tiled": []}
    whole = [(512, 512, 3, 4), (512, 512, 3, 0), (1024, 1024, 3, 0), (1024, 1024, 3, 2),
             (1024, 1024, 3, 4), (1024, 1024, 3, 8), (1024, 1024, 3, 16), (1024, 1024, 3, 32),
             (1024, 1024, 4, 8), (256, 256, 1, 4), (256, 256, 2, 4), (256, 256, 1, -1),
             (256, 256, 3, -1), (300, 200, 3, 6), (64, 64, 5, 3)]
    if os.environ.get("GOLDEN_BIG"):
        whole.append((4096, 4096, 3, 4))
    for (w, h, c, n) in whole:
        img = oracle.generate(w, h, c, n, 1234)
        s, src = stream_of(img)
        assert oracle.fnv1a64(s) == fnv_py(s)
        assert (oracle.decompress(s) == img).all()
        out["whole"].append({"w": w, "h": h, "c": c, "n": n, "seed": 1234, "bytes": len(s),
                             "fnv1a64": f"{fnv_py(s):016x}", "bins": oracle.count_bins(img), "source": src})
        print(out["whole"][-1])
    # per-tile payloads == reference(tile)[6:]
    for (w, h, c, n, tw, th) in [(1024, 1024, 3, 4, 512, 512), (1024, 1024, 3, 4, 256, 256),
                                 (1024, 1024, 3, 4, 1024, 64), (600, 500, 3, 4, 256, 128),
                                 (512, 512, 1, -1, 128, 128)]:
        img = oracle.generate(w, h, c, n, 1234)
        tiles = []
        for y0 in range(0, h, th):
            for x0 in range(0, w, tw):
                t = np.ascontiguousarray(img[y0:y0 + th, x0:x0 + tw])
                s, src = stream_of(t)
                assert oracle.encode_tile(img, x0, y0, t.shape[1], t.shape[0]) == s[6:]
                tiles.append({"x0": x0, "y0": y0, "bytes": len(s) - 6, "fnv1a64": f"{fnv_py(s[6:]):016x}",
                              "source": src})
        out["tiled"].append({"w": w, "h": h, "c": c, "n": n, "seed": 1234, "tile_w": tw, "tile_h": th,
                             "payload_bytes": sum(t["bytes"] for t in tiles), "tiles": tiles})
        print(w, h, c, n, tw, th, out["tiled"][-1]["payload_bytes"])
    with open(os.path.join(HERE, "streams.json"), "w") as f:
        json.dump(out, f, indent=1)

    # small explicit fixtures (pixels + stream stored verbatim)
    rng = np.random.default_rng(20261018)
    small = []
    shapes = [(1, 1, 3), (1, 1, 1), (2, 2, 3), (3, 1, 3), (1, 3, 3), (5, 7, 3), (7, 5, 4), (17, 33, 3),
              (33, 17, 1), (16, 16, 2), (31, 9, 3), (9, 31, 3), (40, 40, 3), (64, 3, 3), (3, 64, 3), (24, 24, 6)]
    for k, (w, h, c) in enumerate(shapes):
        amp = [0, 1, 3, 8, 40, 128][k % 6]
        base = (np.add.outer(np.arange(h) * 3, np.arange(w) * 2)[:, :, None] + np.arange(c) * 17) % 256
        noise = rng.integers(-amp, amp + 1, size=(h, w, c)) if amp else 0
        img = np.clip(base + noise, 0, 255).astype(np.uint8)
        s, src = stream_of(img)
        small.append({"w": w, "h": h, "c": c, "pixels": img.flatten().tolist(), "stream": s.hex(), "source": src})
    with open(os.path.join(HERE, "small_fixtures.json"), "w") as f:
        json.dump({"fixtures": small}, f)
    print("small fixtures:", len(small), [x["source"] for x in small])


if __name__ == "__main__":
    main()
