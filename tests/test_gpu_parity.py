"""Parity of the CUDA path with the CPU oracle, through the C ABI (llcomp_b200.Codec -> ctypes ->
libllcomp_b200.so).  Bit-exact everywhere: this is integer/byte work.

Mirrors the test list of SURVEY.md section 4: known answers, differential vs the reference semantics,
round trips, tail padding, carry propagation, plus size-independent properties at BASELINE sizes.
"""
import json
import os
import struct

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    import llcomp_b200
    return llcomp_b200.default_codec(0)


import contextlib


@contextlib.contextmanager
def switched(codec, **env):
    """Run a block with LLCOMP_* test switches set (the library samples them at context creation, so tell it)."""
    for k, v in env.items():
        os.environ[k] = str(v)
    codec.reload_switches()
    try:
        yield
    finally:
        for k in env:
            del os.environ[k]
        codec.reload_switches()


def split_container(stream: bytes):
    """(w, h, c, tile_w, tile_h, [payload per slice]) of either stream layout."""
    if stream[0] == 0x79:
        c, w, h = stream[1], stream[2] | stream[3] << 8, stream[4] | stream[5] << 8
        return w, h, c, w, h, [stream[6:]]
    assert stream[0] == 0xB2 and stream[1] == 1
    c = stream[2]
    w, h, tw, th, n = struct.unpack_from("<5I", stream, 4)
    lens = struct.unpack_from(f"<{n}I", stream, 24)
    pos, out = 24 + 4 * n, []
    for L in lens:
        out.append(stream[pos:pos + L])
        pos += L
    assert pos == len(stream)
    return w, h, c, tw, th, out


def tiles_of(w, h, tw, th):
    return [(x0, y0, min(tw, w - x0), min(th, h - y0)) for y0 in range(0, h, th) for x0 in range(0, w, tw)]


def rand_image(rng, w, h, c, amp):
    base = (np.add.outer(np.arange(h) * 2, np.arange(w) * 3)[:, :, None] + np.arange(c) * 23) % 256
    noise = rng.integers(-amp, amp + 1, size=(h, w, c)) if amp else 0
    return np.clip(base + noise, 0, 255).astype(np.uint8)


# ---- known answers ----------------------------------------------------------------------------
def test_kat_streams_byte_identical(codec, kats):
    for name, px, stream in kats:
        h, w, c = px.shape
        assert codec.compress(px, w, h, c) == stream, name


def test_kat_decode(codec, kats):
    for name, px, stream in kats:
        img = codec.decompress(stream)
        assert (img.width, img.height, img.channels) == (px.shape[1], px.shape[0], px.shape[2]), name
        assert (img.pixels == px).all(), name


def test_small_fixtures(codec):
    with open(os.path.join(GOLDEN, "small_fixtures.json")) as f:
        fx = json.load(f)["fixtures"]
    for x in fx:
        px = np.array(x["pixels"], dtype=np.uint8).reshape(x["h"], x["w"], x["c"])
        s = bytes.fromhex(x["stream"])
        assert codec.compress(px, x["w"], x["h"], x["c"]) == s
        assert (codec.decompress(s).pixels == px).all()


def test_module_level_functions_mirror_reference_api():
    import llcomp_b200
    px = oracle.generate(40, 30, 3, 4, 2)
    s = llcomp_b200.compressImage(px.reshape(-1), 40, 30, 3)
    assert s == oracle.compress(px)
    pixels, width, height, channels = llcomp_b200.decompressImage(s)   # structured binding, llcompd.cpp:26
    assert (width, height, channels) == (40, 30, 3) and (pixels == px).all()
    assert llcomp_b200.ext == ".llcomp" and llcomp_b200.magic_revision == 0x79


# ---- front end (K1) -----------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,c,tw,th,amp", [
    (64, 48, 3, 0, 0, 8), (65, 33, 3, 16, 16, 4), (37, 29, 1, 0, 0, 16), (37, 29, 2, 8, 32, 2),
    (50, 40, 4, 32, 8, 128), (19, 23, 5, 7, 5, 6), (1, 1, 3, 0, 0, 0), (1, 9, 3, 0, 0, 9), (9, 1, 3, 4, 1, 9),
    (2, 2, 3, 1, 1, 50), (300, 70, 3, 128, 64, 5),
    # the streaming kernel (W*C % 16 == 0, tile_w % 4 == 0): whole rows, partial last region, tile rows that end
    # inside a region, one-group tiles (left and right edge in the same thread), rows that are all h < 2
    (1024, 70, 3, 0, 0, 4), (1040, 67, 3, 512, 32, 6), (528, 40, 4, 132, 9, 5), (16, 5, 3, 4, 2, 30),
    (32, 40, 3, 8, 3, 10), (2048, 35, 3, 1024, 33, 64), (64, 100, 4, 64, 100, 128)])
def test_frontend_records_equal_oracle(codec, w, h, c, tw, th, amp):
    import torch
    rng = np.random.default_rng(w * 1000 + h)
    n_img = 2
    imgs = np.stack([rand_image(rng, w, h, c, amp) for _ in range(n_img)])
    g = codec.geometry(w, h, c, tw, th, n_img)
    sym = codec.frontend_device(torch.from_numpy(imgs).cuda(), g)
    codec.finish()
    got = sym.cpu().numpy().view(np.uint32)
    want = []
    for k in range(n_img):
        for (x0, y0, sw, sh) in tiles_of(w, h, tw or w, th or h):
            want.append(oracle.frontend(imgs[k], x0, y0, sw, sh))
    want = np.concatenate(want)
    assert got.shape == want.shape
    assert (got == want).all()


# ---- whole streams ------------------------------------------------------------------------------
def test_cfg1_512_rgb_byte_identical(codec, golden_streams):
    g = next(x for x in golden_streams["whole"] if (x["w"], x["h"], x["c"], x["n"]) == (512, 512, 3, 4))
    img = oracle.generate(512, 512, 3, 4, 1234)
    s = codec.compress(img, 512, 512, 3)
    assert len(s) == g["bytes"] == 402823
    assert f"{oracle.fnv1a64(s):016x}" == g["fnv1a64"]      # hash of the unmodified reference's stream
    assert s == oracle.compress(img)
    assert (codec.decompress(s).pixels == img).all()


@pytest.mark.parametrize("w,h,c,n", [(1024, 1024, 3, 0), (1024, 1024, 3, 4), (1024, 1024, 3, 32), (1024, 1024, 4, 8),
                                     (256, 256, 1, 4), (256, 256, 2, 4), (300, 200, 3, 6), (64, 64, 5, 3)])
def test_golden_streams(codec, golden_streams, w, h, c, n):
    g = next(x for x in golden_streams["whole"] if (x["w"], x["h"], x["c"], x["n"]) == (w, h, c, n))
    img = oracle.generate(w, h, c, n, 1234)
    s = codec.compress(img, w, h, c)
    assert len(s) == g["bytes"] and f"{oracle.fnv1a64(s):016x}" == g["fnv1a64"]
    assert (codec.decompress(s).pixels == img).all()


@pytest.mark.parametrize("c", [1, 3])
def test_uniform_noise_stream_longer_than_raw(codec, c):
    # the reference overflows here (llcomp.hpp:362); the oracle restatement defines the answer
    img = oracle.generate(256, 256, c, -1, 1234)
    s = codec.compress(img, 256, 256, c)
    assert len(s) > img.size
    assert s == oracle.compress(img)
    assert (codec.decompress(s).pixels == img).all()


def test_constant_image_carry_runs(codec):
    for v in (0, 77, 128, 255):
        img = np.full((96, 80, 3), v, np.uint8)
        s = codec.compress(img, 80, 96, 3)
        assert s == oracle.compress(img)
        assert (codec.decompress(s).pixels == img).all()


def test_differential_random(codec):
    rng = np.random.default_rng(11)
    for k in range(60):
        w, h = int(rng.integers(1, 97)), int(rng.integers(1, 97))
        c = int(rng.choice([1, 2, 3, 3, 3, 4, 6]))
        amp = int(rng.choice([0, 1, 2, 4, 16, 64, 128]))
        img = rand_image(rng, w, h, c, amp)
        s = codec.compress(img, w, h, c)
        assert s == oracle.compress(img), (w, h, c, amp)
        assert (codec.decompress(s).pixels == img).all(), (w, h, c, amp)


# ---- slices ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,c,n,tw,th", [(1024, 1024, 3, 4, 512, 512), (600, 500, 3, 4, 256, 128),
                                           (512, 512, 1, -1, 128, 128), (1024, 1024, 3, 4, 1024, 64)])
def test_slice_payloads_equal_reference_on_tile(codec, golden_streams, w, h, c, n, tw, th):
    g = next(t for t in golden_streams["tiled"] if (t["w"], t["h"], t["c"], t["n"], t["tile_w"], t["tile_h"]) ==
             (w, h, c, n, tw, th))
    img = oracle.generate(w, h, c, n, 1234)
    s = codec.compress(img, w, h, c, tw, th)
    W, H, Cc, TW, TH, payloads = split_container(s)
    assert (W, H, Cc, TW, TH) == (w, h, c, tw, th) and len(payloads) == len(g["tiles"])
    assert sum(map(len, payloads)) == g["payload_bytes"]
    for p, t in zip(payloads, g["tiles"]):
        assert len(p) == t["bytes"] and f"{oracle.fnv1a64(p):016x}" == t["fnv1a64"]
    assert (codec.decompress(s).pixels == img).all()
    assert codec.peek(s) == (w, h, c, tw, th)


def test_ragged_tiles_random(codec):
    rng = np.random.default_rng(5)
    for k in range(12):
        w, h = int(rng.integers(8, 200)), int(rng.integers(8, 200))
        c = int(rng.choice([1, 3, 4]))
        tw, th = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
        img = rand_image(rng, w, h, c, int(rng.choice([0, 3, 40])))
        s = codec.compress(img, w, h, c, tw, th)
        _, _, _, _, _, payloads = split_container(s)
        for p, (x0, y0, sw, sh) in zip(payloads, tiles_of(w, h, tw, th)):
            assert p == oracle.encode_tile(img, x0, y0, sw, sh), (w, h, c, tw, th, x0, y0)
        assert (codec.decompress(s).pixels == img).all()


def test_single_slice_grid_is_reference_stream(codec):
    img = oracle.generate(128, 96, 3, 4, 3)
    assert codec.compress(img, 128, 96, 3, 128, 96) == oracle.compress(img)
    assert codec.compress(img, 128, 96, 3, 4096, 4096) == oracle.compress(img)


# ---- decoder edge behaviour ---------------------------------------------------------------------
def test_decode_of_reference_encoder_output(codec):
    for (w, h, c, n) in [(200, 150, 3, 8), (64, 64, 4, 2), (90, 70, 1, 16)]:
        img = oracle.generate(w, h, c, n, 77)
        assert (codec.decompress(oracle.compress(img)).pixels == img).all()


def test_truncated_and_padded_streams_match_oracle(codec):
    img = oracle.generate(48, 40, 3, 8, 21)
    s = oracle.compress(img)
    assert (codec.decompress(s + bytes(16)).pixels == img).all()          # explicit zeros == zero fill
    for cut in (1, 2, 5, 40, len(s) - 7):
        t = s[:len(s) - cut]
        assert (codec.decompress(t).pixels == oracle.decompress(t)).all(), cut   # same garbage, llcomp.hpp:476-477
    t = s + b"\xff" * 4
    assert (codec.decompress(t).pixels == oracle.decompress(t)).all()


def test_error_behaviour(codec):
    from llcomp_b200 import LlcompError
    with pytest.raises(LlcompError, match="Invalid magic number") as e:      # llcomp.hpp:466
        codec.decompress(bytes([0x77, 3, 4, 0, 4, 0, 1, 2, 3]))
    assert e.value.code == 1
    # a nonzero flag followed by all ones: the exponent runs away -> "Invalid exponent" (llcomp.hpp:233)
    bad = bytes([0x79, 1, 4, 0, 4, 0, 2, 250]) + b"\xff" * 40
    with pytest.raises(oracle.OracleError) as eo:
        oracle.decompress(bad)
    assert eo.value.code == 2
    with pytest.raises(LlcompError, match="Invalid exponent") as e:
        codec.decompress(bad)
    assert e.value.code == 2
    with pytest.raises(ValueError):
        codec.compress(np.zeros(10, np.uint8), 2, 2, 3)                       # assert at llcomp.hpp:361


# ---- batches and device-resident buffers ----------------------------------------------------------
def test_batch_streams_are_standalone_reference_streams(codec):
    imgs = np.stack([oracle.generate(160, 120, 3, 4, 1234 + k) for k in range(9)])
    buf, off = codec.compress_batch(imgs)
    for k in range(9):
        assert buf[int(off[k]):int(off[k + 1])].tobytes() == oracle.compress(imgs[k])
    assert (codec.decompress_batch(buf, off) == imgs).all()
    buf2, off2 = codec.compress_batch(imgs, 64, 64)
    assert (codec.decompress_batch(buf2, off2) == imgs).all()
    assert int(off2[-1]) > int(off[-1])


def test_split_coder_path_and_launch_groups(codec):
    """The two-kernel coder (model pass -> bin queue in HBM -> range pass) stays available behind
    LLCOMP_CODER_SPLIT; a small queue budget forces it to code the slices in several launch groups."""
    import torch
    imgs = np.stack([oracle.generate(128, 96, 3, 8, 70 + k) for k in range(7)])
    g = codec.geometry(128, 96, 3, 64, 48, 7)
    d_px = torch.from_numpy(imgs).cuda()
    ref_payload, ref_off = codec.encode_device(d_px, g)          # default: fused coder
    codec.finish()
    codec.set_queue_budget(200_000)                 # ~3 slices per group instead of all 28
    try:
        with switched(codec, LLCOMP_CODER_SPLIT=1):
            payload, off = codec.encode_device(d_px, g)
            codec.finish()
            assert codec.last_bin_count() == sum(oracle.count_bins(imgs[k][y0:y0 + 48, x0:x0 + 64]) for k in range(7)
                                                 for (x0, y0, _, _) in tiles_of(128, 96, 64, 48))
    finally:
        codec.set_queue_budget(64 << 30)
    assert torch.equal(off, ref_off)
    n = int(off[-1])
    assert torch.equal(payload[:n], ref_payload[:n])
    assert payload[int(off[4]):int(off[5])].cpu().numpy().tobytes() == oracle.encode_tile(imgs[1], 0, 0, 64, 48)


def test_many_slices_decode_with_state_behind_l1(codec):
    """More slices than shared-memory slots (3 x 148): the decoder keeps the state rows in global memory."""
    import torch
    imgs = np.stack([oracle.generate(256, 192, 3, 6, 300 + k) for k in range(16)])
    g = codec.geometry(256, 192, 3, 32, 32, 16)                      # 48 tiles x 16 images = 768 slices
    assert codec.slice_count(g) == 768
    d_px = torch.from_numpy(imgs).cuda()
    payload, offsets = codec.encode_device(d_px, g)
    codec.finish()
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    assert torch.equal(out.view(16, 192, 256, 3), d_px)
    # same answer with the state in shared memory, and from the two-kernel coder
    with switched(codec, LLCOMP_DECODER_SMEM_STATE=1, LLCOMP_CODER_SPLIT=1):
        out2 = codec.decode_device(payload, offsets, g)
        codec.finish()
        payload2, offsets2 = codec.encode_device(d_px, g)
        codec.finish()
    assert torch.equal(out2, out)
    assert torch.equal(offsets2, offsets) and torch.equal(payload2[: int(offsets[-1])], payload[: int(offsets[-1])])
    off = offsets.cpu().numpy()
    k = 48 * 5 + 17                                                   # one slice against the oracle
    x0, y0, sw, sh = tiles_of(256, 192, 32, 32)[17]
    assert payload[int(off[k]):int(off[k + 1])].cpu().numpy().tobytes() == oracle.encode_tile(imgs[5], x0, y0, sw, sh)


def test_fused_coder_variants_agree(codec):
    """The fused coder serves 1..7 slices per CTA with one CTA per SM (the default picks by slice count; 11..17 force
    it, 23..27 give the recurrence warp a sub-partition of its own), or 1, 2, 4 slices per chain warp in the round-1
    arrangement (LLCOMP_FUSED_NS), always with the state rows
    behind L1, or one slice per CTA with the rows in shared memory (LLCOMP_MODEL_SMEM_STATE): same bytes from all of
    them, with ragged slices and a slice count (378) that leaves the last CTA of every form partly empty."""
    import torch
    imgs = np.stack([oracle.generate(200, 180, 3, 9, 4100 + k) for k in range(9)])
    g = codec.geometry(200, 180, 3, 32, 32, 9)                       # 7 x 6 tiles (last column 8 wide, last row 20 high)
    assert codec.slice_count(g) == 378
    d_px = torch.from_numpy(imgs).cuda()
    payload, offsets = codec.encode_device(d_px, g)
    codec.finish()
    n = int(offsets[-1])
    for var, val in (("LLCOMP_FUSED_NS", "1"), ("LLCOMP_FUSED_NS", "2"), ("LLCOMP_FUSED_NS", "4"),
                     ("LLCOMP_FUSED_NS", "11"), ("LLCOMP_FUSED_NS", "12"), ("LLCOMP_FUSED_NS", "13"),
                     ("LLCOMP_FUSED_NS", "14"), ("LLCOMP_FUSED_NS", "15"), ("LLCOMP_FUSED_NS", "16"),
                     ("LLCOMP_FUSED_NS", "17"), ("LLCOMP_FUSED_NS", "23"), ("LLCOMP_FUSED_NS", "24"),
                     ("LLCOMP_FUSED_NS", "25"), ("LLCOMP_FUSED_NS", "27"), ("LLCOMP_MODEL_SMEM_STATE", "1")):
        for pixels in (0, 1):                                        # records from K1's record array / from the pixels
            with switched(codec, **{var: val, "LLCOMP_CODER_PIXELS": pixels}):
                p2, o2 = codec.encode_device(d_px, g)
                codec.finish()
            assert torch.equal(o2, offsets) and torch.equal(p2[:n], payload[:n]), (var, val, pixels)
    off = offsets.cpu().numpy()
    tiles = tiles_of(200, 180, 32, 32)
    for img, t in ((0, 0), (3, 6), (8, 41), (5, 20)):                 # corners and an interior slice against the oracle
        k = img * 42 + t
        x0, y0, sw, sh = tiles[t]
        assert payload[int(off[k]):int(off[k + 1])].cpu().numpy().tobytes() == oracle.encode_tile(imgs[img], x0, y0, sw, sh)
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    assert torch.equal(out.view(imgs.shape), d_px)


@pytest.mark.parametrize("c", [1, 2, 4])
def test_many_slices_other_channel_counts(codec, c):
    """480 slices of 1-, 2- and 4-channel images, half of them flat: the state rows of both the fused coder and the
    decoder live behind L1, and flat content makes consecutive samples share a context row, which is the case the
    decoder's row forwarding (rows requested one pixel ahead, rewritten in between) exists for."""
    import torch
    imgs = np.stack([oracle.generate(128, 96, c, 0 if k % 2 else 6, 7000 + 10 * c + k) for k in range(10)])
    g = codec.geometry(128, 96, c, 16, 16, 10)
    assert codec.slice_count(g) == 480
    d_px = torch.from_numpy(imgs).cuda()
    payload, offsets = codec.encode_device(d_px, g)
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    assert torch.equal(out.view(imgs.shape), d_px)
    off = offsets.cpu().numpy()
    tiles = tiles_of(128, 96, 16, 16)
    for img, t in ((0, 0), (1, 47), (6, 13), (9, 30)):
        k = img * 48 + t
        x0, y0, sw, sh = tiles[t]
        assert payload[int(off[k]):int(off[k + 1])].cpu().numpy().tobytes() == oracle.encode_tile(imgs[img], x0, y0, sw, sh)


def test_alternate_kernels_agree(codec):
    """Every stage has a plain variant behind a switch (one-thread-per-pixel front end, plain decoder chain,
    two-kernel coder); all of them must produce the same bytes / pixels as the default kernels."""
    import torch
    imgs = np.stack([oracle.generate(208, 144, 3, 5, 900 + k) for k in range(3)])    # 208*3 % 16 == 0: streaming K1
    g = codec.geometry(208, 144, 3, 96, 64, 3)
    d_px = torch.from_numpy(imgs).cuda()
    payload, offsets = codec.encode_device(d_px, g)
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    n = int(offsets[-1])
    assert torch.equal(out.view(imgs.shape), d_px)
    # (LLCOMP_CODER_PIXELS: the fused coder computes its records from the pixels itself, no K1 and no record array)
    for env in ({"LLCOMP_CODER_PIXELS": 1}, {"LLCOMP_FRONTEND_SIMPLE": 1}, {"LLCOMP_FRONTEND_TILED": 1},
                {"LLCOMP_FRONTEND_VARIANT": 1}, {"LLCOMP_FRONTEND_VARIANT": 5},
                {"LLCOMP_FRONTEND_VARIANT": 4},
                {"LLCOMP_DECODER_SIMPLE": 1}, {"LLCOMP_DECODER_V1": 1}, {"LLCOMP_DECODER_V1": 1, "LLCOMP_DECODER_SMEM_STATE": 1},
                {"LLCOMP_DECODER_SMEM_STATE": 1}, {"LLCOMP_CODER_SPLIT": 1}):
        switch = "+".join(env)
        with switched(codec, **env):
            p2, o2 = codec.encode_device(d_px, g)
            out2 = codec.decode_device(p2, o2, g)
            codec.finish()
        assert torch.equal(o2, offsets) and torch.equal(p2[:n], payload[:n]), switch
        assert torch.equal(out2, out), switch


def test_decoder_forms_agree_on_many_slices(codec):
    """The chain decoder (default; decoder_chain.cuh) with its state rows behind L1 (many slices) and in shared memory,
    its measurement variants, the round-1 fast decoder and the plain chain: the same pixels from all of them, and the
    pixels are the originals.  1, 2, 3 and 4 channels; noise from flat to uniform random bytes."""
    import torch
    for c, noise, w, h, tw, th, n in ((3, 4, 320, 96, 64, 32, 40), (1, -1, 256, 64, 32, 32, 64), (4, 16, 128, 64, 32, 16, 40),
                                      (2, 0, 96, 64, 48, 16, 80), (3, 64, 1024, 8, 1024, 8, 4)):
        imgs = np.stack([oracle.generate(w, h, c, noise, 4000 + k) for k in range(n)])
        g = codec.geometry(w, h, c, tw, th, n)
        d_px = torch.from_numpy(imgs).cuda()
        payload, offsets = codec.encode_device(d_px, g)
        out = codec.decode_device(payload, offsets, g)
        codec.finish()
        assert torch.equal(out.view(imgs.shape), d_px), (c, noise)
        for env in ({"LLCOMP_DECODER_SMEM_STATE": 1}, {"LLCOMP_DECODER_V1": 1}, {"LLCOMP_DECODER_SIMPLE": 1},
                    {"LLCOMP_DECODER_VARIANT": 1}, {"LLCOMP_DECODER_VARIANT": 2}, {"LLCOMP_DECODER_VARIANT": 4},
                    {"LLCOMP_DECODER_MAX_CARVEOUT": 1}):
            with switched(codec, **env):
                out2 = codec.decode_device(payload, offsets, g)
                codec.finish()
            assert torch.equal(out2, out), (c, noise, env)


def test_very_wide_slices_decode(codec):
    """Slices 16384 samples-times-3 wide: the chain decoder's two row buffers take 197 KB of shared memory (state rows
    behind L1); four channels at that width do not fit and take the plain chain.  Streams equal the oracle's, pixels
    come back exactly."""
    for c in (3, 4, 1):
        img = oracle.generate(16384, 5, c, 6, 600 + c)
        s = codec.compress(img, 16384, 5, c)
        assert s == oracle.compress(img), c
        assert (codec.decompress(s).pixels == img).all(), c


def test_batch_of_megapixel_images_matches_reference(codec):
    """BASELINE configs[3] in small: 1024x1024 RGB images, one slice each, streams byte-identical to the oracle."""
    imgs = np.stack([oracle.generate(1024, 1024, 3, 4, 1234 + k) for k in range(3)])
    buf, off = codec.compress_batch(imgs)
    for k in range(3):
        assert buf[int(off[k]):int(off[k + 1])].tobytes() == oracle.compress(imgs[k])
    assert (codec.decompress_batch(buf, off) == imgs).all()


def test_device_resident_round_trip(codec):
    import torch
    imgs = np.stack([oracle.generate(256, 192, 3, 6, 50 + k) for k in range(5)])
    g = codec.geometry(256, 192, 3, 128, 64, 5)
    d_px = torch.from_numpy(imgs).cuda()
    payload, offsets = codec.encode_device(d_px, g)
    codec.finish()
    off = offsets.cpu().numpy()
    pay = payload.cpu().numpy()
    k = 0
    for i in range(5):
        for (x0, y0, sw, sh) in tiles_of(256, 192, 128, 64):
            assert pay[off[k]:off[k + 1]].tobytes() == oracle.encode_tile(imgs[i], x0, y0, sw, sh)
            k += 1
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    assert (out.cpu().numpy().reshape(imgs.shape) == imgs).all()


# ---- BASELINE sizes: size-independent properties + sampled oracle checks ---------------------------
def test_cfg2_4096_rgb_64_slices(codec):
    import torch
    img = oracle.generate(4096, 4096, 3, 4, 1234)
    g = codec.geometry(4096, 4096, 3, 512, 512, 1)
    d_px = torch.from_numpy(img).cuda()
    payload, offsets = codec.encode_device(d_px, g)
    codec.finish()
    off = offsets.cpu().numpy()
    assert off.size == 65 and (np.diff(off) > 0).all()
    pay = payload[: int(off[-1])].cpu().numpy()
    tiles = tiles_of(4096, 4096, 512, 512)
    for k in (0, 7, 27, 36, 56, 63):                                   # corners + interior, vs the oracle
        x0, y0, sw, sh = tiles[k]
        assert pay[off[k]:off[k + 1]].tobytes() == oracle.encode_tile(img, x0, y0, sw, sh), k
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    assert torch.equal(out.view(4096, 4096, 3), d_px)                   # exact round trip of every pixel
    bpp = 8.0 * int(off[-1]) / (4096 * 4096)
    assert 11.9 < bpp < 12.6                                           # single slice: 11.994


def test_cfg3_gray_noise_round_trip(codec):
    import torch
    img = oracle.generate(2048, 2048, 1, -1, 1234)                       # content of config 3 at 1/16 area
    g = codec.geometry(2048, 2048, 1, 512, 512, 1)
    d_px = torch.from_numpy(img).cuda()
    payload, offsets = codec.encode_device(d_px, g)
    codec.finish()
    off = offsets.cpu().numpy()
    assert int(off[-1]) > img.size                                      # ~1.23x raw
    k = 5
    x0, y0, sw, sh = tiles_of(2048, 2048, 512, 512)[k]
    assert payload[int(off[k]):int(off[k + 1])].cpu().numpy().tobytes() == oracle.encode_tile(img, x0, y0, sw, sh)
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    assert torch.equal(out.view(2048, 2048, 1), d_px)


def test_wide_tile_uses_global_row_scratch(codec):
    # a 20000-wide strip does not fit the shared-memory row buffers of the decoder
    img = oracle.generate(20000, 6, 3, 4, 8)
    s = codec.compress(img, 20000, 6, 3)
    assert s == oracle.compress(img)
    assert (codec.decompress(s).pixels == img).all()


def test_coder_works_from_pixels_when_the_record_array_does_not_fit(codec):
    """no 4-byte-per-sample intermediate.  By default K1 writes a record array for the coder
    (faster); when the array exceeds its budget the coder's model warps compute the records from the pixels inside
    the CTA.  Same bytes either way; RGB and RGBA; whole-image slices and ragged tiles; host-buffer path too."""
    import torch
    for c, tw, th in ((3, 0, 0), (4, 48, 40), (3, 100, 7)):
        imgs = np.stack([oracle.generate(136, 90, c, 6, 600 + k) for k in range(5)])
        g = codec.geometry(136, 90, c, tw, th, 5)
        d_px = torch.from_numpy(imgs).cuda()
        want_p, want_o = codec.encode_device(d_px, g)
        codec.finish()
        assert not codec.last_encode_from_pixels()
        n = int(want_o[-1])
        codec.set_record_budget(1000)                         # far less than 4 bytes x 183,600 samples
        try:
            got_p, got_o = codec.encode_device(d_px, g)
            codec.finish()
            assert codec.last_encode_from_pixels()
            buf, off = codec.compress_batch(imgs, tw, th)
            assert codec.last_encode_from_pixels()
        finally:
            codec.set_record_budget(60 << 30)
        assert torch.equal(got_o, want_o) and torch.equal(got_p[:n], want_p[:n]), (c, tw, th)
        buf2, off2 = codec.compress_batch(imgs, tw, th)
        assert (off == off2).all() and (buf[: int(off[-1])] == buf2[: int(off[-1])]).all()
        if tw == 0:
            assert buf[: int(off[1])].tobytes() == oracle.compress(imgs[0])
