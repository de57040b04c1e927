"""Pins the CPU oracle (oracle/llcomp_oracle.c) before anything trusts it.

Checks, in order of authority:
  1. the unmodified reference header (oracle/_ref, when built in this container)
  2. the known-answer vectors of SURVEY.md appendix B (tests/golden/kat.json)
  3. committed golden streams generated from the unmodified reference
     (tests/golden/streams.json, small_fixtures.json; generator script beside them)
"""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference)")


# ---- constant tables vs the reference arrays (llcomp.hpp:252-281, :297-333) ----
@needs_ref
def test_tables_match_reference():
    L, R = oracle.lib(), oracle.ref()
    for s in range(128):
        assert L.llo_next_state_mps(s) == R.ref_next_state_mps(s)
        assert L.llo_next_state_lps(s) == R.ref_next_state_lps(s)
        assert L.llo_state_probability(s) == R.ref_state_probability(s)
    for x in range(-600, 601):
        assert L.llo_quant11(x) == R.ref_quant11(x)
        assert L.llo_quant5(x) == R.ref_quant5(x)
    assert R.ref_magic() == 0x79 and R.ref_states_nb() == 133104


@needs_ref
def test_median_matches_reference():
    L, R = oracle.lib(), oracle.ref()
    rng = np.random.default_rng(1)
    for a, b, c in rng.integers(-600, 600, size=(2000, 3)):
        assert L.llo_median(int(a), int(b), int(c)) == R.ref_median(int(a), int(b), int(c))


@needs_ref
def test_binarization_matches_reference():
    import ctypes as C
    R = oracle.ref()
    for d in range(-510, 511):
        buf = (C.c_uint8 * 40)()
        n = R.ref_binarize(d, buf)
        assert oracle.binarize(d) == [(buf[i] >> 1, buf[i] & 1) for i in range(n)]


def test_binarization_examples():
    # SURVEY.md section 8(a) row a9
    assert oracle.binarize(0) == [(0, 1)]
    assert oracle.binarize(-1) == [(0, 0), (1, 0), (7, 1)]
    assert oracle.binarize(-4) == [(0, 0), (1, 1), (2, 1), (3, 0), (5, 0), (6, 0), (7, 1)]
    assert len(oracle.binarize(510)) == 19


# ---- known-answer vectors -------------------------------------------------
def test_kat_encode(kats):
    for name, px, stream in kats:
        assert oracle.compress(px) == stream, name


def test_kat_decode(kats):
    for name, px, stream in kats:
        assert (oracle.decompress(stream) == px).all(), name


def test_kat_row_vs_column(kats):
    d = {n: s for n, _, s in kats}
    assert d["2x1 rgb"][6:] == d["1x2 rgb"][6:]      # w==0 => l=top rule, llcomp.hpp:417


@needs_ref
def test_kat_on_unmodified_reference(kats):
    for name, px, stream in kats:
        if len(stream) <= px.size:                   # D1: reference overflows otherwise
            assert oracle.ref_compress(px) == stream, name
            if px.shape[2] >= 3:
                assert (oracle.ref_decompress(stream) == px).all(), name


# ---- committed golden streams ---------------------------------------------
def test_golden_small_fixtures():
    with open(os.path.join(GOLDEN, "small_fixtures.json")) as f:
        fx = json.load(f)["fixtures"]
    assert any(x["source"] == "ref" for x in fx)
    for x in fx:
        px = np.array(x["pixels"], dtype=np.uint8).reshape(x["h"], x["w"], x["c"])
        s = bytes.fromhex(x["stream"])
        assert oracle.compress(px) == s
        assert (oracle.decompress(s) == px).all()


def test_golden_whole_streams(golden_streams):
    for g in golden_streams["whole"]:
        if g["w"] * g["h"] * g["c"] > 1024 * 1024 * 3 or (g["w"] == 1024 and g["n"] not in (0, 4)):
            continue                                  # keep the CPU suite short
        img = oracle.generate(g["w"], g["h"], g["c"], g["n"], g["seed"])
        s = oracle.compress(img)
        assert len(s) == g["bytes"]
        assert f"{oracle.fnv1a64(s):016x}" == g["fnv1a64"]


def test_golden_survey_lengths(golden_streams):
    # stream sizes listed in SURVEY.md appendix B (hashes there are not reproducible, sizes are)
    want = {(512, 512, 3, 4): 402823, (512, 512, 3, 0): 24048, (1024, 1024, 3, 0): 42489,
            (1024, 1024, 3, 2): 1232905, (1024, 1024, 3, 4): 1585151, (1024, 1024, 3, 8): 1974230,
            (1024, 1024, 3, 16): 2415570, (1024, 1024, 3, 32): 2900180, (1024, 1024, 4, 8): 2589361}
    got = {(g["w"], g["h"], g["c"], g["n"]): g["bytes"] for g in golden_streams["whole"]}
    for k, v in want.items():
        assert got[k] == v


def test_golden_tiles(golden_streams):
    g = next(t for t in golden_streams["tiled"] if t["tile_w"] == 256 and t["w"] == 600)
    img = oracle.generate(g["w"], g["h"], g["c"], g["n"], g["seed"])
    for t in g["tiles"]:
        tw = min(g["tile_w"], g["w"] - t["x0"])
        th = min(g["tile_h"], g["h"] - t["y0"])
        p = oracle.encode_tile(img, t["x0"], t["y0"], tw, th)
        assert len(p) == t["bytes"] and f"{oracle.fnv1a64(p):016x}" == t["fnv1a64"]
        assert (oracle.decode_tile(p, tw, th, g["c"]) == img[t["y0"]:t["y0"] + th, t["x0"]:t["x0"] + tw]).all()


# ---- differential vs the unmodified header --------------------------------
@needs_ref
def test_differential_random_images():
    rng = np.random.default_rng(7)
    n_ref = 0
    for k in range(120):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        c = int(rng.choice([1, 2, 3, 3, 3, 4]))
        amp = int(rng.choice([0, 1, 2, 4, 16, 64, 128]))
        base = rng.integers(0, 256, size=3)
        img = (np.add.outer(np.arange(h) * int(base[0] % 5), np.arange(w) * int(base[1] % 7))[:, :, None]
               + np.arange(c) * 11 + int(base[2]))
        img = np.clip(img % 256 + rng.integers(-amp, amp + 1, size=(h, w, c)), 0, 255).astype(np.uint8)
        s = oracle.compress(img)
        assert (oracle.decompress(s) == img).all()
        if len(s) <= img.size:
            n_ref += 1
            assert oracle.ref_compress(img) == s
            if c >= 3:
                assert (oracle.ref_decompress(s) == img).all()
    assert n_ref > 40


def test_frontend_then_coder_equals_encoder():
    img = oracle.generate(96, 80, 3, 6, 5)
    sym = oracle.frontend(img)
    assert sym.size == img.size
    assert oracle.encode_symbols(sym) == oracle.encode_tile(img)
    assert (sym >> 11).max() <= 7925


def test_tile_is_standalone_image():
    img = oracle.generate(200, 120, 3, 4, 9)
    tile = np.ascontiguousarray(img[40:100, 64:192])
    assert oracle.encode_tile(img, 64, 40, 128, 60) == oracle.compress(tile)[6:]


# ---- decoder edge behaviour -----------------------------------------------
def test_zero_fill_and_tail_sensitivity():
    img = oracle.generate(24, 24, 3, 8, 3)
    s = oracle.compress(img)
    assert (oracle.decompress(s + b"\x00" * 8) == img).all()   # zero fill == explicit zeros, llcomp.hpp:476-477
    k = len(s)
    while s[k - 1] == 0:
        k -= 1
    assert (oracle.decompress(s[:k]) == img).all()


def test_bad_magic():
    with pytest.raises(oracle.OracleError) as e:
        oracle.decompress(bytes([0x77, 3, 1, 0, 1, 0, 0, 0]))
    assert e.value.code == 1 and str(e.value) == "Invalid magic number"


def test_gray_and_noise_round_trip():
    for c in (1, 2):
        img = oracle.generate(64, 48, c, -1, 11)
        s = oracle.compress(img)
        assert len(s) > img.size                                  # D1 territory
        assert (oracle.decompress(s) == img).all()                # D2 territory


def test_generator_is_std_mt19937():
    # first outputs of std::mt19937(5489) are 3499211612, 581869302; G draws rng()%(2n+1)-n
    img = oracle.generate(2, 1, 1, 128, 5489)
    assert int(img[0, 0, 0]) == min(255, max(0, 0 + 3499211612 % 257 - 128))
    assert int(img[0, 1, 0]) == min(255, max(0, (1 * 255 // 2) // 2 + 581869302 % 257 - 128))
