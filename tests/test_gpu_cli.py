"""llcompc / llcompd (C++ host over the C ABI) keep the reference tools' contract (llcompc.cpp, llcompd.cpp):
one positional argument, `<image>.llcomp` next to the input, exit codes 0/1, reference error text."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "llcomp_b200", "host")


@pytest.fixture(scope="module")
def tools():
    from llcomp_b200.build import build
    build()
    return os.path.join(HOST, "llcompc"), os.path.join(HOST, "llcompd")


def write_pnm(path, img):
    h, w, c = img.shape
    with open(path, "wb") as f:
        if c in (1, 3):
            f.write(b"%s\n%d %d\n255\n" % (b"P5" if c == 1 else b"P6", w, h))
        else:
            f.write(b"P7\nWIDTH %d\nHEIGHT %d\nDEPTH %d\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n" % (w, h, c))
        f.write(img.tobytes())


def read_pnm_payload(path, n):
    with open(path, "rb") as f:
        return np.frombuffer(f.read()[-n:], np.uint8)


@pytest.mark.parametrize("c,ext", [(3, ".ppm"), (1, ".pgm"), (4, ".pam")])
def test_round_trip_through_the_tools(tools, tmp_path, c, ext):
    llcompc, llcompd = tools
    img = oracle.generate(200, 120, c, 4, 99)
    src = str(tmp_path / ("in" + ext))
    write_pnm(src, img)
    assert subprocess.run([llcompc, src]).returncode == 0
    with open(src + ".llcomp", "rb") as f:
        stream = f.read()
    assert stream == oracle.compress(img)                       # byte-identical to the reference format
    assert subprocess.run([llcompd, src + ".llcomp"]).returncode == 0
    assert (read_pnm_payload(src + ".llcomp" + ext, img.size) == img.reshape(-1)).all()


def test_tile_option_and_errors(tools, tmp_path):
    llcompc, llcompd = tools
    img = oracle.generate(300, 200, 3, 4, 5)
    src = str(tmp_path / "t.ppm")
    write_pnm(src, img)
    assert subprocess.run([llcompc, src, "--tile", "128x64"]).returncode == 0
    with open(src + ".llcomp", "rb") as f:
        assert f.read(1) == b"\xb2"
    assert subprocess.run([llcompd, src + ".llcomp"]).returncode == 0
    assert (read_pnm_payload(src + ".llcomp.ppm", img.size) == img.reshape(-1)).all()
    # reference behaviour: usage / missing file / bad magic all exit 1 (llcompc.cpp:19-29, llcompd.cpp:12-20, :32-35)
    assert subprocess.run([llcompc], capture_output=True).returncode == 1
    assert subprocess.run([llcompc, str(tmp_path / "nope.ppm")], capture_output=True).returncode == 1
    bad = str(tmp_path / "bad.llcomp")
    with open(bad, "wb") as f:
        f.write(bytes([0x77, 3, 1, 0, 1, 0, 0]))
    r = subprocess.run([llcompd, bad], capture_output=True, text=True)
    assert r.returncode == 1 and "Error decompressing image: Invalid magic number" in r.stderr


def test_many_files_are_coded_as_batches(tools, tmp_path):
    """More than one file: runs of equally sized images go through one batch call; every stream is still the
    reference's bytes for that image, and the decoder tool groups streams by header the same way."""
    llcompc, llcompd = tools
    imgs = [oracle.generate(96, 64, 3, 4, 10 + k) for k in range(5)] + [oracle.generate(50, 40, 3, 6, 77)] + \
           [oracle.generate(96, 64, 3, 0, 31)]
    srcs = []
    for k, img in enumerate(imgs):
        srcs.append(str(tmp_path / ("f%d.ppm" % k)))
        write_pnm(srcs[-1], img)
    assert subprocess.run([llcompc] + srcs).returncode == 0
    for src, img in zip(srcs, imgs):
        with open(src + ".llcomp", "rb") as f:
            assert f.read() == oracle.compress(img), src
    assert subprocess.run([llcompd] + [s + ".llcomp" for s in srcs]).returncode == 0
    for src, img in zip(srcs, imgs):
        assert (read_pnm_payload(src + ".llcomp.ppm", img.size) == img.reshape(-1)).all(), src
    # tiled containers batch too
    assert subprocess.run([llcompc] + srcs[:3] + ["--tile", "32x32"]).returncode == 0
    assert subprocess.run([llcompd] + [s + ".llcomp" for s in srcs[:3]]).returncode == 0
    for src, img in zip(srcs[:3], imgs[:3]):
        assert (read_pnm_payload(src + ".llcomp.ppm", img.size) == img.reshape(-1)).all(), src


def test_gpus_option_shards_without_changing_bytes(tools, tmp_path):
    """--devices 0,0 (two shards; works on a one-GPU box) and --gpus N: a batch goes image-wise, a tiled image by
    bands of tile rows; every output file is the same bytes as without the option."""
    llcompc, llcompd = tools
    imgs = [oracle.generate(96, 64, 3, 4, 40 + k) for k in range(5)]
    srcs = []
    for k, img in enumerate(imgs):
        srcs.append(str(tmp_path / ("g%d.ppm" % k)))
        write_pnm(srcs[-1], img)
    assert subprocess.run([llcompc] + srcs + ["--devices", "0,0"]).returncode == 0
    for src, img in zip(srcs, imgs):
        with open(src + ".llcomp", "rb") as f:
            assert f.read() == oracle.compress(img), src
    assert subprocess.run([llcompd] + [s + ".llcomp" for s in srcs] + ["--devices", "0,0"]).returncode == 0
    for src, img in zip(srcs, imgs):
        assert (read_pnm_payload(src + ".llcomp.ppm", img.size) == img.reshape(-1)).all(), src
    big = oracle.generate(320, 256, 3, 4, 3)
    src = str(tmp_path / "big.ppm")
    write_pnm(src, big)
    assert subprocess.run([llcompc, src, "--tile", "64x64"]).returncode == 0
    with open(src + ".llcomp", "rb") as f:
        want = f.read()
    assert subprocess.run([llcompc, src, "--tile", "64x64", "--devices", "0,0,0"]).returncode == 0
    with open(src + ".llcomp", "rb") as f:
        assert f.read() == want
    assert subprocess.run([llcompd, src + ".llcomp", "--devices", "0,0,0"]).returncode == 0
    assert (read_pnm_payload(src + ".llcomp.ppm", big.size) == big.reshape(-1)).all()
    import torch
    if torch.cuda.device_count() >= 2:
        assert subprocess.run([llcompc, src, "--tile", "64x64", "--gpus", "2"]).returncode == 0
        with open(src + ".llcomp", "rb") as f:
            assert f.read() == want


def test_unmodified_reference_tools_run_on_the_gpu_library(tmp_path):
    """llcomp_b200/host/_ref_cli/ holds the reference's own llcompc.cpp / llcompd.cpp, compiled UNCHANGED against
    llcomp_b200/host/llcomp.hpp (make -C llcomp_b200/host ref_cli; stb replaced by the PNM stubs of tests/stubs).
    They must behave like the reference tools: same stream bytes, exact pixels back."""
    c_tool, d_tool = (os.path.join(HOST, "_ref_cli", n) for n in ("llcompc", "llcompd"))
    if not (os.path.exists(c_tool) and os.path.exists(d_tool)):
        pytest.skip("reference CLIs were not built (no /root/reference at build time)")
    img = oracle.generate(160, 100, 3, 4, 2024)
    src = str(tmp_path / "r.ppm")
    write_pnm(src, img)
    assert subprocess.run([c_tool, src]).returncode == 0
    with open(src + ".llcomp", "rb") as f:
        assert f.read() == oracle.compress(img)
    assert subprocess.run([d_tool, src + ".llcomp"]).returncode == 0
    assert (read_pnm_payload(src + ".llcomp.png", img.size) == img.reshape(-1)).all()    # the stub writes PNM bytes
    bad = str(tmp_path / "bad.llcomp")
    with open(bad, "wb") as f:
        f.write(bytes([0x77, 3, 1, 0, 1, 0, 0]))
    r = subprocess.run([d_tool, bad], capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid magic number" in r.stderr
