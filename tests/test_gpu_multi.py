"""Multi-device entry points (llcomp_b200_multi_*) and the pipelined host-buffer
decode, through the C ABI.  The sharding logic does not care which physical device a shard lands on, so a box with one
GPU exercises it by naming device 0 several times; with two or more GPUs the same tests use distinct devices."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 4))))
    return lists


@pytest.fixture(scope="module")
def codec():
    import llcomp_b200
    return llcomp_b200.default_codec(0)


@pytest.mark.parametrize("tile", [(0, 0), (64, 32)])
def test_batch_sharded_by_images_equals_single_device(codec, tile):
    import llcomp_b200
    imgs = np.stack([oracle.generate(128, 96, 3, 6, 500 + k) for k in range(7)])
    want, want_off = codec.compress_batch(imgs, *tile)
    n = int(want_off[-1])
    for devs in device_lists():
        m = llcomp_b200.MultiCodec(devs)
        got, off = m.compress_batch(imgs, *tile)
        assert (off == want_off).all(), devs
        assert (got[:n] == want[:n]).all(), devs
        assert (m.decompress_batch(got, off) == imgs).all(), devs
        assert m.launch_count() > 0
        m.close()
    if tile == (0, 0):                                    # one slice per image: every stream is the reference's own
        for k in (0, 3, 6):
            assert want[int(want_off[k]):int(want_off[k + 1])].tobytes() == oracle.compress(imgs[k])


def test_single_image_sharded_by_tile_rows_equals_single_device(codec):
    """One 320x200 image in 64x48 tiles (5 tile rows, the last one ragged): bands of tile rows go to different
    devices, the container is the single-device container byte for byte, every payload is the oracle's tile coding."""
    import llcomp_b200
    img = oracle.generate(320, 200, 3, 5, 77)
    want = codec.compress(img, 320, 200, 3, 64, 48)
    for devs in device_lists():
        m = llcomp_b200.MultiCodec(devs)
        got = m.compress(img, 320, 200, 3, 64, 48)
        assert got == want, devs
        back = m.decompress(got)
        assert (back.pixels == img).all() and (back.width, back.height, back.channels) == (320, 200, 3)
        # truncated container: the missing tail reads as zero in every band, as on one device
        cut = got[: len(got) - 700]
        assert (m.decompress(cut).pixels == codec.decompress(cut).pixels).all()
        m.close()
    from test_gpu_parity import split_container, tiles_of
    _, _, _, _, _, payloads = split_container(want)
    for k, (x0, y0, sw, sh) in enumerate(tiles_of(320, 200, 64, 48)):
        if k in (0, 7, 24):
            assert payloads[k] == oracle.encode_tile(img, x0, y0, sw, sh)


def test_more_devices_than_work(codec):
    import llcomp_b200
    m = llcomp_b200.MultiCodec([0, 0, 0, 0])
    img = oracle.generate(40, 30, 3, 9, 5)
    s = m.compress(img, 40, 30, 3)                          # one slice: nothing to shard
    assert s == oracle.compress(img)
    assert (m.decompress(s).pixels == img).all()
    imgs = np.stack([img, img[::-1].copy()])
    out, off = m.compress_batch(imgs)                       # two images on four devices
    assert out[:int(off[1])].tobytes() == s
    assert (m.decompress_batch(out, off) == imgs).all()
    m.close()


def test_pipelined_host_decode_matches_oracle(codec):
    """16 images: the host-buffer decode runs in four groups on four streams (uploads, decoder launches and downloads
    of different groups overlap); sliced and unsliced, complete and truncated streams."""
    imgs = np.stack([oracle.generate(96, 80, 3, 7, 900 + k) for k in range(16)])
    for tile in ((0, 0), (32, 32)):
        buf, off = codec.compress_batch(imgs, *tile)
        assert (codec.decompress_batch(buf, off) == imgs).all()
    # reference streams, some of them truncated: same pixels as the oracle decodes (zero fill, llcomp.hpp:476-477)
    streams = [oracle.compress(imgs[k]) for k in range(16)]
    for k in (2, 9, 15):
        streams[k] = streams[k][: len(streams[k]) - 40 * (k + 1)]
    blob = np.frombuffer(b"".join(streams), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum([len(s) for s in streams])]).astype(np.uint64)
    got = codec.decompress_batch(blob, offs)
    for k in range(16):
        assert (got[k] == oracle.decompress(streams[k])).all(), k
