import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def kat_pixels(v) -> np.ndarray:
    """Pixels of one known-answer vector of tests/golden/kat.json."""
    w, h, c = v["w"], v["h"], v["c"]
    if "pixels" in v:
        return np.array(v["pixels"], dtype=np.uint8).reshape(h, w, c)
    if "fill" in v:
        return np.full((h, w, c), v["fill"], dtype=np.uint8)
    a = np.zeros((h, w, c), dtype=np.uint8)
    for y in range(h):
        for x in range(w):
            if v["formula"] == "ramp4":       # px(x,y)=(16x+y, 8x+8y, 255-16x-y)
                a[y, x] = (16 * x + y, 8 * x + 8 * y, 255 - 16 * x - y)
            elif v["formula"] == "rgba3":     # px(x,y)=(40x,40y,20xy,200-x)
                a[y, x] = (40 * x, 40 * y, 20 * x * y, 200 - x)
            else:
                raise KeyError(v["formula"])
    return a


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        vs = json.load(f)["vectors"]
    return [(v["name"], kat_pixels(v), bytes.fromhex(v["stream"].replace(" ", ""))) for v in vs]


@pytest.fixture(scope="session")
def golden_streams():
    with open(os.path.join(GOLDEN, "streams.json")) as f:
        return json.load(f)
