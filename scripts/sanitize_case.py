"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every default kernel once -- streaming front end,
fused coder (records from pixels and from the record array, several slices per CTA), scan, compaction, fast decoder --
checked against the oracle so that a "clean" run is also a correct one.
    compute-sanitizer --tool memcheck  python scripts/sanitize_case.py
    compute-sanitizer --tool racecheck python scripts/sanitize_case.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import llcomp_b200  # noqa: E402
import oracle  # noqa: E402

codec = llcomp_b200.Codec(0)
n_img, w, h, c, tw, th = 6, 64, 48, 3, 32, 16          # 6 x (2 x 3) = 36 slices: several slices per coder CTA
imgs = np.stack([oracle.generate(w, h, c, 6, 40 + k) for k in range(n_img)])
g = codec.geometry(w, h, c, tw, th, n_img)
d_px = torch.from_numpy(imgs).cuda()
for env in ({}, {"LLCOMP_CODER_PIXELS": "1"}, {"LLCOMP_FUSED_NS": "13"}, {"LLCOMP_FUSED_NS": "24"}):
    os.environ.update(env)
    codec.reload_switches()
    payload, offsets = codec.encode_device(d_px, g)
    codec.finish()
    off = offsets.cpu().numpy()
    k = 0
    for i in range(n_img):
        for y0 in range(0, h, th):
            for x0 in range(0, w, tw):
                got = payload[int(off[k]):int(off[k + 1])].cpu().numpy().tobytes()
                assert got == oracle.encode_tile(imgs[i], x0, y0, tw, th), (env, k)
                k += 1
    out = codec.decode_device(payload, offsets, g)
    codec.finish()
    assert torch.equal(out.view(imgs.shape), d_px), env
    for name in env:
        del os.environ[name]
codec.reload_switches()
sym = codec.frontend_device(d_px, g)                   # K1 alone (the streaming kernel: 64*3 % 16 == 0)
codec.finish()
buf, boff = codec.compress_batch(imgs)                 # host-buffer path, one slice per image
assert (codec.decompress_batch(buf, boff) == imgs).all()
assert buf[: int(boff[1])].tobytes() == oracle.compress(imgs[0])
print("sanitize_case ok:", codec.launch_count(), "kernel launches")
