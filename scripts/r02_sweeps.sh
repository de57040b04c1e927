#!/bin/bash
# The BASELINE configs other than the headline one, through bench.py (one JSON line each) -> $1
# configs[1] 4096^2 / 64 slices; configs[2] 8192^2 gray noise by tile grid; configs[4] slice-count sweep 1..4096 on 16384^2;
# entropy dependence of the headline batch.  Parity at these sizes: tests/test_gpu_fullsize.py.
out=${1:-gpurun_out/r02_sweeps.jsonl}; : > $out
B="timeout 900 python bench.py --no-cpu --no-e2e --scaling weak"
$B --steps 2 --warmup 1 --images 1 --size 4096 --tile 512 | tail -1 >> $out                                   # configs[1]
for t in 1024 512 256 128; do $B --steps 2 --warmup 1 --images 1 --size 8192 --channels 1 --noise -1 --tile $t | tail -1 >> $out; done   # configs[2]
for t in 16384 8192; do $B --steps 1 --warmup 0 --no-decode --images 1 --size 16384 --tile $t | tail -1 >> $out; done   # configs[4]: 1, 4 slices (encode; a 200 M-sample chain decodes for minutes)
$B --steps 1 --warmup 0 --images 1 --size 16384 --tile 4096 | tail -1 >> $out                                  # 16 slices
for t in 2048 1024 512 256; do $B --steps 2 --warmup 1 --images 1 --size 16384 --tile $t | tail -1 >> $out; done   # 64 .. 4096 slices
for n in 0 2 8 16 32; do $B --steps 2 --warmup 1 --images 1024 --strips 1 --noise $n | tail -1 >> $out; done    # entropy
python - <<PY
import json
for l in open("$out"):
    try: d=json.loads(l)
    except Exception: print("bad line", l[:80]); continue
    print(d["config"]["workload"][:70], "| slices", d["config"]["slices_per_gpu"], "| enc", round(d["value"],3), "dec", round(d["decode"]["value"],3) if d["decode"]["value"]==d["decode"]["value"] else None, "GB/s | bpp", round(d["bits_per_pixel"],4), d["round_trip_exact"])
PY
