#!/bin/bash
# The BASELINE configs other than the headline one, through bench.py (one JSON line each) -> $1
out=${1:-gpurun_out/other_configs.jsonl}; : > $out
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 1"
$B --images 1 --size 4096 --tile 512 | tail -1 >> $out                      # configs[1]: 64 slices
$B --images 1 --size 8192 --channels 1 --noise -1 --tile 512 | tail -1 >> $out   # configs[2]: gray noise
for t in 2048 1024 512 256; do $B --images 1 --size 16384 --tile $t | tail -1 >> $out; done   # configs[4] slice sweep
$B --images 1024 --noise 0 | tail -1 >> $out                                # smooth batch
$B --images 1024 --noise 32 | tail -1 >> $out                               # +-32 noise
