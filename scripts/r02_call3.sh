#!/bin/bash
# chain recurrence variants (v0 shift, v1 IMAD.HI) and where the state rows live, by slices per SM
p=gpurun_out/r02_c3
B="timeout 300 python bench.py --no-e2e --no-cpu --no-decode --steps 3 --warmup 2"
run() { # name, env, images
  env $2 $B --images $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    print("$1", "$2", $3, "img:", round(d["value"],2), "GB/s", {k["name"]: round(k["ms"],3) for k in d["kernels"]}, d["clocks"])
except Exception as e: print("$1", "failed", e)
PY
}
V0=LLCOMP_B200_LIB=$PWD/llcomp_b200/lib/v0.so
V1=LLCOMP_B200_LIB=$PWD/llcomp_b200/lib/v1.so
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or kat or golden or alternate" 2>&1 | tail -2
run v0_148 $V0 148
run v1_148 $V1 148
run v1_148_g1 "$V1 LLCOMP_FUSED_NS=1" 148
run v1_296_g1 "$V1 LLCOMP_FUSED_NS=1" 296
run v1_296_g2 "$V1 LLCOMP_FUSED_NS=2" 296
run v1_296 $V1 296
run v1_592 $V1 592
run v0_1024_old4 "$V0 LLCOMP_FUSED_NS=4" 1024
run v1_1024_old4 "$V1 LLCOMP_FUSED_NS=4" 1024
run v1_1024 $V1 1024
