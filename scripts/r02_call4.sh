#!/bin/bash
# fp32 chain (v2) against the integer one (v0); state rows in shared memory against behind L1 at low slice counts
p=gpurun_out/r02_c4
B="timeout 300 python bench.py --no-e2e --no-cpu --no-decode --steps 3 --warmup 2"
run() { # name, env, images
  env $2 $B --images $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    print("$1", $3, "img:", round(d["value"],2), "GB/s", {k["name"]: round(k["ms"],3) for k in d["kernels"]})
except Exception as e: print("$1", "failed", e)
PY
}
V0=LLCOMP_B200_LIB=$PWD/llcomp_b200/lib/v0.so
V2=LLCOMP_B200_LIB=$PWD/llcomp_b200/lib/v2.so
env $V2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run v2_148 $V2 148
run v2_1024_old4 "$V2 LLCOMP_FUSED_NS=4" 1024
run v2_1024 $V2 1024
run v0_148_g1 "$V0 LLCOMP_FUSED_NS=1" 148
run v0_296_smem $V0 296
run v0_296_g1 "$V0 LLCOMP_FUSED_NS=1" 296
run v0_296_g2 "$V0 LLCOMP_FUSED_NS=2" 296
run v0_296_s3 "$V0 LLCOMP_FUSED_NS=13" 296
run v0_444_g1 "$V0 LLCOMP_FUSED_NS=1" 444
