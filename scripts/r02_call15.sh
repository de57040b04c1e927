#!/bin/bash
p=gpurun_out/r02_c15
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15
B="timeout 300 python bench.py --no-cpu --no-decode --no-e2e --steps 3 --warmup 2"
run() { # name, env, extra args
  env $2 $B $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["value"],2), "GB/s enc", {k["name"]: round(k["ms"],3) for k in d["kernels"]}, d["round_trip_exact"])
except Exception as e: print("$1", "failed", e)
PY
}
run pix_1024 X=1 ""
run rec_1024 LLCOMP_CODER_RECORDS=1 ""
run pix_592 X=1 "--images 592 --scaling weak --strips 1"
run rec_592 LLCOMP_CODER_RECORDS=1 "--images 592 --scaling weak --strips 1"
run pix_148 X=1 "--images 148 --scaling weak --strips 1"
run rec_148 LLCOMP_CODER_RECORDS=1 "--images 148 --scaling weak --strips 1"
run pix_296 X=1 "--images 296 --scaling weak --strips 1"
tail -3 ${p}.err
