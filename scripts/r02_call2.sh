#!/bin/bash
# Coder time against slices per SM (one CTA per SM, NS = ceil(n / 148)); old arrangement beside it.
p=gpurun_out/r02_c2
B="timeout 300 python bench.py --no-e2e --no-cpu --no-decode --steps 3 --warmup 2"
run() { # name, env, images
  env $2 $B --images $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    print("$1", "$2", $3, "img:", round(d["value"],2), "GB/s", {k["name"]: round(k["ms"],3) for k in d["kernels"]})
except Exception as e: print("$1", "failed", e)
PY
}
run n148 X=1 148
run n296 X=1 296
run n297 X=1 297
run n444 X=1 444
run n592 X=1 592
run n592_old LLCOMP_FUSED_NS=4 592
run n740 X=1 740
run n888 X=1 888
run n1036 X=1 1036
run n1024_ns2old LLCOMP_FUSED_NS=2 1024
