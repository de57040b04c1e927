#!/bin/bash
# Front-end variants (K1 time), decoder after the micro-changes, GPU tests
p=gpurun_out/r02_fe
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="timeout 600 python bench.py --no-cpu --no-e2e --steps 3 --warmup 2"
run() { # name, env, extra args
  env $2 $B $3 > ${p}_$1.json 2>> ${p}.err
  python - <<PY
import json
try:
    d=json.loads(open("${p}_$1.json").read().strip().splitlines()[-1])
    print("$1", "enc", round(d["value"],2), {k["name"]: round(k["ms"],3) for k in d["kernels"]}, d["round_trip_exact"])
except Exception as e: print("$1", "failed", e)
PY
}
run fe_v0 LLCOMP_FRONTEND_VARIANT=0 "--no-decode"
run fe_v1 LLCOMP_FRONTEND_VARIANT=1 "--no-decode"
run fe_v2 LLCOMP_FRONTEND_VARIANT=2 "--no-decode"
run fe_v3 LLCOMP_FRONTEND_VARIANT=3 "--no-decode"
run fe_v4 LLCOMP_FRONTEND_VARIANT=4 "--no-decode"
run full X=1 ""
tail -3 ${p}.err
