#!/bin/bash
# Store coalescing: microbenchmark, then the front end with staged (default) and direct (variant 1) stores; parity tests of the front end
p=gpurun_out/r02_fe2
true
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
B="timeout 600 python bench.py --no-cpu --no-e2e --no-decode --steps 3 --warmup 2"
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
run fe_staged LLCOMP_FRONTEND_VARIANT=0 ""
run fe_direct LLCOMP_FRONTEND_VARIANT=1 ""
run fe_legacy_staged LLCOMP_FRONTEND_VARIANT=4 ""
run fe_legacy_direct LLCOMP_FRONTEND_VARIANT=5 ""
run fe_c4_staged LLCOMP_FRONTEND_VARIANT=0 "--channels 4 --images 768"
run fe_c4_direct LLCOMP_FRONTEND_VARIANT=1 "--channels 4 --images 768"
tail -3 ${p}.err
