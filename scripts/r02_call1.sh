#!/bin/bash
# Round 2, first GPU call: parity of the new front end / solo coder, then stage times of the variants.
p=gpurun_out/r02_c1
timeout 1200 python -m pytest tests -m gpu -x -q > ${p}_pytest.log 2>&1; tail -3 ${p}_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-decode --steps 3 --warmup 2"
$B > ${p}_default.json 2> ${p}_default.err
LLCOMP_FUSED_NS=4 $B > ${p}_ns4_old.json 2>> ${p}_default.err
LLCOMP_FUSED_NS=14 $B > ${p}_ns4_solo.json 2>> ${p}_default.err
LLCOMP_FRONTEND_TILED=1 $B > ${p}_fe_tiled.json 2>> ${p}_default.err
for f in default ns4_old ns4_solo fe_tiled; do python - <<PY
import json
try:
    d=json.loads(open("${p}_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],2), "GB/s", round(d["ms_per_step"],2), "ms", {k["name"]: round(k["ms"],3) for k in d["kernels"]}, d["round_trip_exact"])
except Exception as e: print("$f", "failed", e)
PY
done
