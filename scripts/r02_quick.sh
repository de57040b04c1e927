#!/bin/bash
# GPU tests + one full bench line (no CPU arm): the check after a kernel change
p=${1:-gpurun_out/r02_quick}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu --steps 3 --warmup 2 > ${p}_bench.json 2> ${p}.err; tail -c 300 ${p}.err
python - <<PY
import json
d=json.loads(open("${p}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],1), "dec", round(d["decode"]["value"],3), "dec e2e", round(d["decode"]["e2e"]["value"],3), d["round_trip_exact"])
print({k["name"]: round(k["ms"],3) for k in d["kernels"]})
PY
