#!/bin/bash
p=gpurun_out/r02_c17
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 > ${p}_bench.json 2> ${p}_bench.err; tail -c 400 ${p}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${p}_bench_reference.json 2>> ${p}_bench.err
python - <<PY
import json
d=json.loads(open("${p}_bench.json").read().strip().splitlines()[-1])
r=json.loads(open("${p}_bench_reference.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],1), "dec", round(d["decode"]["value"],3), "dec e2e", round(d["decode"]["e2e"]["value"],3), d["round_trip_exact"], "| reference", round(r["value"],4), "dec", round(r["decode"]["value"],4))
print({k["name"]: round(k["ms"],3) for k in d["kernels"]}); print(d.get("identity")); print(d["roofline"]["frac"], d["clocks"])
PY
